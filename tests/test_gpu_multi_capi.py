"""Several GPUs from ONE process through the C ABI (lfgpu_multi_*, lehrfempp_b200/csrc/multi.cu): the call a C++ user of
AssembleMatrixLocally makes with a device list.  The handle may list the same device several times -- every entry gets its own
context and its own sub-problem -- so the whole path is checked on a single-GPU box; with two GPUs the same test also runs on
[0, 1].  Bars: the owned blocks of all devices together are the oracle's matrix, pattern bit-exact, values within 1e-12."""
import numpy as np
import pytest

from oracle import lfo
from tests.helpers import BUILTIN

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(scope="module")
def lf():
    import lehrfempp_b200 as lf
    return lf


def device_lists():
    import torch
    n = torch.cuda.device_count()
    lists = [[0, 0, 0], [0]]
    if n >= 2:
        lists.append([0, 1])
    if n >= 4:
        lists.append([0, 1, 2, 3])
    return lists


def check_against(multi, n_dev, o_outer, o_inner, o_vals, n_rows):
    covered = np.zeros(n_rows, np.int32)
    for k in range(n_dev):
        rows, ptr, cols, vals = multi.part(k)
        covered[rows] += 1
        lens = np.diff(ptr)
        assert np.array_equal(lens, o_outer[rows + 1] - o_outer[rows])
        idx_g = np.repeat(o_outer[rows] - ptr[:-1], lens) + np.arange(ptr[-1])
        assert np.array_equal(cols, o_inner[idx_g]), "pattern of an owned row differs from the reference's"
        assert np.abs(vals - o_vals[idx_g]).max() <= TOL * np.abs(o_vals).max()
    assert np.all(covered == 1)


@pytest.mark.parametrize("kind,degree", [("tria", 1), ("hybrid", 1), ("hybrid", 2), ("tria", 3)])
def test_multi_device_handle_against_the_oracle(lf, kind, degree):
    om = lfo.Mesh.tp_tria(33, 27) if kind == "tria" else lfo.Mesh.hybrid(22, 0.2, 12345)
    ex = om.export()
    dofs, nl = om.cell_dofs(degree)
    n_dofs = om.num_dofs(degree)
    for devs in device_lists():
        for major in (lf.ROW_MAJOR, lf.COL_MAJOR):
            multi = lf.MultiAssembler(devs)
            multi.setup(ex["node_coords"], ex["cell_nodes"], n_dofs, dofs, nl, major=major)
            o = om.assemble_rd(degree, lfo.coeff.const2x2([[2.0, 0.5], [-0.25, 1.0]]), lfo.coeff.const(0.75), csr=(major == lf.ROW_MAJOR))
            multi.assemble_reaction_diffusion(degree, lf.Coeff.const2x2([[2.0, 0.5], [-0.25, 1.0]]), lf.Coeff.const(0.75))
            check_against(multi, len(devs), o[0], o[1], o[2], n_dofs)
            # accumulate like the reference's void overload (assembler.h:84-88), then start again from zero
            multi.assemble_reaction_diffusion(degree, lf.Coeff.const2x2([[2.0, 0.5], [-0.25, 1.0]]), lf.Coeff.const(0.75), accumulate=True)
            check_against(multi, len(devs), o[0], o[1], 2 * o[2], n_dofs)
            multi.set_zero()
            multi.assemble_reaction_diffusion(degree, lf.Coeff.const(1.0), lf.Coeff.const(0.0), accumulate=True)
            o1 = om.assemble_rd(degree, lfo.coeff.const(1.0), lfo.coeff.const(0.0), csr=(major == lf.ROW_MAJOR))
            check_against(multi, len(devs), o1[0], o1[1], o1[2], n_dofs)
            sizes = [multi.part_sizes(k) for k in range(len(devs))]
            assert sum(s["rows"] for s in sizes) == n_dofs and sum(s["nnz"] for s in sizes) == o[2].size
            multi.close()


def test_multi_device_per_point_coefficients_from_host_tables(lf):
    # PER_QP tables are host arrays over the cells of the WHOLE mesh; every device receives the entries of its cells
    om = lfo.Mesh.hybrid(20, 0.2, 99)
    ex = om.export()
    dofs, nl = om.cell_dofs(1)
    ctx = lf.Context(0)
    gm = ctx.mesh_upload(ex["node_coords"], ex["cell_nodes"])
    xy = gm.qp_coords(1, 4).to_host().reshape(om.n_cells, 4, 2)
    a_tab = np.ascontiguousarray(BUILTIN[1](xy[..., 0], xy[..., 1]))
    g_tab = np.ascontiguousarray(BUILTIN[2](xy[..., 0], xy[..., 1]))
    del gm
    ctx.close()
    o = om.assemble_rd(1, lfo.coeff.builtin(1), lfo.coeff.builtin(2), csr=True)
    for devs in device_lists():
        multi = lf.MultiAssembler(devs)
        multi.setup(ex["node_coords"], ex["cell_nodes"], om.num_dofs(1), dofs, nl, major=lf.ROW_MAJOR)
        multi.assemble_reaction_diffusion(1, lf.Coeff.per_qp(a_tab, 4), lf.Coeff.per_qp(g_tab, 4))
        check_against(multi, len(devs), o[0], o[1], o[2], om.num_dofs(1))
        multi.close()
