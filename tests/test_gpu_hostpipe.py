"""GPU tests of the host-buffer entry point lfgpu_assemble_reaction_diffusion_host (pipelined upload / kernel / download).

It must return exactly what the three separate calls return (bitwise: same kernels, same order of operations inside a
row), for any node numbering, and match the oracle within the value tolerance of the path."""
import numpy as np
import pytest

from oracle import lfo
from tests.helpers import rel_max_err

pytestmark = pytest.mark.gpu

TOL = 1e-12


@pytest.fixture(scope="module")
def lf():
    import lehrfempp_b200 as lf
    return lf


@pytest.fixture(scope="module")
def ctx(lf):
    c = lf.Context(0)
    yield c
    c.close()


def moved(xy, seed):
    """interior-safe perturbation: small enough to keep every cell of a unit-ish grid non-degenerate"""
    rng = np.random.default_rng(seed)
    h = np.sqrt(np.ptp(xy[:, 0]) * np.ptp(xy[:, 1]) / len(xy))
    return xy + 0.15 * h * (rng.random(xy.shape) - 0.5)


def separate_calls(ctx, lf, gm, pat, degree, alpha, gamma, xy):
    gm.update_node_coords(xy)
    return pat.assemble_reaction_diffusion(degree, alpha, gamma).to_host()


@pytest.mark.parametrize("n_blocks", [0, 2, 7])
@pytest.mark.parametrize("pinned", [True, False])
def test_host_entry_equals_separate_calls_p1_fan(ctx, lf, n_blocks, pinned):
    gm = ctx.mesh_tp_tria(300, 200)
    dm = gm.dofmap_lagrange(1)
    pat = dm.symbolic(major=lf.ROW_MAJOR)
    xy = moved(gm.download()["node_coords"], 1)
    alpha, gamma = lf.Coeff.const(1.0), lf.Coeff.const(0.0)
    ref = separate_calls(ctx, lf, gm, pat, 1, alpha, gamma, xy)
    # start from different coordinates so that a missing upload shows
    gm.update_node_coords(moved(xy, 2))
    if pinned:
        h_xy = ctx.pinned(xy.size)
        h_xy[:] = xy.ravel()
        h_vals = ctx.pinned(pat.nnz)
    else:
        h_xy = np.ascontiguousarray(xy.ravel())
        h_vals = np.empty(pat.nnz)
    h_vals[:] = np.nan
    launches0 = ctx.kernel_launches
    dvals = pat.assemble_reaction_diffusion_host(1, alpha, gamma, h_xy, h_vals, n_blocks=n_blocks)
    assert ctx.kernel_launches > launches0
    assert np.array_equal(h_vals, ref)
    assert np.array_equal(dvals.to_host(), ref)
    assert np.array_equal(gm.download()["node_coords"], xy)


def test_host_entry_against_oracle(ctx, lf):
    om0 = lfo.Mesh.tp_tria(96, 64)
    ex = om0.export()
    xy = moved(ex["node_coords"], 3)
    om = lfo.Mesh.from_arrays(xy, ex["cell_nodes"])
    o_outer, o_inner, o_vals, _, _ = om.assemble_rd(1, lfo.coeff.const(2.0), lfo.coeff.const(0.5), csr=True)
    gm = ctx.mesh_upload(ex["node_coords"], ex["cell_nodes"])
    dm = gm.dofmap_lagrange(1)
    pat = dm.symbolic(major=lf.ROW_MAJOR)
    outer, inner = pat.download()
    assert np.array_equal(outer, o_outer) and np.array_equal(inner, o_inner)
    h_xy = ctx.pinned(xy.size)
    h_xy[:] = xy.ravel()
    h_vals = ctx.pinned(pat.nnz)
    pat.assemble_reaction_diffusion_host(1, lf.Coeff.const(2.0), lf.Coeff.const(0.5), h_xy, h_vals, n_blocks=3)
    assert rel_max_err(h_vals, o_vals) <= TOL


def test_host_entry_with_scattered_numbering(ctx, lf):
    # a node numbering without locality: every block needs (almost) all coordinates -- still the same result
    om0 = lfo.Mesh.tp_tria(128, 96)
    ex = om0.export()
    rng = np.random.default_rng(11)
    n = len(ex["node_coords"])
    perm = rng.permutation(n)            # new index of old node i
    xy = np.empty_like(ex["node_coords"])
    xy[perm] = ex["node_coords"]
    cn = ex["cell_nodes"].copy()
    cn[:, :3] = perm[cn[:, :3]]
    gm = ctx.mesh_upload(xy, cn)
    dm = gm.dofmap_lagrange(1)
    pat = dm.symbolic(major=lf.COL_MAJOR)
    xy2 = moved(xy, 5)
    alpha, gamma = lf.Coeff.const2x2([[2.0, 0.5], [0.25, 1.0]]), lf.Coeff.const(1.0)
    ref = separate_calls(ctx, lf, gm, pat, 1, alpha, gamma, xy2)
    gm.update_node_coords(xy)
    h_xy = ctx.pinned(xy2.size)
    h_xy[:] = xy2.ravel()
    h_vals = ctx.pinned(pat.nnz)
    pat.assemble_reaction_diffusion_host(1, alpha, gamma, h_xy, h_vals, n_blocks=3)
    assert np.array_equal(h_vals, ref)


@pytest.mark.parametrize("degree", [1, 2])
def test_host_entry_generic_path(ctx, lf, degree):
    # quads present: no fan kernel, the call degrades to upload -> assemble -> download
    gm = ctx.mesh_hybrid(40, 0.0, 1)
    dm = gm.dofmap_lagrange(degree)
    pat = dm.symbolic(major=lf.ROW_MAJOR)
    xy = moved(gm.download()["node_coords"], 7)
    alpha, gamma = lf.Coeff.const(1.0), lf.Coeff.const(3.0)
    ref = separate_calls(ctx, lf, gm, pat, degree, alpha, gamma, xy)
    gm.update_node_coords(moved(xy, 8))
    h_vals = np.empty(pat.nnz)
    pat.assemble_reaction_diffusion_host(degree, alpha, gamma, np.ascontiguousarray(xy.ravel()), h_vals)
    assert np.array_equal(h_vals, ref)


@pytest.mark.parametrize("parts", [2, 3])
def test_host_entry_row_ranges_tile_the_matrix(ctx, lf, parts):
    """The _range form (one GPU's share in the owner_rows mode): ranges that tile the rows reproduce the full result; each
    call uploads only the coordinate window its rows refer to."""
    from lehrfempp_b200.distributed import row_ranges
    import torch
    gm = ctx.mesh_tp_tria(260, 190)
    dm = gm.dofmap_lagrange(1)
    pat = dm.symbolic(major=lf.ROW_MAJOR)
    xy0 = gm.download()["node_coords"]
    xy = moved(xy0, 21)
    alpha, gamma = lf.Coeff.const(1.5), lf.Coeff.const(0.25)
    ref = separate_calls(ctx, lf, gm, pat, 1, alpha, gamma, xy)
    outer, _ = pat.download()
    bounds = row_ranges(torch.as_tensor(outer), parts)
    h_xy = ctx.pinned(xy.size)
    h_xy[:] = xy.ravel()
    got = np.full(pat.nnz, np.nan)
    for k in range(parts):
        r0, r1 = bounds[k], bounds[k + 1]
        # poison the device coordinates so that every range must bring its own window
        gm.update_node_coords((xy0 + 1.0) * 1e30)  # far away, but still a valid (non-degenerate) geometry
        ctx.synchronize()
        h_part = ctx.pinned(int(outer[r1] - outer[r0]))
        h_part[:] = np.nan
        pat.assemble_reaction_diffusion_host_range(1, alpha, gamma, h_xy, h_part, r0, r1 - r0, n_blocks=3)
        got[outer[r0]:outer[r1]] = h_part
        dev_xy = gm.download()["node_coords"]
        touched = np.flatnonzero(dev_xy[:, 0] < 1e29)
        assert touched.size < (0.75 if parts == 2 else 0.6) * len(xy)  # a window, not the whole array
        assert np.array_equal(dev_xy[touched], xy[touched])
    assert np.array_equal(got, ref)


def test_host_entry_optional_buffers(ctx, lf):
    gm = ctx.mesh_tp_tria(150, 150)
    dm = gm.dofmap_lagrange(1)
    pat = dm.symbolic()
    alpha, gamma = lf.Coeff.const(1.0), lf.Coeff.const(0.0)
    ref = pat.assemble_reaction_diffusion(1, alpha, gamma).to_host()
    h_vals = ctx.pinned(pat.nnz)
    h_vals[:] = 0.0
    pat.assemble_reaction_diffusion_host(1, alpha, gamma, None, h_vals, n_blocks=4)      # download only
    assert np.array_equal(h_vals, ref)
    xy = moved(gm.download()["node_coords"], 9)
    h_xy = ctx.pinned(xy.size)
    h_xy[:] = xy.ravel()
    dv = pat.assemble_reaction_diffusion_host(1, alpha, gamma, h_xy, None, n_blocks=4)   # upload only
    assert np.array_equal(dv.to_host(), separate_calls(ctx, lf, gm, pat, 1, alpha, gamma, xy))


@pytest.mark.parametrize("n_blocks", [1, 4])
def test_host_entry_checks_the_uploaded_geometry(ctx, lf, n_blocks):
    """Node positions that arrive through the host-buffer call are checked like those of lfgpu_mesh_upload (the reference asserts on
    a degenerate cell when the geometry object is built, tria_o1.cc:10-48): LFGPU_ERR_DEGENERATE, and the mesh recovers with the
    next valid upload."""
    gm = ctx.mesh_tp_tria(96, 96)
    pat = gm.dofmap_lagrange(1).symbolic(major=lf.ROW_MAJOR)
    xy = gm.download()["node_coords"].copy()
    alpha, gamma = lf.Coeff.const(1.0), lf.Coeff.const(0.0)
    h_vals = np.empty(pat.nnz)
    good = pat.assemble_reaction_diffusion_host(1, alpha, gamma, np.ascontiguousarray(xy.ravel()), h_vals, n_blocks=n_blocks).to_host()
    bad = xy.copy()
    bad[1000] = bad[1001]  # one collapsed edge
    with pytest.raises(lf.LfgpuError) as e:
        pat.assemble_reaction_diffusion_host(1, alpha, gamma, np.ascontiguousarray(bad.ravel()), h_vals, n_blocks=n_blocks)
    assert e.value.code == -5
    again = pat.assemble_reaction_diffusion_host(1, alpha, gamma, np.ascontiguousarray(xy.ravel()), h_vals, n_blocks=n_blocks).to_host()
    assert np.array_equal(good, again) and np.array_equal(h_vals, good)
