"""GPU parity of the P1 row kernel for quadrilaterals / hybrid meshes / variable coefficients / activity masks / cell corners that
differ from the node positions (lehrfempp_b200/csrc/assemble_p1h.cu; BASELINE config C2) against the oracle.

LFGPU_ALGO_FAN insists on a kernel that owns rows in registers (it fails with LFGPU_ERR_UNSUPPORTED instead of falling back to
the generic kernels), so every case below is known to have run in the new kernel.  Bars: values within 1e-12 relative in
max-norm; index arrays are those of the symbolic pass (bit-exact, tests/test_gpu_parity.py)."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import lfo
from tests.helpers import BUILTIN, per_qp_scalar, per_qp_tensor100, rel_max_err, upload_oracle_mesh
from tests.test_gpu_parity import gpu_mesh, oracle_mesh

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


def per_cell(ctx, lf, gm, fid):
    xy = gm.qp_coords(1, 4).to_host().reshape(gm.n_cells, 4, 2)
    vals = np.ascontiguousarray(BUILTIN[fid](xy[:, 0, 0], xy[:, 0, 1]))
    return lf.Coeff.per_cell(ctx.to_device(vals)), vals


def coefficient_cases(ctx, lf, gm):
    """(name, oracle alpha, oracle gamma, gpu alpha, gpu gamma) for every coefficient kind of include/lfgpu.h"""
    c = lfo.coeff
    A = [[3.0, 0.5], [1.0, 2.0]]
    ga1, _ = per_qp_scalar(ctx, gm, 1, 1)
    gg2, _ = per_qp_scalar(ctx, gm, 1, 2)
    gg4, _ = per_qp_scalar(ctx, gm, 1, 4)
    gt = per_qp_tensor100(ctx, gm, 1)
    pc1, v1 = per_cell(ctx, lf, gm, 1)
    pc2, v2 = per_cell(ctx, lf, gm, 2)
    return [
        ("const", c.const(1.5), c.const(0.75), lf.Coeff.const(1.5), lf.Coeff.const(0.75)),
        ("laplace", c.const(1.0), c.const(0.0), lf.Coeff.const(1.0), lf.Coeff.const(0.0)),
        ("const2x2", c.const2x2(A), c.const(0.25), lf.Coeff.const2x2(A), lf.Coeff.const(0.25)),
        ("per_cell", c.table(v1), c.table(v2), pc1, pc2),
        ("per_qp", c.builtin(1), c.builtin(2), ga1, gg2),
        ("per_qp_2x2", c.builtin(100), c.builtin(4), gt, gg4),
        ("mixed", c.builtin(1), c.const(0.0), ga1, lf.Coeff.const(0.0)),
    ]


MESHES = ["hybrid:8", "hybrid:9", "hybrid:10", "tp_quad:5", "tp_tria:6", "golden0", "golden1", "golden5", "golden8"]


@pytest.mark.parametrize("kind", MESHES)
@pytest.mark.parametrize("major", [0, 1])
def test_p1_rows_all_coefficient_kinds(ctx, lf, golden_meshes, kind, major):
    om = oracle_mesh(kind, golden_meshes)
    gm = gpu_mesh(ctx, kind, golden_meshes, om)
    pat = gm.dofmap_lagrange(1).symbolic(major=major)
    for name, oa, og, ga, gg in coefficient_cases(ctx, lf, gm):
        if name in ("const", "laplace", "const2x2") and gm.n_quad == 0 and kind.startswith("tp_tria"):
            continue  # the vertex-fan kernel's case (tests/test_gpu_parity.py)
        o = om.assemble_rd(1, oa, og, csr=(major == lf.ROW_MAJOR))
        vals = pat.assemble_reaction_diffusion(1, ga, gg, algo=lf.ALGO_FAN).to_host()
        assert rel_max_err(vals, o[2]) <= TOL, (kind, name)
        gen = pat.assemble_reaction_diffusion(1, ga, gg, algo=lf.ALGO_GATHER).to_host()
        assert rel_max_err(vals, gen) <= TOL, (kind, name)


@pytest.mark.parametrize("kind", ["hybrid:7", "tp_quad:4", "tp_tria:5", "golden0"])
def test_p1_rows_mask_and_accumulate(ctx, lf, golden_meshes, kind):
    om = oracle_mesh(kind, golden_meshes)
    gm = gpu_mesh(ctx, kind, golden_meshes, om)
    pat = gm.dofmap_lagrange(1).symbolic(major=lf.ROW_MAJOR)
    outer, inner = pat.download()
    rng = np.random.default_rng(11)
    active = (rng.random(om.n_cells) < 0.6).astype(np.uint8)
    ga, _ = per_qp_scalar(ctx, gm, 1, 1)
    gg, _ = per_qp_scalar(ctx, gm, 1, 2)
    o = om.assemble_rd(1, lfo.coeff.builtin(1), lfo.coeff.builtin(2), csr=True, active=active)
    dact = ctx.to_device(active)
    vals = pat.assemble_reaction_diffusion(1, ga, gg, active=dact, algo=lf.ALGO_FAN)
    Ao = sp.csr_matrix((o[2], o[1], o[0]), shape=o[3])
    Ag = sp.csr_matrix((vals.to_host(), inner, outer), shape=o[3])
    assert abs(Ao - Ag).max() <= TOL * np.abs(o[2]).max()
    # accumulate (assembler.h:84-88): a second call with beta = 1 adds; beta = -0.5 scales what is there first
    full = om.assemble_rd(1, lfo.coeff.builtin(1), lfo.coeff.builtin(2), csr=True)
    v = pat.assemble_reaction_diffusion(1, ga, gg, algo=lf.ALGO_FAN)
    pat.assemble_reaction_diffusion(1, ga, gg, beta=1.0, out=v, algo=lf.ALGO_FAN)
    assert rel_max_err(v.to_host(), 2 * full[2]) <= TOL
    pat.assemble_reaction_diffusion(1, ga, gg, beta=-0.5, out=v, algo=lf.ALGO_FAN)
    assert np.abs(v.to_host()).max() <= 1e-12 * np.abs(full[2]).max()


def test_p1_rows_cell_coords_differ_from_nodes(ctx, lf, golden_meshes):
    # a mesh whose cell geometries are NOT bitwise its node positions (what MeshHierarchy hands out, tria_o1.cc:99-151): the
    # kernel reads the corners from cell_coords, like the reference's Eval does through cell.Geometry()
    om = lfo.Mesh.hybrid(9, 0.2, 777)
    ex = om.export()
    rng = np.random.default_rng(3)
    cc = ex["cell_coords"].copy()
    cc += 1e-3 * rng.standard_normal(cc.shape) * (np.abs(cc) > 0)  # per-cell corners moved independently
    gm = ctx.mesh_upload(ex["node_coords"], ex["cell_nodes"], cc)
    om2 = lfo.Mesh.from_arrays(ex["node_coords"], ex["cell_nodes"], cell_coords=cc, cell_geo=np.ones(om.n_cells, np.uint8))
    pat = gm.dofmap_lagrange(1).symbolic(major=lf.ROW_MAJOR)
    ga, _ = per_qp_scalar(ctx, gm, 1, 1)
    gg, _ = per_qp_scalar(ctx, gm, 1, 2)
    o = om2.assemble_rd(1, lfo.coeff.builtin(1), lfo.coeff.builtin(2), csr=True)
    vals = pat.assemble_reaction_diffusion(1, ga, gg, algo=lf.ALGO_FAN).to_host()
    assert rel_max_err(vals, o[2]) <= TOL
    vals = pat.assemble_reaction_diffusion(1, lf.Coeff.const(2.0), lf.Coeff.const(0.0), algo=lf.ALGO_FAN).to_host()
    o = om2.assemble_rd(1, lfo.coeff.const(2.0), lfo.coeff.const(0.0), csr=True)
    assert rel_max_err(vals, o[2]) <= TOL


def test_p1_rows_user_rules(ctx, lf, golden_meshes):
    # explicit make_QuadRule(., 2) == the default (loc_comp_test.cc:86-152) runs in the row kernel; a rule with other point
    # counts keeps the generic kernels under AUTO and is refused under FAN
    om = lfo.Mesh.hybrid(8, 0.2, 12345)
    gm = ctx.mesh_hybrid(8, 0.2, 12345)
    pat = gm.dofmap_lagrange(1).symbolic(major=lf.ROW_MAJOR)
    qt, qq = lf.QuadRule(*lfo.quad_rule(3, 2)), lf.QuadRule(*lfo.quad_rule(4, 2))
    ga, _ = per_qp_scalar(ctx, gm, 1, 1, qt, qq)
    gg, _ = per_qp_scalar(ctx, gm, 1, 2, qt, qq)
    o = om.assemble_rd(1, lfo.coeff.builtin(1), lfo.coeff.builtin(2), qr_tria=2, qr_quad=2, csr=True)
    v = pat.assemble_reaction_diffusion(1, ga, gg, qt, qq, algo=lf.ALGO_FAN).to_host()
    assert rel_max_err(v, o[2]) <= TOL
    qt4, qq4 = lf.QuadRule(*lfo.quad_rule(3, 4)), lf.QuadRule(*lfo.quad_rule(4, 4))
    ga4, _ = per_qp_scalar(ctx, gm, 1, 1, qt4, qq4)
    gg4, _ = per_qp_scalar(ctx, gm, 1, 2, qt4, qq4)
    o4 = om.assemble_rd(1, lfo.coeff.builtin(1), lfo.coeff.builtin(2), qr_tria=4, qr_quad=4, csr=True)
    v4 = pat.assemble_reaction_diffusion(1, ga4, gg4, qt4, qq4, algo=lf.ALGO_AUTO).to_host()
    assert rel_max_err(v4, o4[2]) <= TOL
    with pytest.raises(lf.LfgpuError):
        pat.assemble_reaction_diffusion(1, ga4, gg4, qt4, qq4, algo=lf.ALGO_FAN)


def test_p1_rows_row_range_and_repeatability(ctx, lf):
    gm = ctx.mesh_hybrid(60, 0.2, 5)
    pat = gm.dofmap_lagrange(1).symbolic(major=lf.ROW_MAJOR)
    outer, _ = pat.download()
    ga, _ = per_qp_scalar(ctx, gm, 1, 1)
    gg, _ = per_qp_scalar(ctx, gm, 1, 2)
    full = pat.assemble_reaction_diffusion(1, ga, gg, algo=lf.ALGO_FAN).to_host()
    again = pat.assemble_reaction_diffusion(1, ga, gg, algo=lf.ALGO_FAN).to_host()
    assert np.array_equal(full, again)  # owner-computes: bitwise repeatable
    out = ctx.zeros(pat.nnz)
    r0, n = 517, 1999
    pat.assemble_reaction_diffusion_range(1, ga, gg, r0, n, out=out, algo=lf.ALGO_FAN)
    h = out.to_host()
    assert np.array_equal(h[outer[r0]:outer[r0 + n]], full[outer[r0]:outer[r0 + n]])
    assert not h[:outer[r0]].any() and not h[outer[r0 + n]:].any()


@pytest.mark.parametrize("n", [300, 1000])
def test_p1_rows_large_hybrid_against_generic_and_oracle(ctx, lf, n):
    # n = 1000: 1.5e6 cells -- more rows than one wave of resident CTAs, so the L2-prefetch branch of the kernel runs
    gm = ctx.mesh_hybrid(n, 0.2, 12345)
    pat = gm.dofmap_lagrange(1).symbolic(major=lf.ROW_MAJOR)
    ga, _ = per_qp_scalar(ctx, gm, 1, 1)
    gg, _ = per_qp_scalar(ctx, gm, 1, 2)
    rows = pat.assemble_reaction_diffusion(1, ga, gg, algo=lf.ALGO_FAN).to_host()
    gen = pat.assemble_reaction_diffusion(1, ga, gg, algo=lf.ALGO_GATHER).to_host()
    assert rel_max_err(rows, gen) <= TOL
    if n <= 300:
        om = lfo.Mesh.hybrid(n, 0.2, 12345)
        o = om.assemble_rd(1, lfo.coeff.builtin(1), lfo.coeff.builtin(2), csr=True)
        assert rel_max_err(rows, o[2]) <= TOL
    # size-independent: constants are in the kernel of the stiffness part, sum of the mass part = |Omega| = 1
    outer, inner = pat.download()
    N = pat.rows
    st = pat.assemble_reaction_diffusion(1, ga, lf.Coeff.const(0.0), algo=lf.ALGO_FAN).to_host()
    A = sp.csr_matrix((st, inner, outer), shape=(N, N))
    assert np.abs(A @ np.ones(N)).max() <= 1e-11 * np.abs(st).max()
    ms = pat.assemble_reaction_diffusion(1, lf.Coeff.const(0.0), lf.Coeff.const(1.0), algo=lf.ALGO_FAN).to_host()
    assert abs(ms.sum() - 1.0) <= 1e-11
