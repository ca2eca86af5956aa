"""GPU parity of the P2 row kernels (lehrfempp_b200/csrc/assemble_p2.cu) against the oracle and against the generic kernels.

The row kernels take every cell with the row's own vertex / edge as local entity 0, so values differ from the generic
kernels by rounding only (bar of the path: 1e-12 relative in max-norm); irregular rows (boundary, valence != 6) come from
the generic gather kernel and must fit in seamlessly.
"""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import lfo
from tests.helpers import rel_max_err, upload_oracle_mesh

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


CASES = [
    ("laplace", lambda lf: lf.Coeff.const(1.0), lambda lf: lf.Coeff.const(0.0), lfo.coeff.const(1.0), lfo.coeff.const(0.0)),
    ("reaction_diffusion", lambda lf: lf.Coeff.const(2.5), lambda lf: lf.Coeff.const(0.75), lfo.coeff.const(2.5), lfo.coeff.const(0.75)),
    ("tensor", lambda lf: lf.Coeff.const2x2([[2.0, 0.5], [-0.25, 1.5]]), lambda lf: lf.Coeff.const(1.25),
     lfo.coeff.const2x2([[2.0, 0.5], [-0.25, 1.5]]), lfo.coeff.const(1.25)),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("major", ["row", "col"])
@pytest.mark.parametrize("shape", [(9, 10), (37, 23)])
def test_p2_rows_against_oracle(ctx, lf, case, major, shape):
    _, ga, gg, oa, og = case
    mj = lf.ROW_MAJOR if major == "row" else lf.COL_MAJOR
    om = lfo.Mesh.tp_tria(shape[0], shape[1], 0.25, -0.5, 1.75, 0.5)
    gm = ctx.mesh_tp_tria(shape[0], shape[1], 0.25, -0.5, 1.75, 0.5)
    pat = gm.dofmap_lagrange(2).symbolic(major=mj)
    o_outer, o_inner, o_vals, _, _ = om.assemble_rd(2, oa, og, csr=(mj == lf.ROW_MAJOR))
    outer, inner = pat.download()
    assert np.array_equal(outer, o_outer) and np.array_equal(inner, o_inner)
    n0 = ctx.kernel_launches
    vals = pat.assemble_reaction_diffusion(2, ga(lf), gg(lf), algo=lf.ALGO_FAN).to_host()
    assert rel_max_err(vals, o_vals) <= TOL
    # the generic kernel on the same input: same numbers up to rounding; AUTO is one of the two
    ref = pat.assemble_reaction_diffusion(2, ga(lf), gg(lf), algo=lf.ALGO_GATHER).to_host()
    assert rel_max_err(vals, ref) <= 1e-13
    auto = pat.assemble_reaction_diffusion(2, ga(lf), gg(lf)).to_host()
    assert rel_max_err(auto, o_vals) <= TOL
    assert ctx.kernel_launches > n0
    # deterministic: bitwise repeatable
    assert np.array_equal(vals, pat.assemble_reaction_diffusion(2, ga(lf), gg(lf), algo=lf.ALGO_FAN).to_host())


def test_p2_rows_overwrite_stale_values(ctx, lf):
    gm = ctx.mesh_tp_tria(12, 12)
    pat = gm.dofmap_lagrange(2).symbolic()
    out = ctx.to_device(np.full(pat.nnz, 123.456))
    vals = pat.assemble_reaction_diffusion(2, lf.Coeff.const(1.0), lf.Coeff.const(0.0), out=out, algo=lf.ALGO_FAN).to_host()
    ref = pat.assemble_reaction_diffusion(2, lf.Coeff.const(1.0), lf.Coeff.const(0.0), algo=lf.ALGO_GATHER).to_host()
    assert rel_max_err(vals, ref) <= 1e-13


def test_p2_rows_refined_mesh_numbering(ctx, lf):
    """MeshHierarchy-refined mesh: other node / edge numbering than the builder's, same kernels."""
    gm = ctx.mesh_tp_tria(5, 4).refine_regular().refine_regular()
    om = lfo.Mesh.tp_tria(5, 4).refine_regular().refine_regular()
    pat = gm.dofmap_lagrange(2).symbolic()
    o = om.assemble_rd(2, lfo.coeff.const(1.0), lfo.coeff.const(3.0), csr=True)
    outer, inner = pat.download()
    assert np.array_equal(outer, o[0]) and np.array_equal(inner, o[1])
    vals = pat.assemble_reaction_diffusion(2, lf.Coeff.const(1.0), lf.Coeff.const(3.0), algo=lf.ALGO_FAN).to_host()
    assert rel_max_err(vals, o[2]) <= TOL


def test_p2_rows_unstructured_mesh(ctx, lf):
    """Gmsh triangle mesh (valences 4..8, most vertex rows irregular): regular edge rows from the row kernel, the rest from
    the generic kernel -- or the generic kernel alone when too few rows qualify; either way the oracle's numbers."""
    from oracle.lfo_gmsh import GmshReader as OracleReader
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "msh", "circle_first_order.msh")
    xy, en, cn, _ = OracleReader(path).arrays()
    om = lfo.Mesh.from_arrays(xy, cn, edge_nodes=en)
    gm = lf.GmshReader(path).mesh(ctx)
    pat = gm.dofmap_lagrange(2).symbolic()
    o = om.assemble_rd(2, lfo.coeff.const(1.0), lfo.coeff.const(1.0), csr=True)
    vals = pat.assemble_reaction_diffusion(2, lf.Coeff.const(1.0), lf.Coeff.const(1.0)).to_host()
    assert rel_max_err(vals, o[2]) <= TOL


def test_p2_rows_not_taken_for_other_inputs(ctx, lf):
    """User rule, activity mask, accumulate, hybrid mesh: the generic kernels (no change of results)."""
    gm = ctx.mesh_tp_tria(8, 8)
    om = lfo.Mesh.tp_tria(8, 8)
    pat = gm.dofmap_lagrange(2).symbolic()
    o = om.assemble_rd(2, lfo.coeff.const(1.0), lfo.coeff.const(1.0), csr=True)
    act = (np.arange(gm.n_cells) % 4 != 1).astype(np.uint8)
    oa = om.assemble_rd(2, lfo.coeff.const(1.0), lfo.coeff.const(1.0), csr=True, active=act)  # pattern of the active cells only
    va = pat.assemble_reaction_diffusion(2, lf.Coeff.const(1.0), lf.Coeff.const(1.0), active=ctx.to_device(act)).to_host()
    outer, inner = pat.download()
    N = outer.size - 1
    diff = sp.csr_matrix((va, inner, outer), shape=(N, N)) - sp.csr_matrix((oa[2], oa[1], oa[0]), shape=(N, N))
    assert abs(diff).max() <= TOL * np.abs(oa[2]).max()
    out = pat.assemble_reaction_diffusion(2, lf.Coeff.const(1.0), lf.Coeff.const(1.0))
    pat.assemble_reaction_diffusion(2, lf.Coeff.const(1.0), lf.Coeff.const(1.0), beta=1.0, out=out)
    assert rel_max_err(out.to_host(), 2 * o[2]) <= TOL
    # round 2: accumulation stays in the row kernels (round 1 refused LFGPU_ALGO_FAN with beta != 0); the mask still does not
    pat.assemble_reaction_diffusion(2, lf.Coeff.const(1.0), lf.Coeff.const(1.0), beta=1.0, out=out, algo=lf.ALGO_FAN)
    assert rel_max_err(out.to_host(), 3 * o[2]) <= TOL
    with pytest.raises(lf.LfgpuError):
        pat.assemble_reaction_diffusion(2, lf.Coeff.const(1.0), lf.Coeff.const(1.0), active=ctx.to_device(act), algo=lf.ALGO_FAN)


def test_p2_rows_large_mesh_properties(ctx, lf):
    """Size-independent checks at 1.4e6 triangles: symmetry, zero row sums of the stiffness matrix, sum of the mass matrix
    = |Omega|, agreement with the generic kernel."""
    n = 850
    gm = ctx.mesh_tp_tria(n, n, 0.0, 0.0, 2.0, 1.0)
    pat = gm.dofmap_lagrange(2).symbolic()
    outer, inner = pat.download()
    N = outer.size - 1
    k = pat.assemble_reaction_diffusion(2, lf.Coeff.const(1.0), lf.Coeff.const(0.0), algo=lf.ALGO_FAN).to_host()
    K = sp.csr_matrix((k, inner, outer), shape=(N, N))
    scale = np.abs(k).max()
    assert abs(K - K.T).max() <= 1e-13 * scale
    assert np.abs(K @ np.ones(N)).max() <= 1e-12 * scale
    m = pat.assemble_reaction_diffusion(2, lf.Coeff.const(0.0), lf.Coeff.const(1.0), algo=lf.ALGO_FAN).to_host()
    assert abs(m.sum() - 2.0) <= 1e-11
    ref = pat.assemble_reaction_diffusion(2, lf.Coeff.const(1.0), lf.Coeff.const(0.0), algo=lf.ALGO_GATHER).to_host()
    assert rel_max_err(k, ref) <= 1e-13
