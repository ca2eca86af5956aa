"""GPU tests of the consumer side: SpMV on both storage orders and the chain assemble -> edge terms -> Dirichlet
elimination -> conjugate gradients, against scipy on the oracle's system."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from oracle import lfo

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lf():
    import lehrfempp_b200 as lf
    return lf


@pytest.fixture(scope="module")
def ctx(lf):
    c = lf.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("degree", [1, 2, 3])
@pytest.mark.parametrize("csr", [True, False])
def test_spmv_matches_scipy(ctx, lf, degree, csr):
    gm = ctx.mesh_hybrid(11, 0.2, 7)
    dm = gm.dofmap_lagrange(degree)
    pat = dm.symbolic(major=lf.ROW_MAJOR if csr else lf.COL_MAJOR)
    # non-symmetric operator (tensor diffusion) so that a transposed product would show
    vals = pat.assemble_reaction_diffusion(degree, lf.Coeff.const2x2([[2.0, 1.0], [-0.5, 1.0]]), lf.Coeff.const(1.0))
    outer, inner = pat.download()
    n = dm.num_dofs
    A = (sp.csr_matrix if csr else sp.csc_matrix)((vals.to_host(), inner, outer), shape=(n, n))
    x = np.random.default_rng(degree).standard_normal(n)
    y = pat.spmv(vals, ctx.to_device(x)).to_host()
    ref = A @ x
    assert np.abs(y - ref).max() <= 1e-13 * np.abs(ref).max()


@pytest.mark.parametrize("degree,n", [(1, 24), (2, 24), (1, 96)])
def test_assemble_fix_solve_chain(ctx, lf, degree, n):
    """-div grad u + u = f on the unit square, u = g on the boundary: the whole chain on the device."""
    om, gm = lfo.Mesh.tp_tria(n, n), ctx.mesh_tp_tria(n, n)
    dm = gm.dofmap_lagrange(degree)
    pat = dm.symbolic(major=lf.ROW_MAJOR)
    vals = pat.assemble_reaction_diffusion(degree, lf.Coeff.const(1.0), lf.Coeff.const(1.0))
    rhs = dm.assemble_load(degree, lf.Coeff.const(2.0))
    # boundary dofs: the dofs of boundary edges (endpoints and edge-interior dofs), flagged through a unit edge load
    bd_mark = dm.assemble_edge_load(degree, lf.Coeff.const(1.0), active_edges=gm.boundary_edges()).to_host()
    fixed = (bd_mark > 0).astype(np.uint8)
    assert fixed.sum() == 4 * n * degree
    xhat = np.where(fixed == 1, 0.25, 0.0)
    o_outer, o_inner, o_vals, o_rhs = om.assemble_fixed(degree, 1.0, 1.0, 2.0, fixed, xhat, csr=True)
    pat.fix_flagged_solution_components(vals, rhs, ctx.to_device(fixed), ctx.to_device(xhat))
    x, iters, res = pat.cg_solve(vals, rhs, rel_tol=1e-12, max_iter=5000)
    N = dm.num_dofs
    ref = spla.spsolve(sp.csr_matrix((o_vals, o_inner, o_outer), shape=(N, N)).tocsc(), o_rhs)
    assert res <= 1e-12 and 0 < iters < 5000
    assert np.abs(x.to_host() - ref).max() <= 1e-9 * np.abs(ref).max()
    assert np.abs(x.to_host()[fixed == 1] - 0.25).max() <= 1e-12


def test_cg_rejects_indefinite_matrix(ctx, lf):
    gm = ctx.mesh_tp_tria(6, 6)
    dm = gm.dofmap_lagrange(1)
    pat = dm.symbolic()
    vals = pat.assemble_reaction_diffusion(1, lf.Coeff.const(-1.0), lf.Coeff.const(-1.0))  # negative definite
    rhs = dm.assemble_load(1, lf.Coeff.const(1.0))
    with pytest.raises(lf.LfgpuError):
        pat.cg_solve(vals, rhs, jacobi=False)
