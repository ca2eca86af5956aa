"""Boundary flags of nodes and dofs on the device (lfgpu_mesh_boundary_nodes, lfgpu_dofmap_boundary_dofs) -- the selector the
reference's drivers build from flagEntitiesOnBoundary(mesh) and dofh.Entity(dof) (homDir_linfe_demo.cc:158-165) -- and the
device-resident pipeline it completes: assemble -> flag boundary dofs -> FixFlaggedSolutionComponents -> CG, against scipy on
the oracle's system.

Added after the round's GPU minutes were spent; sorts last on purpose."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from oracle import lfo

pytestmark = pytest.mark.gpu
LAYOUT = {1: (1, 0, 0, 0), 2: (1, 1, 0, 1), 3: (1, 2, 1, 4)}  # uniform_scalar_fe_space.h:334-341


@pytest.fixture(scope="module")
def lf():
    import lehrfempp_b200 as lf
    return lf


@pytest.fixture(scope="module")
def ctx(lf):
    c = lf.Context(0)
    yield c
    c.close()


def expected_flags(om, degree):
    ex = om.export()
    bd_edge = om.boundary_edges().astype(bool)
    bd_node = np.zeros(om.n_nodes, bool)
    bd_node[ex["edge_nodes"][bd_edge].ravel()] = True
    codim, idx = lfo.DofHandler(om, *LAYOUT[degree]).dof_entities()
    flags = np.zeros(codim.size, np.uint8)
    flags[codim == 2] = bd_node[idx[codim == 2]]
    flags[codim == 1] = bd_edge[idx[codim == 1]]
    return bd_node.astype(np.uint8), flags


@pytest.mark.parametrize("kind", ["tp_tria", "tp_quad", "hybrid"])
@pytest.mark.parametrize("degree", [1, 2, 3])
def test_boundary_flags_match_oracle(ctx, kind, degree):
    if kind == "tp_tria":
        om, gm = lfo.Mesh.tp_tria(6, 5), ctx.mesh_tp_tria(6, 5)
    elif kind == "tp_quad":
        om, gm = lfo.Mesh.tp_quad(4, 7), ctx.mesh_tp_quad(4, 7)
    else:
        om, gm = lfo.Mesh.hybrid(6, 0.2, 3), ctx.mesh_hybrid(6, 0.2, 3)
    dm = gm.dofmap_lagrange(degree)
    node_flags, dof_flags = expected_flags(om, degree)
    assert np.array_equal(gm.boundary_nodes().to_host()[: om.n_nodes], node_flags)
    assert np.array_equal(dm.boundary_dofs().to_host(), dof_flags)


def test_uploaded_dof_table_is_refused(ctx, lf):
    om = lfo.Mesh.tp_tria(3, 3)
    gm = ctx.mesh_tp_tria(3, 3)
    od, onl = om.cell_dofs(2)
    dm = gm.dofmap_upload(om.num_dofs(2), od, onl)
    with pytest.raises(lf.LfgpuError) as e:
        dm.boundary_dofs()
    assert e.value.code == -7


@pytest.mark.parametrize("degree", [1, 2])
def test_device_resident_poisson_pipeline(ctx, lf, degree):
    """-Laplace u = 1 on the unit square, u = 0 on the boundary: everything after mesh generation stays on the device."""
    n = 24
    om, gm = lfo.Mesh.tp_tria(n, n), ctx.mesh_tp_tria(n, n)
    dm = gm.dofmap_lagrange(degree)
    pat = dm.symbolic(major=lf.ROW_MAJOR)
    vals = pat.assemble_reaction_diffusion(degree, lf.Coeff.const(1.0), lf.Coeff.const(0.0))
    rhs = dm.assemble_load(degree, lf.Coeff.const(1.0))
    fixed = dm.boundary_dofs()
    pat.fix_flagged_solution_components(vals, rhs, fixed, ctx.zeros(dm.num_dofs))
    x, iters, res = pat.cg_solve(vals, rhs, rel_tol=1e-12)
    assert res <= 1e-12 and iters > 0
    # the oracle's system, eliminated on the host
    _, flags = expected_flags(om, degree)
    of = om.assemble_fixed(degree, 1.0, 0.0, 1.0, flags, np.zeros(flags.size), csr=True)
    A = sp.csr_matrix((of[2], of[1], of[0]), shape=(flags.size, flags.size))
    ref = spla.spsolve(A.tocsc(), of[3])
    xh = x.to_host()
    assert np.abs(xh - ref).max() <= 1e-9 * np.abs(ref).max()
    assert np.all(xh[flags == 1] == 0.0)
    assert abs(xh.max() - 0.0736713) < 2e-3  # max of the torsion function of the unit square


# ---- lf::fe::InitEssentialConditionFromFunction on the device (fe/fe_tools.h:301-356) ------------------------------------------
def oracle_dof_coords(om, degree):
    """Global(EvaluationNodes) of every cell, scattered through the cell's dof list (position b carries shape function b)."""
    ex = om.export()
    dofs, nl = om.cell_dofs(degree)
    xy = np.full((om.num_dofs(degree), 2), np.nan)
    nodes = {3: lfo.eval_fe(degree, 3, np.zeros((2, 1)))[2], 4: lfo.eval_fe(degree, 4, np.zeros((2, 1)))[2]}
    for c in range(om.n_cells):
        t = int(ex["cell_type"][c])
        p = ex["cell_coords"][c, :t]
        x0, x1 = nodes[t]
        if t == 3:
            w = np.stack([1 - x0 - x1, x0, x1])
        else:
            w = np.stack([(1 - x0) * (1 - x1), x0 * (1 - x1), x0 * x1, (1 - x0) * x1])
        xy[dofs[c, : nl[c]]] = (w.T @ p)
    return xy


@pytest.mark.parametrize("kind", ["tp_tria", "hybrid"])
@pytest.mark.parametrize("degree", [1, 2, 3])
def test_dof_positions_and_essential_condition(ctx, lf, kind, degree):
    if kind == "tp_tria":
        om, gm = lfo.Mesh.tp_tria(5, 4, 0.0, 0.0, 2.0, 1.0), ctx.mesh_tp_tria(5, 4, 0.0, 0.0, 2.0, 1.0)
    else:
        om, gm = lfo.Mesh.hybrid(5, 0.2, 9), ctx.mesh_hybrid(5, 0.2, 9)
    dm = gm.dofmap_lagrange(degree)
    want = oracle_dof_coords(om, degree)
    got = dm.dof_coords(degree)
    assert not np.isnan(want).any()
    assert np.abs(got - want).max() <= 1e-14
    # step 1 of InitEssentialConditionFromFunction with the boundary edges as selector = the boundary dofs
    flags = dm.edge_dof_flags(gm.boundary_edges()).to_host()
    assert np.array_equal(flags, dm.boundary_dofs().to_host())
    assert np.array_equal(flags, expected_flags(om, degree)[1])
    # a partial selector: only the edges on {y = 0}
    ex = om.export()
    mid_y = ex["node_coords"][ex["edge_nodes"]][:, :, 1]
    sel = (np.abs(mid_y).max(axis=1) < 1e-12).astype(np.uint8)
    part = dm.edge_dof_flags(ctx.to_device(sel)).to_host()
    assert part.sum() > 0 and np.all(np.abs(want[part == 1][:, 1]) < 1e-12) and np.all(part <= flags)
    on_line = np.abs(want[:, 1]) < 1e-12
    assert np.array_equal(part.astype(bool), on_line)
