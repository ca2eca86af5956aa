"""GPU parity of the edge (codim-1) path: MassEdgeMatrixProvider / ScalarLoadEdgeVectorProvider + AssembleXLocally(1, ...)
(uscalfe/loc_comp_ellbvp.h:367-529, 784-921) against the oracle.  Index arrays bit-exact (the edge entries live inside
the cell pattern), values within 1e-12 of the max entry."""
import numpy as np
import pytest

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


def meshes(ctx, kind):
    if kind == "hybrid":
        return lfo.Mesh.hybrid(9, 0.2, 12345), ctx.mesh_hybrid(9, 0.2, 12345)
    if kind == "tp_tria":
        return lfo.Mesh.tp_tria(12, 9, 0.0, 0.0, 2.0, 1.0), ctx.mesh_tp_tria(12, 9, 0.0, 0.0, 2.0, 1.0)
    return lfo.Mesh.tp_quad(7, 11), ctx.mesh_tp_quad(7, 11)


@pytest.mark.parametrize("kind", ["hybrid", "tp_tria", "tp_quad"])
def test_boundary_edge_flags(ctx, kind):
    om, gm = meshes(ctx, kind)
    assert np.array_equal(gm.boundary_edges().to_host(), om.boundary_edges())


def test_boundary_flags_on_golden_mesh(ctx, golden_meshes):
    om = lfo.Mesh.from_golden(golden_meshes["0"])
    gm = upload_oracle_mesh(ctx, om)[0]
    f = gm.boundary_edges().to_host()
    assert np.flatnonzero(f).tolist() == [11, 12, 13, 14, 15, 16, 17]  # assembly_tests.cc:552-558: rows 3..9 of the golden


@pytest.mark.parametrize("kind", ["hybrid", "tp_tria", "tp_quad"])
@pytest.mark.parametrize("degree", [1, 2, 3])
@pytest.mark.parametrize("csr", [False, True])
def test_cell_plus_edge_matrix_matches_oracle(ctx, lf, kind, degree, csr):
    om, gm = meshes(ctx, kind)
    c = lfo.coeff
    bd = om.boundary_edges()
    o_outer, o_inner, o_vals = om.assemble_rd_edge(degree, c.const(1.5), c.const(0.5), c.const(2.0), edge_mask=bd, csr=csr)
    dm = gm.dofmap_lagrange(degree)
    pat = dm.symbolic(major=lf.ROW_MAJOR if csr else lf.COL_MAJOR)
    vals = pat.assemble_reaction_diffusion(degree, lf.Coeff.const(1.5), lf.Coeff.const(0.5))
    pat.assemble_edge_mass(dm, degree, lf.Coeff.const(2.0), vals, active_edges=gm.boundary_edges())
    outer, inner = pat.download()
    assert np.array_equal(outer, o_outer) and np.array_equal(inner, o_inner)
    assert rel_max_err(vals.to_host(), o_vals) <= TOL


@pytest.mark.parametrize("degree", [1, 2, 3])
def test_edge_terms_with_variable_coefficients(ctx, lf, degree):
    """eta(x) = 1 + |x|^2 and g(x) = sin(2 pi x) sin(2 pi y) through per-point tables (what the shim builds from a
    MeshFunctionGlobal), on ALL edges (no selector)."""
    om, gm = meshes(ctx, "hybrid")
    c = lfo.coeff
    dm = gm.dofmap_lagrange(degree)
    pat = dm.symbolic(major=lf.ROW_MAJOR)
    xy = gm.edge_qp_coords(degree)
    eta = np.array([[lfo.builtin_scalar(1, x, y) for (x, y) in e] for e in xy])
    g = np.array([[lfo.builtin_scalar(3, x, y) for (x, y) in e] for e in xy])
    d_eta, d_g = ctx.to_device(eta.ravel()), ctx.to_device(g.ravel())
    vals = pat.assemble_reaction_diffusion(degree, lf.Coeff.const(1.0), lf.Coeff.const(0.0))
    pat.assemble_edge_mass(dm, degree, lf.Coeff.per_qp(d_eta, eta.shape[1]), vals)
    o_outer, o_inner, o_vals = om.assemble_rd_edge(degree, c.const(1.0), c.const(0.0), c.builtin(1), csr=True)
    assert rel_max_err(vals.to_host(), o_vals) <= TOL
    vec = dm.assemble_load(degree, lf.Coeff.const(1.0))
    dm.assemble_edge_load(degree, lf.Coeff.per_qp(d_g, g.shape[1]), out=vec)
    ov, _ = om.assemble_load(degree, c.const(1.0))
    om.assemble_edge_load(degree, c.builtin(3), out=ov)
    assert rel_max_err(vec.to_host(), ov) <= TOL


def test_edge_load_per_edge_table_and_custom_rule(ctx, lf):
    om, gm = meshes(ctx, "tp_tria")
    degree = 2
    dm = gm.dofmap_lagrange(degree)
    bd = gm.boundary_edges()
    per_edge = np.linspace(0.5, 2.0, gm.n_edges)
    qr = lf.default_quad_rule(2, 9)  # 5-point Gauss rule on the segment
    vec = dm.assemble_edge_load(degree, lf.Coeff.per_cell(ctx.to_device(per_edge)), qr_segment=qr, active_edges=bd)
    ov = om.assemble_edge_load(degree, lfo.coeff.table(per_edge), edge_mask=om.boundary_edges(), qr_degree=9)
    assert rel_max_err(vec.to_host(), ov) <= TOL
    # accumulate semantics (assembler.h:291-293)
    dm.assemble_edge_load(degree, lf.Coeff.per_cell(ctx.to_device(per_edge)), qr_segment=qr, active_edges=bd, out=vec)
    assert rel_max_err(vec.to_host(), 2 * ov) <= TOL


def test_edge_mass_needs_device_built_lagrange_dofs(ctx, lf, golden_meshes):
    om = lfo.Mesh.from_golden(golden_meshes["0"])
    gm = upload_oracle_mesh(ctx, om)[0]
    od, onl = om.cell_dofs(2)
    dm = gm.dofmap_upload(om.num_dofs(2), od, onl)  # uploaded table: edge dofs unknown
    pat = dm.symbolic()
    with pytest.raises(lf.LfgpuError) as e:
        pat.assemble_edge_mass(dm, 2, lf.Coeff.const(1.0), ctx.zeros(pat.nnz))
    assert e.value.code == -7
