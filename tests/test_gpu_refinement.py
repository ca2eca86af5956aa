"""GPU tests on MeshHierarchy-refined meshes (BASELINE config 4 family): the device-side regular refinement reproduces the
oracle's refined mesh bit for bit (numbering, orientations, coordinates), and the assembly on refined meshes -- device
generated or uploaded from the host -- matches the oracle."""
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


def pair(ctx, kind):
    if kind == "tp_tria":
        return lfo.Mesh.tp_tria(7, 5, 0.0, 0.0, 2.0, 1.0), ctx.mesh_tp_tria(7, 5, 0.0, 0.0, 2.0, 1.0)
    if kind == "tp_quad":
        return lfo.Mesh.tp_quad(4, 6), ctx.mesh_tp_quad(4, 6)
    return lfo.Mesh.hybrid(7, 0.2, 5), ctx.mesh_hybrid(7, 0.2, 5)


def assert_same_mesh(gm, om):
    ex = om.export()
    d = gm.download(topology=True)
    assert (gm.n_nodes, gm.n_edges, gm.n_cells, gm.n_tria, gm.n_quad) == (om.n_nodes, om.n_edges, om.n_cells, om.n_tria, om.n_quad)
    assert np.array_equal(d["cell_nodes"], ex["cell_nodes"])
    assert np.array_equal(d["node_coords"].view(np.uint64), ex["node_coords"].view(np.uint64))
    assert np.array_equal(d["cell_coords"].view(np.uint64), ex["cell_coords"].view(np.uint64))
    assert np.array_equal(d["edge_nodes"], ex["edge_nodes"])
    assert np.array_equal(d["cell_edges"], ex["cell_edges"])
    assert np.array_equal(d["cell_edge_ori"], ex["cell_edge_ori"])


@pytest.mark.parametrize("kind", ["tp_tria", "tp_quad", "hybrid"])
@pytest.mark.parametrize("levels", [1, 2])
def test_device_refinement_bit_exact(ctx, kind, levels):
    om, gm = pair(ctx, kind)
    for _ in range(levels):
        om, gm = om.refine_regular(), gm.refine_regular()
    assert_same_mesh(gm, om)


@pytest.mark.parametrize("kind", ["tp_tria", "hybrid"])
@pytest.mark.parametrize("degree", [1, 3])
def test_assembly_on_refined_mesh(ctx, lf, kind, degree):
    """config 4: FeLagrangeO3 stiffness + mass on a refined mesh (and P1, which takes the fan kernel on triangles)"""
    om, gm = pair(ctx, kind)
    om, gm = om.refine_regular(), gm.refine_regular()
    dm = gm.dofmap_lagrange(degree)
    od, onl = om.cell_dofs(degree)
    gd, gnl = dm.download()
    assert np.array_equal(gd, od) and np.array_equal(gnl, onl)
    pat = dm.symbolic(major=lf.ROW_MAJOR)
    o_outer, o_inner, o_vals, _, _ = om.assemble_rd(degree, lfo.coeff.const(1.0), lfo.coeff.const(1.0), csr=True)
    outer, inner = pat.download()
    assert np.array_equal(outer, o_outer) and np.array_equal(inner, o_inner)
    vals = pat.assemble_reaction_diffusion(degree, lf.Coeff.const(1.0), lf.Coeff.const(1.0))
    assert rel_max_err(vals.to_host(), o_vals) <= TOL
    vec = dm.assemble_load(degree, lf.Coeff.const(1.0)).to_host()
    ov, _ = om.assemble_load(degree, lfo.coeff.const(1.0))
    assert rel_max_err(vec, ov) <= TOL


def test_uploaded_refined_mesh(ctx, lf, golden_meshes):
    """the host path: a refined reference mesh (explicit child geometries, all edges explicit) flattened and uploaded"""
    om = lfo.Mesh.from_golden(golden_meshes["0"]).refine_regular()
    gm = upload_oracle_mesh(ctx, om)[0]
    gm.build_topology(om.export()["edge_nodes"])
    dm = gm.dofmap_lagrange(3)
    od, onl = om.cell_dofs(3)
    gd, gnl = dm.download()
    assert np.array_equal(gd, od) and np.array_equal(gnl, onl)
    pat = dm.symbolic()
    o_outer, o_inner, o_vals, _, _ = om.assemble_rd(3, lfo.coeff.const(2.0), lfo.coeff.const(0.5))
    outer, inner = pat.download()
    assert np.array_equal(outer, o_outer) and np.array_equal(inner, o_inner)
    vals = pat.assemble_reaction_diffusion(3, lf.Coeff.const(2.0), lf.Coeff.const(0.5))
    assert rel_max_err(vals.to_host(), o_vals) <= TOL
