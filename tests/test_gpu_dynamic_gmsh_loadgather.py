"""GPU parity of the three late round-1 additions (30 tests green on B200, profiles/r01_gpu_tests_late_additions.log).

  * lfgpu_dofmap_dynamic      = lf::assemble::DynamicFEDofHandler (assemble/dofhandler.h:514-789), bit-exact dof tables
  * lfgpu_assemble_load(GATHER) = AssembleVectorLocally with the additions in the reference's order (assembler.h:322-324)
  * lfgpu_gmsh_mesh           = lf::io::GmshReader::mesh() on the device (io/gmsh_reader.cc), numbering bit-exact
"""
import os

import numpy as np
import pytest

from oracle import lfo
from oracle.lfo_gmsh import GmshReader as OracleReader
from tests.helpers import per_qp_scalar, rel_max_err, upload_oracle_mesh

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


def golden_pair(ctx, golden_meshes, sel):
    om = lfo.Mesh.from_golden(golden_meshes[sel])
    gm = upload_oracle_mesh(ctx, om)[0]
    entry = golden_meshes[sel]
    if "cells" in entry:
        gm.build_topology(cell_has_geometry=[c["coords"] is not None for c in entry["cells"]])
    else:
        gm.build_topology(om.export()["edge_nodes"])
    return om, gm


# ---- DynamicFEDofHandler ------------------------------------------------------------------------------------------------
def test_dynamic_layout_1234_equals_uniform(ctx, golden_meshes):
    om, gm = golden_pair(ctx, golden_meshes, "0")
    ex = om.export()
    cells = np.where(ex["cell_type"] == 3, 3, 4)
    dyn = gm.dofmap_dynamic(np.full(gm.n_nodes, 1), np.full(gm.n_edges, 2), cells)
    uni = gm.dofmap_uniform(1, 2, 3, 4)
    od, onl = lfo.DofHandler.dynamic(om, np.full(om.n_nodes, 1), np.full(om.n_edges, 2), cells).cell_dofs()
    for dm in (dyn, uni):
        gd, gnl = dm.download()
        assert dm.num_dofs == 10 + 36 + cells.sum()
        assert np.array_equal(gd, od) and np.array_equal(gnl, onl)


@pytest.mark.parametrize("sel", ["0", "1", "5"])
@pytest.mark.parametrize("seed", [1, 2])
def test_dynamic_variable_layout_bit_exact(ctx, golden_meshes, sel, seed):
    om, gm = golden_pair(ctx, golden_meshes, sel)
    rng = np.random.default_rng(seed)
    nn, ne, nc = rng.integers(0, 2, om.n_nodes), rng.integers(0, 3, om.n_edges), rng.integers(0, 3, om.n_cells)  # <= 14 per cell
    nn[0] = 1  # at least one dof
    odh = lfo.DofHandler.dynamic(om, nn, ne, nc)
    dm = gm.dofmap_dynamic(nn, ne, nc)
    od, onl = odh.cell_dofs()
    gd, gnl = dm.download()
    assert dm.num_dofs == odh.num_dofs and dm.stride == odh.stride
    assert np.array_equal(gnl, onl) and np.array_equal(gd, od)


def test_dynamic_nodes_only_and_edges_only(ctx, golden_meshes):
    om, gm = golden_pair(ctx, golden_meshes, "0")
    dm = gm.dofmap_dynamic(n_int_node=np.ones(gm.n_nodes))
    od, onl = lfo.DofHandler.dynamic(om, n_int_node=np.ones(om.n_nodes)).cell_dofs()
    gd, gnl = dm.download()
    assert dm.num_dofs == 10 and np.array_equal(gd, od) and np.array_equal(gnl, onl)
    dm = gm.dofmap_dynamic(n_int_edge=np.full(gm.n_edges, 2))
    od, onl = lfo.DofHandler.dynamic(om, n_int_edge=np.full(om.n_edges, 2)).cell_dofs()
    gd, gnl = dm.download()
    assert dm.num_dofs == 36 and np.array_equal(gd, od) and np.array_equal(gnl, onl)


def test_dynamic_structured_mesh_lagrange_layouts(ctx):
    """On the builder meshes a dynamic handler with the O2 / O3 Lagrange counts reproduces FeSpaceLagrangeO2/O3's table,
    and its pattern is the one of the uniform handler."""
    gm = ctx.mesh_tp_tria(7, 5)
    om = lfo.Mesh.tp_tria(7, 5)
    for degree, (n_seg, n_tri) in ((2, (1, 0)), (3, (2, 1))):
        dm = gm.dofmap_dynamic(np.ones(gm.n_nodes), np.full(gm.n_edges, n_seg), np.full(gm.n_cells, n_tri))
        od, onl = om.cell_dofs(degree)  # row length max(tria, quad) (dofhandler.cc:138); the dynamic table is as long as needed
        gd, gnl = dm.download()
        assert dm.stride == onl.max() and np.all(od[:, dm.stride:] == -1)
        assert np.array_equal(gd, od[:, :dm.stride]) and np.array_equal(gnl, onl)
        o1, i1 = dm.symbolic().download()
        o2, i2 = gm.dofmap_lagrange(degree).symbolic().download()
        assert np.array_equal(o1, o2) and np.array_equal(i1, i2)


def test_dynamic_too_many_local_dofs_rejected(ctx, lf):
    gm = ctx.mesh_tp_quad(2, 2)
    with pytest.raises(lf.LfgpuError) as e:
        gm.dofmap_dynamic(np.full(gm.n_nodes, 5))  # 20 dofs on a quadrilateral
    assert e.value.code == -7
    with pytest.raises(lf.LfgpuError) as e:
        gm.dofmap_dynamic(np.zeros(gm.n_nodes))
    assert e.value.code == -1


# ---- load vector, gather variant ----------------------------------------------------------------------------------------
def oracle_and_gpu(ctx, kind, golden_meshes):
    if kind.startswith("golden"):
        return golden_pair(ctx, golden_meshes, kind[6:])
    name, n = kind.split(":")
    n = int(n)
    if name == "tp_tria":
        return lfo.Mesh.tp_tria(n, n + 1, 0.25, -0.5, 1.75, 0.5), ctx.mesh_tp_tria(n, n + 1, 0.25, -0.5, 1.75, 0.5)
    return lfo.Mesh.hybrid(n, 0.2, 12345), ctx.mesh_hybrid(n, 0.2, 12345)


@pytest.mark.parametrize("kind", ["tp_tria:9", "hybrid:8", "golden0", "golden6"])
@pytest.mark.parametrize("degree", [1, 2, 3])
def test_load_vector_gather(ctx, lf, golden_meshes, kind, degree):
    om, gm = oracle_and_gpu(ctx, kind, golden_meshes)
    dm = gm.dofmap_lagrange(degree)
    gf, _ = per_qp_scalar(ctx, gm, degree, 3)
    ov, _ = om.assemble_load(degree, lfo.coeff.builtin(3))
    gv = dm.assemble_load(degree, gf, algo=lf.ALGO_GATHER).to_host()
    assert rel_max_err(gv, ov) <= TOL
    # same numbers as the atomic kernel up to the order of the additions
    ga = dm.assemble_load(degree, gf, algo=lf.ALGO_ATOMIC).to_host()
    assert rel_max_err(gv, ga) <= 1e-14
    # bitwise repeatable
    assert np.array_equal(gv, dm.assemble_load(degree, gf, algo=lf.ALGO_GATHER).to_host())
    # accumulate on top (assembler.h:291-293: the vector is not zeroed) and the activity mask (isActive)
    out = dm.assemble_load(degree, lf.Coeff.const(2.0), algo=lf.ALGO_GATHER)
    dm.assemble_load(degree, lf.Coeff.const(2.0), beta=1.0, out=out, algo=lf.ALGO_GATHER)
    ov2, _ = om.assemble_load(degree, lfo.coeff.const(2.0))
    assert rel_max_err(out.to_host(), 2 * ov2) <= TOL
    act = (np.arange(om.n_cells) % 3 != 0).astype(np.uint8)
    ov3, _ = om.assemble_load(degree, lfo.coeff.const(1.5), active=act)
    gv3 = dm.assemble_load(degree, lf.Coeff.const(1.5), active=ctx.to_device(act), algo=lf.ALGO_GATHER).to_host()
    assert rel_max_err(gv3, ov3) <= TOL


def test_load_vector_gather_larger_mesh_properties(ctx, lf):
    """Size-independent checks at 2.9e6 triangles: sum of the load vector of f = 1 is |Omega|; gather == atomic."""
    gm = ctx.mesh_tp_tria(1200, 1200, 0.0, 0.0, 2.0, 1.0)
    for degree in (1, 2):
        dm = gm.dofmap_lagrange(degree)
        v = dm.assemble_load(degree, lf.Coeff.const(1.0), algo=lf.ALGO_GATHER).to_host()
        assert abs(v.sum() - 2.0) <= 1e-11
        a = dm.assemble_load(degree, lf.Coeff.const(1.0), algo=lf.ALGO_ATOMIC).to_host()
        assert rel_max_err(v, a) <= 1e-13


# ---- Gmsh input -> device mesh (lfgpu_gmsh_mesh = reader.mesh()) ------------------------------------------------------------
MSH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "msh")


@pytest.mark.parametrize("name", ["two_element_hybrid_2d.msh", "two_element_hybrid_2d_v4_binary.msh", "lecturedemomesh.msh",
                                  "circle_first_order.msh", "circle_first_order_v4.msh", "piece_of_cake.msh"])
def test_gmsh_mesh_numbering_and_assembly(ctx, lf, name):
    path = os.path.join(MSH, name)
    o = OracleReader(path)
    xy, en, cn, _ = o.arrays()
    om = lfo.Mesh.from_arrays(xy, cn, edge_nodes=en)
    g = lf.GmshReader(path)
    gm = g.mesh(ctx)
    assert (gm.n_nodes, gm.n_edges, gm.n_cells) == (om.n_nodes, om.n_edges, om.n_cells)
    d, e = gm.download(topology=True), om.export()
    for key in ("cell_nodes", "edge_nodes", "cell_edges", "cell_edge_ori"):
        assert np.array_equal(d[key], e[key]), key
    assert np.array_equal(d["node_coords"], e["node_coords"])
    # P2 reaction-diffusion on the mesh: dof numbers depend on the explicit-edge numbering of the reader
    dm = gm.dofmap_lagrange(2)
    od, onl = om.cell_dofs(2)
    gd, gnl = dm.download()
    assert np.array_equal(gd, od) and np.array_equal(gnl, onl)
    pat = dm.symbolic(major=lf.COL_MAJOR)
    oo = om.assemble_rd(2, lfo.coeff.const(1.5), lfo.coeff.const(0.5))
    outer, inner = pat.download()
    assert np.array_equal(outer, oo[0]) and np.array_equal(inner, oo[1])
    vals = pat.assemble_reaction_diffusion(2, lf.Coeff.const(1.5), lf.Coeff.const(0.5)).to_host()
    assert rel_max_err(vals, oo[2]) <= TOL
    # Dirichlet-type selector from a physical group: flags on the edges of the device mesh
    for nr, _name in g.physical_entities(1):
        flags = g.physical_flags(1, nr, gm.n_edges)
        want = [o.is_physical_entity(1, i, nr) for i in range(gm.n_edges)]
        assert list(flags) == [int(w) for w in want]


def test_gmsh_second_order_mesh_is_rejected_on_the_device(ctx, lf):
    g = lf.GmshReader(os.path.join(MSH, "circle_second_order.msh"))
    assert g.geometry_order == 2
    with pytest.raises(lf.LfgpuError) as e:
        g.mesh(ctx)
    assert e.value.code == -7


# ---- two-pass load vector (LFGPU_ALGO_AUTO): element vectors once per cell, then added per dof in the reference's order ------------
@pytest.mark.parametrize("kind", ["tp_tria:9", "hybrid:8", "golden0", "golden6"])
@pytest.mark.parametrize("degree", [1, 2, 3])
def test_load_vector_two_pass(ctx, lf, golden_meshes, kind, degree):
    om, gm = oracle_and_gpu(ctx, kind, golden_meshes)
    dm = gm.dofmap_lagrange(degree)
    gf, _ = per_qp_scalar(ctx, gm, degree, 3)
    ov, _ = om.assemble_load(degree, lfo.coeff.builtin(3))
    gv = dm.assemble_load(degree, gf).to_host()
    assert rel_max_err(gv, ov) <= TOL
    # the same additions in the same order as the one-pass gather kernel; bitwise repeatable
    gg = dm.assemble_load(degree, gf, algo=lf.ALGO_GATHER).to_host()
    assert rel_max_err(gv, gg) <= 1e-14  # (P1 on triangles: the ring kernel adds in ring order, everything else is bitwise the gather kernel's)
    assert np.array_equal(gv, dm.assemble_load(degree, gf).to_host())
    out = dm.assemble_load(degree, gf)
    dm.assemble_load(degree, gf, beta=1.0, out=out)
    assert rel_max_err(out.to_host(), 2 * ov) <= TOL
    act = (np.arange(om.n_cells) % 3 != 0).astype(np.uint8)
    ov3, _ = om.assemble_load(degree, lfo.coeff.builtin(3), active=act)
    gv3 = dm.assemble_load(degree, gf, active=ctx.to_device(act)).to_host()
    assert rel_max_err(gv3, ov3) <= TOL


def test_load_vector_two_pass_large(ctx, lf):
    """1.0e6 hybrid cells / 2.9e6 triangles, P2 and P3: against the atomic kernel; sum of the load vector of f = 1 is |Omega|"""
    for gm, area in ((ctx.mesh_hybrid(816, 0.2, 7), 1.0), (ctx.mesh_tp_tria(1200, 1200, 0.0, 0.0, 2.0, 1.0), 2.0)):
        for degree in (2, 3):
            dm = gm.dofmap_lagrange(degree)
            v = dm.assemble_load(degree, lf.Coeff.const(1.0)).to_host()
            assert abs(v.sum() - area) <= 1e-11
            a = dm.assemble_load(degree, lf.Coeff.const(1.0), algo=lf.ALGO_ATOMIC).to_host()
            assert rel_max_err(v, a) <= 1e-13
