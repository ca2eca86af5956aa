"""Load vector on the vertex rings (lehrfempp_b200/csrc/assemble_p1.cu: k_load_p1_fan; LFGPU_ALGO_AUTO for P1 on triangles with a
constant source) against the oracle's AssembleVectorLocally + ScalarLoadElementVectorProvider and against the gather kernel."""
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


@pytest.mark.parametrize("nx,ny", [(1, 1), (2, 3), (37, 29), (128, 64)])
def test_load_fan_structured(ctx, lf, nx, ny):
    om = lfo.Mesh.tp_tria(nx, ny, 0.25, -0.5, 1.75, 0.5)
    gm = ctx.mesh_tp_tria(nx, ny, 0.25, -0.5, 1.75, 0.5)
    dm = gm.dofmap_lagrange(1)
    ov, _ = om.assemble_load(1, lfo.coeff.const(2.5))
    v = dm.assemble_load(1, lf.Coeff.const(2.5))
    assert rel_max_err(v.to_host(), ov) <= TOL
    dm.assemble_load(1, lf.Coeff.const(2.5), beta=1.0, out=v)  # assembler.h:291-293: the vector is not zeroed
    assert rel_max_err(v.to_host(), 2 * ov) <= TOL
    again = dm.assemble_load(1, lf.Coeff.const(2.5)).to_host()
    assert np.array_equal(again, dm.assemble_load(1, lf.Coeff.const(2.5)).to_host())  # bitwise repeatable
    # explicit rule of another degree (same lhat for every local index: the kernel applies)
    q, qq = lf.QuadRule(*lfo.quad_rule(3, 4)), lf.QuadRule(*lfo.quad_rule(4, 4))
    ov4, _ = om.assemble_load(1, lfo.coeff.const(1.0), qr_tria=4, qr_quad=4)
    assert rel_max_err(dm.assemble_load(1, lf.Coeff.const(1.0), qr_tria=q, qr_quad=qq).to_host(), ov4) <= TOL


@pytest.mark.parametrize("sel", ["6", "4"])
def test_load_fan_on_reference_test_meshes(ctx, lf, golden_meshes, sel):
    # pure triangle meshes of GenerateHybrid2DTestMesh: irregular valences, boundary fans
    om = lfo.Mesh.from_golden(golden_meshes[sel])
    if om.n_quad:
        pytest.skip("triangle meshes only")
    gm = upload_oracle_mesh(ctx, om)[0]
    gm.build_topology(om.export()["edge_nodes"])
    dm = gm.dofmap_lagrange(1)
    ov, _ = om.assemble_load(1, lfo.coeff.const(1.0))
    assert rel_max_err(dm.assemble_load(1, lf.Coeff.const(1.0)).to_host(), ov) <= TOL


def test_load_fan_unstructured_and_large(ctx, lf):
    from scipy.spatial import Delaunay
    pts = np.random.default_rng(7).random((4000, 2))
    tri = Delaunay(pts).simplices
    cn = np.full((tri.shape[0], 4), 0xFFFFFFFF, dtype=np.uint32)
    cn[:, :3] = tri
    gm = ctx.mesh_upload(pts, cn)
    om = lfo.Mesh.from_arrays(pts, cn)
    dm = gm.dofmap_lagrange(1)
    ov, _ = om.assemble_load(1, lfo.coeff.const(3.0))
    assert rel_max_err(dm.assemble_load(1, lf.Coeff.const(3.0)).to_host(), ov) <= TOL
    # 4.5e6 triangles: more rows than one wave of CTAs (prefetch branch), compact ring plan; sum = f * |Omega|
    big = ctx.mesh_tp_tria(1500, 1500)
    dmb = big.dofmap_lagrange(1)
    v = dmb.assemble_load(1, lf.Coeff.const(1.0)).to_host()
    g = dmb.assemble_load(1, lf.Coeff.const(1.0), algo=lf.ALGO_GATHER).to_host()
    assert rel_max_err(v, g) <= TOL and abs(v.sum() - 1.0) <= 1e-11


def test_coordinate_update_rechecks_the_geometry(ctx, lf):
    # lfgpu_mesh_update_node_coords: the reference asserts on a degenerate cell (tria_o1.cc:10-48); the check of the new positions
    # runs asynchronously and surfaces at the next synchronize
    gm = ctx.mesh_tp_tria(4, 4)
    xy = gm.download()["node_coords"].copy()
    gm.update_node_coords(xy)
    ctx.synchronize()  # fine
    bad = xy.copy()
    bad[:] = bad[0]  # every node in one point
    gm.update_node_coords(bad)
    with pytest.raises(lf.LfgpuError) as e:
        ctx.synchronize()
    assert e.value.code == -5
    gm.update_node_coords(xy)
    ctx.synchronize()


# ---- the same rings with a source tabulated per cell / per quadrature point (k_load_p1_fan_src) -------------------------------------
def tabulated_sources(ctx, lf, gm, n_cells):
    """(name, oracle coefficient, gpu coefficient): per cell and per quadrature point (table stride 4 = the larger of the two default
    rules of degree 2, as the C ABI asks for)"""
    xy = gm.qp_coords(1, 4).to_host().reshape(n_cells, 4, 2)
    per_qp = np.ascontiguousarray(1.0 + np.sin(3.0 * xy[..., 0]) * xy[..., 1] + xy[..., 0] ** 2)
    per_cell = np.ascontiguousarray(per_qp[:, :3].mean(axis=1))
    return [("per_cell", lfo.coeff.table(per_cell), lf.Coeff.per_cell(ctx.to_device(per_cell))),
            ("per_qp", lfo.coeff.table(per_qp), lf.Coeff.per_qp(ctx.to_device(per_qp), 4))]


@pytest.mark.parametrize("nx,ny", [(1, 1), (2, 3), (37, 29)])
def test_load_fan_tabulated_source_structured(ctx, lf, nx, ny):
    om = lfo.Mesh.tp_tria(nx, ny, 0.25, -0.5, 1.75, 0.5)
    gm = ctx.mesh_tp_tria(nx, ny, 0.25, -0.5, 1.75, 0.5)
    dm = gm.dofmap_lagrange(1)
    for name, oc, gc in tabulated_sources(ctx, lf, gm, om.n_cells):
        ov, _ = om.assemble_load(1, oc)
        n0 = ctx.kernel_launches
        v = dm.assemble_load(1, gc)
        assert rel_max_err(v.to_host(), ov) <= TOL, name
        g = dm.assemble_load(1, gc, algo=lf.ALGO_GATHER).to_host()
        assert rel_max_err(v.to_host(), g) <= 1e-14, name
        assert np.array_equal(v.to_host(), dm.assemble_load(1, gc).to_host()), name  # fixed order of additions
        dm.assemble_load(1, gc, beta=1.0, out=v)
        assert rel_max_err(v.to_host(), 2 * ov) <= TOL, name
        assert ctx.kernel_launches > n0


def test_load_fan_tabulated_source_other_rule_and_meshes(ctx, lf, golden_meshes):
    # a rule with four points (degree 3) stays on the rings; six points (degree 4) go to the two-pass kernels -- same numbers
    om = lfo.Mesh.tp_tria(23, 17)
    gm = ctx.mesh_tp_tria(23, 17)
    dm = gm.dofmap_lagrange(1)
    for deg in (3, 4):
        q, qq = lf.QuadRule(*lfo.quad_rule(3, deg)), lf.QuadRule(*lfo.quad_rule(4, deg))
        stride = max(q.weights.size, qq.weights.size)
        xy = gm.qp_coords(1, stride, q, qq).to_host().reshape(om.n_cells, stride, 2)
        tab = np.ascontiguousarray(np.cos(xy[..., 0]) + xy[..., 1])
        ov, _ = om.assemble_load(1, lfo.coeff.table(tab), qr_tria=deg, qr_quad=deg)
        v = dm.assemble_load(1, lf.Coeff.per_qp(ctx.to_device(tab), stride), qr_tria=q, qr_quad=qq).to_host()
        assert rel_max_err(v, ov) <= TOL, deg
    # irregular valences, boundary fans, unstructured numbering
    om = lfo.Mesh.from_golden(golden_meshes["6"])
    if not om.n_quad:
        gm = upload_oracle_mesh(ctx, om)[0]
        gm.build_topology(om.export()["edge_nodes"])
        dm = gm.dofmap_lagrange(1)
        for name, oc, gc in tabulated_sources(ctx, lf, gm, om.n_cells):
            ov, _ = om.assemble_load(1, oc)
            assert rel_max_err(dm.assemble_load(1, gc).to_host(), ov) <= TOL, name
    from scipy.spatial import Delaunay
    pts = np.random.default_rng(8).random((3000, 2))
    tri = Delaunay(pts).simplices
    cn = np.full((tri.shape[0], 4), 0xFFFFFFFF, dtype=np.uint32)
    cn[:, :3] = tri
    gm = ctx.mesh_upload(pts, cn)
    om = lfo.Mesh.from_arrays(pts, cn)
    dm = gm.dofmap_lagrange(1)
    for name, oc, gc in tabulated_sources(ctx, lf, gm, om.n_cells):
        ov, _ = om.assemble_load(1, oc)
        assert rel_max_err(dm.assemble_load(1, gc).to_host(), ov) <= TOL, name


def test_load_fan_tabulated_source_large(ctx, lf):
    # 4.5e6 triangles against the two kernels that do not use the rings
    big = ctx.mesh_tp_tria(1500, 1500)
    dm = big.dofmap_lagrange(1)
    xy = big.qp_coords(1, 4).to_host().reshape(big.n_cells, 4, 2)
    tab = np.ascontiguousarray(1.0 + xy[..., 0] * xy[..., 1])
    gc = lf.Coeff.per_qp(ctx.to_device(tab), 4)
    v = dm.assemble_load(1, gc).to_host()
    g = dm.assemble_load(1, gc, algo=lf.ALGO_GATHER).to_host()
    a = dm.assemble_load(1, gc, algo=lf.ALGO_ATOMIC).to_host()
    assert rel_max_err(v, g) <= 1e-14 and rel_max_err(v, a) <= 1e-13
    assert abs(v.sum() - 1.25) <= 1e-9  # integral of 1 + x y over the unit square (midpoint rule: exact for quadratics)
