"""GPU parity tests: the CUDA path, called through the C ABI, against the oracle on the same inputs.

Bars (BASELINE.json north_star): mesh numbering, dof tables and sparsity pattern BIT-EXACT; matrix / vector values
within 1e-12 relative in max-norm.
"""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import lfo
from tests.helpers import BUILTIN, per_qp_scalar, per_qp_tensor100, rel_max_err, upload_oracle_mesh

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


def oracle_mesh(kind, golden_meshes):
    if kind.startswith("golden"):
        return lfo.Mesh.from_golden(golden_meshes[kind[6:]])
    name, n = kind.split(":")
    n = int(n)
    if name == "tp_tria":
        return lfo.Mesh.tp_tria(n, n + 1, 0.25, -0.5, 1.75, 0.5)
    if name == "tp_quad":
        return lfo.Mesh.tp_quad(n + 2, n, -1.0, 0.0, 1.0, 3.0)
    return lfo.Mesh.hybrid(n, 0.2, 12345)


def gpu_mesh(ctx, kind, golden_meshes, om):
    if kind.startswith("golden"):
        gm = upload_oracle_mesh(ctx, om)[0]
        entry = golden_meshes[kind[6:]]
        if "cells" in entry:
            # same optional-geometry policy as the MeshFactory calls of the fixture (test_meshes.cc)
            gm.build_topology(cell_has_geometry=[c["coords"] is not None for c in entry["cells"]])
        else:
            gm.build_topology(om.export()["edge_nodes"])  # selector 4 is the triangle builder: explicit edges
        return gm
    name, n = kind.split(":")
    n = int(n)
    if name == "tp_tria":
        return ctx.mesh_tp_tria(n, n + 1, 0.25, -0.5, 1.75, 0.5)
    if name == "tp_quad":
        return ctx.mesh_tp_quad(n + 2, n, -1.0, 0.0, 1.0, 3.0)
    return ctx.mesh_hybrid(n, 0.2, 12345)


MESHES = ["tp_tria:7", "tp_quad:5", "hybrid:9", "hybrid:10", "golden0", "golden1", "golden3", "golden4", "golden5", "golden6", "golden8"]


# ---- mesh numbering ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", MESHES)
def test_mesh_and_topology_bit_exact(ctx, golden_meshes, kind):
    om = oracle_mesh(kind, golden_meshes)
    gm = gpu_mesh(ctx, kind, golden_meshes, om)
    ex = om.export()
    assert (gm.n_nodes, gm.n_cells, gm.n_tria, gm.n_quad) == (om.n_nodes, om.n_cells, om.n_tria, om.n_quad)
    d = gm.download(topology=True)
    assert gm.n_edges == om.n_edges
    assert np.array_equal(d["cell_nodes"], ex["cell_nodes"])
    assert np.array_equal(d["cell_type"], ex["cell_type"])
    assert np.array_equal(d["node_coords"].view(np.uint64), ex["node_coords"].view(np.uint64))  # bitwise
    assert np.array_equal(d["cell_coords"].view(np.uint64), ex["cell_coords"].view(np.uint64))
    assert np.array_equal(d["edge_nodes"], ex["edge_nodes"])
    assert np.array_equal(d["cell_edges"], ex["cell_edges"])
    assert np.array_equal(d["cell_edge_ori"], ex["cell_edge_ori"])


def test_explicit_edges_keep_index_and_direction(ctx):
    # hybrid2d/mesh.cc:240-274: supplied edges keep their position as index and their orientation
    om = lfo.Mesh.tp_tria(3, 2)
    ex = om.export()
    gm = ctx.mesh_upload(ex["node_coords"], ex["cell_nodes"])
    gm.build_topology(ex["edge_nodes"][::-1].copy()[:, ::-1].copy())  # reversed list, flipped directions
    d = gm.download(topology=True)
    om2 = lfo.Mesh.from_arrays(ex["node_coords"], ex["cell_nodes"], edge_nodes=ex["edge_nodes"][::-1].copy()[:, ::-1].copy())
    e2 = om2.export()
    assert np.array_equal(d["edge_nodes"], e2["edge_nodes"])
    assert np.array_equal(d["cell_edges"], e2["cell_edges"])
    assert np.array_equal(d["cell_edge_ori"], e2["cell_edge_ori"])


def test_partial_explicit_edges(ctx, golden_meshes):
    om0 = lfo.Mesh.from_golden(golden_meshes["0"])
    ex = om0.export()
    some = ex["edge_nodes"][[5, 2, 11]].copy()
    om = lfo.Mesh.from_arrays(ex["node_coords"], ex["cell_nodes"], edge_nodes=some)
    e2 = om.export()
    gm = ctx.mesh_upload(ex["node_coords"], ex["cell_nodes"])
    gm.build_topology(some)
    d = gm.download(topology=True)
    assert np.array_equal(d["edge_nodes"], e2["edge_nodes"])
    assert np.array_equal(d["cell_edges"], e2["cell_edges"])
    assert np.array_equal(d["cell_edge_ori"], e2["cell_edge_ori"])


def test_degenerate_cell_rejected(ctx, lf):
    xy = np.array([[0.0, 0.0], [1.0, 0.0], [2.0, 0.0]])
    cn = np.array([[0, 1, 2, lf.api.NIL]], dtype=np.uint32)
    with pytest.raises(lf.LfgpuError) as e:
        ctx.mesh_upload(xy, cn)
    assert e.value.code == -5


# ---- dof handler ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", MESHES)
@pytest.mark.parametrize("degree", [1, 2, 3])
def test_lagrange_dofs_bit_exact(ctx, golden_meshes, kind, degree):
    om = oracle_mesh(kind, golden_meshes)
    gm = gpu_mesh(ctx, kind, golden_meshes, om)
    dm = gm.dofmap_lagrange(degree)
    od, onl = om.cell_dofs(degree)
    gd, gnl = dm.download()
    assert dm.num_dofs == om.num_dofs(degree)
    assert np.array_equal(gnl, onl)
    assert np.array_equal(gd, od)


def test_uniform_two_dofs_per_edge(ctx, golden_meshes):
    # the layout of the 36x36 golden (assembly_tests.cc:460-470)
    om = lfo.Mesh.from_golden(golden_meshes["0"])
    gm = upload_oracle_mesh(ctx, om)[0]
    dm = gm.dofmap_uniform(n_seg=2)
    od, onl = lfo.DofHandler(om, n_seg=2).cell_dofs()
    gd, gnl = dm.download()
    assert dm.num_dofs == 36
    assert np.array_equal(gd, od) and np.array_equal(gnl, onl)


def test_uploaded_dofmap_roundtrip(ctx, golden_meshes):
    om = lfo.Mesh.from_golden(golden_meshes["0"])
    gm = upload_oracle_mesh(ctx, om)[0]
    od, onl = om.cell_dofs(3)
    dm = gm.dofmap_upload(om.num_dofs(3), od, onl)
    gd, gnl = dm.download()
    assert np.array_equal(gd, od) and np.array_equal(gnl, onl)


# ---- pattern + values ------------------------------------------------------------------------------------------------
def assemble_both(ctx, lf, om, gm, degree, oalpha, ogamma, galpha, ggamma, major, algo, qr=None, active=None, repeat=1):
    q = -1 if qr is None else qr
    o_outer, o_inner, o_vals, shape, _ = om.assemble_rd(degree, oalpha, ogamma, qr_tria=q, qr_quad=q, csr=(major == lf.ROW_MAJOR),
                                                        active=active, repeat=repeat)
    dm = gm.dofmap_lagrange(degree)
    pat = dm.symbolic(major=major)
    qt = qq = None
    if qr is not None:
        qt = lf.QuadRule(*lfo.quad_rule(3, qr))
        qq = lf.QuadRule(*lfo.quad_rule(4, qr))
    dact = ctx.to_device(np.asarray(active, dtype=np.uint8)) if active is not None else None
    vals = pat.assemble_reaction_diffusion(degree, galpha, ggamma, qt, qq, active=dact, algo=algo)
    for _ in range(repeat - 1):
        pat.assemble_reaction_diffusion(degree, galpha, ggamma, qt, qq, active=dact, beta=1.0, out=vals, algo=algo)
    g_outer, g_inner = pat.download()
    return (o_outer, o_inner, o_vals), (g_outer, g_inner, vals.to_host()), shape


@pytest.mark.parametrize("kind", MESHES)
@pytest.mark.parametrize("degree", [1, 2, 3])
@pytest.mark.parametrize("major", [0, 1])
def test_pattern_bit_exact_and_values_const(ctx, lf, golden_meshes, kind, degree, major):
    om = oracle_mesh(kind, golden_meshes)
    gm = gpu_mesh(ctx, kind, golden_meshes, om)
    for algo in (lf.ALGO_ATOMIC, lf.ALGO_GATHER):
        o, g, _ = assemble_both(ctx, lf, om, gm, degree, lfo.coeff.const(1.5), lfo.coeff.const(0.75), lf.Coeff.const(1.5),
                                lf.Coeff.const(0.75), major, algo)
        assert np.array_equal(o[0], g[0]), "outer index array differs"
        assert np.array_equal(o[1], g[1]), "inner index array differs"
        assert rel_max_err(g[2], o[2]) <= TOL


@pytest.mark.parametrize("kind", ["tp_tria:6", "hybrid:8", "golden0", "golden1"])
@pytest.mark.parametrize("degree", [1, 2, 3])
@pytest.mark.parametrize("major", [0, 1])
def test_values_nonsymmetric_tensor_and_variable_coefficients(ctx, lf, golden_meshes, kind, degree, major):
    om = oracle_mesh(kind, golden_meshes)
    gm = gpu_mesh(ctx, kind, golden_meshes, om)
    A = [[3.0, 0.0], [1.0, 2.0]]
    for algo in (lf.ALGO_ATOMIC, lf.ALGO_GATHER):
        # constant non-symmetric 2x2 tensor: separates rows from columns
        o, g, _ = assemble_both(ctx, lf, om, gm, degree, lfo.coeff.const2x2(A), lfo.coeff.const(0.0), lf.Coeff.const2x2(A),
                                lf.Coeff.const(0.0), major, algo)
        assert np.array_equal(o[0], g[0]) and np.array_equal(o[1], g[1])
        assert rel_max_err(g[2], o[2]) <= TOL
        # alpha = 1 + |x|^2, gamma = 1/(1+|x|^2) evaluated per quadrature point (config C2's coefficients)
        ga, _ = per_qp_scalar(ctx, gm, degree, 1)
        gg, _ = per_qp_scalar(ctx, gm, degree, 2)
        o, g, _ = assemble_both(ctx, lf, om, gm, degree, lfo.coeff.builtin(1), lfo.coeff.builtin(2), ga, gg, major, algo)
        assert rel_max_err(g[2], o[2]) <= TOL
        # variable non-symmetric tensor [1 x; y xy] (lagr_fe_tests.cc:818-820)
        gt = per_qp_tensor100(ctx, gm, degree)
        gg4, _ = per_qp_scalar(ctx, gm, degree, 4)
        o, g, _ = assemble_both(ctx, lf, om, gm, degree, lfo.coeff.builtin(100), lfo.coeff.builtin(4), gt, gg4, major, algo)
        assert rel_max_err(g[2], o[2]) <= TOL


def test_reference_bilinear_form_known_answers_on_gpu(ctx, lf, golden_meshes):
    # lagr_fe_tests.cc:807-882 re-run with the GPU-assembled matrix: 7911/8, 81, 1996731/280
    om = lfo.Mesh.from_golden(golden_meshes["0"])
    gm = upload_oracle_mesh(ctx, om)[0]
    c = lfo.coeff
    cases = [(1, 4, "tensor", 9, 10, 7911.0 / 8.0), (2, 6, 5, 8, 7, 81.0), (3, 8, 6, 11, 12, 1996731.0 / 280.0)]
    for degree, qr, alpha, fa, fb, expect in cases:
        qt = lf.QuadRule(*lfo.quad_rule(3, qr))
        qq = lf.QuadRule(*lfo.quad_rule(4, qr))
        if alpha == "tensor":
            ga = per_qp_tensor100(ctx, gm, degree, qt, qq)
        else:
            ga, _ = per_qp_scalar(ctx, gm, degree, alpha, qt, qq)
        gg, _ = per_qp_scalar(ctx, gm, degree, 4, qt, qq)
        dm = gm.dofmap_lagrange(degree)
        pat = dm.symbolic(major=lf.COL_MAJOR)
        vals = pat.assemble_reaction_diffusion(degree, ga, gg, qt, qq).to_host()
        outer, inner = pat.download()
        A = sp.csc_matrix((vals, inner, outer), shape=(pat.rows, pat.cols))
        av = om.nodal_projection(degree, c.builtin(fa))
        bv = om.nodal_projection(degree, c.builtin(fb))
        assert abs(av @ (A @ bv) - expect) < 1e-10 * abs(expect)


@pytest.mark.parametrize("degree", [1, 2, 3])
def test_active_mask_and_accumulate(ctx, lf, golden_meshes, degree):
    om = oracle_mesh("hybrid:7", golden_meshes)
    gm = gpu_mesh(ctx, "hybrid:7", golden_meshes, om)
    rng = np.random.default_rng(5)
    active = (rng.random(om.n_cells) < 0.6).astype(np.uint8)
    for algo in (lf.ALGO_ATOMIC, lf.ALGO_GATHER):
        # isActive (assembler.h:127): inactive cells contribute nothing but the pattern is the full one on the GPU;
        # the oracle's COO only holds triplets of active cells, so compare as matrices
        o, g, shape = assemble_both(ctx, lf, om, gm, degree, lfo.coeff.const(1.0), lfo.coeff.const(1.0), lf.Coeff.const(1.0),
                                    lf.Coeff.const(1.0), lf.ROW_MAJOR, algo, active=active)
        Ao = sp.csr_matrix((o[2], o[1], o[0]), shape=shape)
        Ag = sp.csr_matrix((g[2], g[1], g[0]), shape=shape)
        assert abs(Ao - Ag).max() <= TOL * np.abs(o[2]).max()
        # accumulate semantics (assembler.h:84-88): assembling twice doubles the entries
        o, g, _ = assemble_both(ctx, lf, om, gm, degree, lfo.coeff.const(1.0), lfo.coeff.const(1.0), lf.Coeff.const(1.0),
                                lf.Coeff.const(1.0), lf.ROW_MAJOR, algo, repeat=2)
        assert rel_max_err(g[2], o[2]) <= TOL


def test_missing_rule_is_an_error(ctx, lf, golden_meshes):
    # loc_comp_test.cc:154-183: only a triangle rule on a hybrid mesh -> LfException in the reference
    om = lfo.Mesh.from_golden(golden_meshes["0"])
    gm = upload_oracle_mesh(ctx, om)[0]
    pat = gm.dofmap_lagrange(1).symbolic()
    qt = lf.QuadRule(*lfo.quad_rule(3, 2))
    with pytest.raises(lf.LfgpuError) as e:
        pat.assemble_reaction_diffusion(1, lf.Coeff.const(1.0), lf.Coeff.const(0.0), qt, None)
    assert e.value.code == -4


# ---- load vector ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["tp_tria:9", "hybrid:8", "golden0", "golden1", "golden6"])
@pytest.mark.parametrize("degree", [1, 2, 3])
def test_load_vector(ctx, lf, golden_meshes, kind, degree):
    om = oracle_mesh(kind, golden_meshes)
    gm = gpu_mesh(ctx, kind, golden_meshes, om)
    dm = gm.dofmap_lagrange(degree)
    gf, _ = per_qp_scalar(ctx, gm, degree, 3)
    ov, _ = om.assemble_load(degree, lfo.coeff.builtin(3))
    gv = dm.assemble_load(degree, gf).to_host()
    assert rel_max_err(gv, ov) <= TOL
    # constant source, accumulate on top (assembler.h:291-293: the vector is not zeroed)
    out = dm.assemble_load(degree, lf.Coeff.const(2.0))
    dm.assemble_load(degree, lf.Coeff.const(2.0), beta=1.0, out=out)
    ov2, _ = om.assemble_load(degree, lfo.coeff.const(2.0))
    assert rel_max_err(out.to_host(), 2 * ov2) <= TOL


# ---- BASELINE config C1 -----------------------------------------------------------------------------------------------
def test_config_c1_p1_laplacian_256(ctx, lf):
    om = lfo.Mesh.tp_tria(256, 256)
    gm = ctx.mesh_tp_tria(256, 256)
    dm = gm.dofmap_lagrange(1)
    for major in (lf.COL_MAJOR, lf.ROW_MAJOR):
        pat = dm.symbolic(major=major)
        o_outer, o_inner, o_vals, shape, _ = om.assemble_rd(1, lfo.coeff.const(1.0), lfo.coeff.const(0.0), csr=(major == lf.ROW_MAJOR))
        outer, inner = pat.download()
        assert pat.nnz == 460289 and dm.num_dofs == 66049
        assert np.array_equal(outer, o_outer) and np.array_equal(inner, o_inner)
        for algo in (lf.ALGO_ATOMIC, lf.ALGO_GATHER, lf.ALGO_FAN):
            vals = pat.assemble_reaction_diffusion(1, lf.Coeff.const(1.0), lf.Coeff.const(0.0), algo=algo).to_host()
            assert rel_max_err(vals, o_vals) <= TOL
            assert (vals == 0.0).sum() >= 2 * 256 * 256  # explicit zeros on the diagonal edges stay in the pattern
    gf, _ = per_qp_scalar(ctx, gm, 1, 3)
    ov, _ = om.assemble_load(1, lfo.coeff.builtin(3))
    assert rel_max_err(dm.assemble_load(1, gf).to_host(), ov) <= TOL


# ---- size-independent properties at a larger size ----------------------------------------------------------------------
@pytest.mark.parametrize("degree,n", [(1, 1500), (2, 600), (3, 300)])
def test_large_mesh_properties(ctx, lf, degree, n):
    gm = ctx.mesh_tp_tria(n, n)
    dm = gm.dofmap_lagrange(degree)
    pat = dm.symbolic(major=lf.ROW_MAJOR)
    outer, inner = pat.download()
    N = dm.num_dofs
    assert np.all(np.diff(outer) > 0)
    # inner indices strictly ascending inside every row
    d = np.diff(inner.astype(np.int64))
    row_start = np.zeros(inner.size, dtype=bool)
    row_start[outer[1:-1]] = True
    assert np.all(d[~row_start[1:]] > 0)
    stiff = pat.assemble_reaction_diffusion(degree, lf.Coeff.const(1.0), lf.Coeff.const(0.0), algo=lf.ALGO_GATHER).to_host()
    A = sp.csr_matrix((stiff, inner, outer), shape=(N, N))
    scale = np.abs(stiff).max()
    assert np.abs(A @ np.ones(N)).max() <= 1e-11 * scale          # constants are in the kernel (bvp_fe_tests.cc:30-60)
    assert abs(A - A.T).max() <= 1e-12 * scale                    # symmetric form
    mass = pat.assemble_reaction_diffusion(degree, lf.Coeff.const(0.0), lf.Coeff.const(1.0), algo=lf.ALGO_GATHER).to_host()
    assert abs(mass.sum() - 1.0) <= 1e-11                         # sum of the mass matrix = |Omega| (loc_comp_test.cc:46-84)
    atom = pat.assemble_reaction_diffusion(degree, lf.Coeff.const(1.0), lf.Coeff.const(0.0), algo=lf.ALGO_ATOMIC).to_host()
    assert rel_max_err(atom, stiff) <= TOL                        # the two scatter strategies agree
    again = pat.assemble_reaction_diffusion(degree, lf.Coeff.const(1.0), lf.Coeff.const(0.0), algo=lf.ALGO_GATHER).to_host()
    assert np.array_equal(again, stiff)                           # gather path is deterministic (bitwise repeatable)


# ---- P1 vertex-fan fast path (LFGPU_ALGO_FAN) -------------------------------------------------------------------------
def fan_meshes():
    """Triangle meshes that stress the fan plan: structured, unstructured, a non-manifold 'bow-tie' vertex, a vertex of
    valence 14 (longer than the ring) and mixed cell orientations."""
    out = {}
    # bow-tie: two triangles that share only vertex 0, plus a regular strip
    xy = np.array([[0.0, 0.0], [1.0, 0.0], [1.0, 1.0], [-1.0, 0.0], [-1.0, -1.0], [2.0, 0.0], [2.0, 1.0]])
    cn = np.array([[0, 1, 2, NIL], [0, 3, 4, NIL], [1, 5, 2, NIL], [5, 6, 2, NIL]], dtype=np.uint32)
    out["bowtie"] = (xy, cn)
    # wheel of 14 triangles around vertex 0, alternating orientation
    k = 14
    ang = 2 * np.pi * np.arange(k) / k
    xy = np.vstack([[0.0, 0.0], np.stack([np.cos(ang), np.sin(ang)], 1) * (1.0 + 0.1 * np.cos(3 * ang))[:, None]])
    cn = np.array([[0, 1 + t, 1 + (t + 1) % k, NIL] if t % 2 == 0 else [1 + (t + 1) % k, 1 + t, 0, NIL] for t in range(k)],
                  dtype=np.uint32)
    out["wheel14"] = (xy, cn)
    # wheel of 7 (closed fan of odd length) + an open fan
    k = 7
    ang = 2 * np.pi * np.arange(k) / k
    xy = np.vstack([[0.0, 0.0], np.stack([np.cos(ang), np.sin(ang)], 1)])
    cn = np.array([[0, 1 + t, 1 + (t + 1) % k, NIL] for t in range(k)], dtype=np.uint32)
    out["wheel7"] = (xy, cn)
    return out


NIL = 0xFFFFFFFF


@pytest.mark.parametrize("kind", ["tp_tria:11", "golden3", "golden4", "bowtie", "wheel14", "wheel7"])
@pytest.mark.parametrize("major", [0, 1])
def test_p1_fan_kernel(ctx, lf, golden_meshes, kind, major):
    if kind in ("bowtie", "wheel14", "wheel7"):
        xy, cn = fan_meshes()[kind]
        om = lfo.Mesh.from_arrays(xy, cn)
        gm = ctx.mesh_upload(xy, cn)
    else:
        om = oracle_mesh(kind, golden_meshes)
        gm = gpu_mesh(ctx, kind, golden_meshes, om)
    dm = gm.dofmap_lagrange(1)
    pat = dm.symbolic(major=major)
    A = [[3.0, 0.5], [1.0, 2.0]]
    cases = [(lfo.coeff.const(1.0), lfo.coeff.const(0.0), lf.Coeff.const(1.0), lf.Coeff.const(0.0)),
             (lfo.coeff.const(2.5), lfo.coeff.const(0.75), lf.Coeff.const(2.5), lf.Coeff.const(0.75)),
             (lfo.coeff.const2x2(A), lfo.coeff.const(1.25), lf.Coeff.const2x2(A), lf.Coeff.const(1.25))]
    for oa, og, ga, gg in cases:
        o_outer, o_inner, o_vals, _, _ = om.assemble_rd(1, oa, og, csr=(major == lf.ROW_MAJOR))
        outer, inner = pat.download()
        assert np.array_equal(outer, o_outer) and np.array_equal(inner, o_inner)
        v = pat.assemble_reaction_diffusion(1, ga, gg, algo=lf.ALGO_FAN)
        assert rel_max_err(v.to_host(), o_vals) <= TOL
        # AUTO takes the same path; accumulate on top (beta = 1) doubles the entries
        pat.assemble_reaction_diffusion(1, ga, gg, beta=1.0, out=v, algo=lf.ALGO_AUTO)
        assert rel_max_err(v.to_host(), 2 * o_vals) <= TOL
        # explicit user rule of higher degree gives the same matrix (constant coefficients are integrated exactly)
        qt = lf.QuadRule(*lfo.quad_rule(3, 6))
        v6 = pat.assemble_reaction_diffusion(1, ga, gg, qt, None, algo=lf.ALGO_FAN).to_host()
        assert rel_max_err(v6, o_vals) <= TOL


def test_p1_fan_not_applicable(ctx, lf, golden_meshes):
    om = lfo.Mesh.from_golden(golden_meshes["0"])          # hybrid mesh
    gm = upload_oracle_mesh(ctx, om)[0]
    pat = gm.dofmap_lagrange(1).symbolic()
    # round 2: hybrid meshes have a row kernel of their own (assemble_p1h.cu), so LFGPU_ALGO_FAN is served ...
    o = om.assemble_rd(1, lfo.coeff.const(1.0), lfo.coeff.const(0.0), csr=False)
    v = pat.assemble_reaction_diffusion(1, lf.Coeff.const(1.0), lf.Coeff.const(0.0), algo=lf.ALGO_FAN).to_host()
    assert rel_max_err(v, o[2]) <= TOL
    # ... unless the rule has other point counts than the kernels are compiled for: LFGPU_ERR_UNSUPPORTED, no silent fallback
    qt, qq = lf.QuadRule(*lfo.quad_rule(3, 6)), lf.QuadRule(*lfo.quad_rule(4, 6))
    with pytest.raises(lf.LfgpuError) as e:
        pat.assemble_reaction_diffusion(1, lf.Coeff.const(1.0), lf.Coeff.const(0.0), qt, qq, algo=lf.ALGO_FAN)
    assert e.value.code == -7
    # AUTO falls back to the generic kernel
    o = om.assemble_rd(1, lfo.coeff.const(1.0), lfo.coeff.const(0.0), csr=True)
    v = pat.assemble_reaction_diffusion(1, lf.Coeff.const(1.0), lf.Coeff.const(0.0)).to_host()
    assert rel_max_err(v, o[2]) <= TOL


def test_p1_fan_row_list(ctx, lf):
    # the multi-GPU building block: only the listed rows are written
    gm = ctx.mesh_tp_tria(40, 30)
    om = lfo.Mesh.tp_tria(40, 30)
    dm = gm.dofmap_lagrange(1)
    pat = dm.symbolic(major=lf.ROW_MAJOR)
    outer, inner = pat.download()
    o = om.assemble_rd(1, lfo.coeff.const(1.0), lfo.coeff.const(0.0), csr=True)
    rows = np.arange(100, 900, dtype=np.int32)
    v = ctx.to_device(np.full(pat.nnz, -7.0))
    pat.assemble_reaction_diffusion(1, lf.Coeff.const(1.0), lf.Coeff.const(0.0), out=v, rows=ctx.to_device(rows))
    h = v.to_host()
    lo, hi = outer[100], outer[900]
    assert rel_max_err(h[lo:hi], o[2][lo:hi]) <= TOL
    assert np.all(h[:lo] == -7.0) and np.all(h[hi:] == -7.0)


# ---- the kernels that own rows in registers, at sizes where their prefetch branches and compact plans are active ---------------
@pytest.mark.parametrize("degree,n", [(1, 707), (2, 400), (3, 250)])
def test_row_kernels_against_the_oracle_at_scale(ctx, lf, degree, n):
    # P1: 1.0e6 triangles = the smallest size of BASELINE config 5; more rows than one wave of resident CTAs (151 552), so the
    # L2-prefetch branch, the compact 16-bit ring plan and the staged line writes of the fan kernel are all compared with the oracle
    om = lfo.Mesh.tp_tria(n, n)
    gm = ctx.mesh_tp_tria(n, n)
    pat = gm.dofmap_lagrange(degree).symbolic(major=lf.ROW_MAJOR)
    o_outer, o_inner, o_vals, _, _ = om.assemble_rd(degree, lfo.coeff.const(1.0), lfo.coeff.const(0.0), csr=True)
    outer, inner = pat.download()
    assert np.array_equal(outer, o_outer) and np.array_equal(inner, o_inner)
    vals = pat.assemble_reaction_diffusion(degree, lf.Coeff.const(1.0), lf.Coeff.const(0.0), algo=lf.ALGO_FAN).to_host()
    assert rel_max_err(vals, o_vals) <= TOL
    o2 = om.assemble_rd(degree, lfo.coeff.const2x2([[2.0, 0.5], [-0.25, 1.0]]), lfo.coeff.const(0.75), csr=True)
    vals = pat.assemble_reaction_diffusion(degree, lf.Coeff.const2x2([[2.0, 0.5], [-0.25, 1.0]]), lf.Coeff.const(0.75), algo=lf.ALGO_FAN).to_host()
    assert rel_max_err(vals, o2[2]) <= TOL


@pytest.mark.parametrize("degree,n", [(1, 1800), (2, 900), (3, 500)])
def test_row_kernels_against_the_generic_kernel_large(ctx, lf, degree, n):
    gm = ctx.mesh_tp_tria(n, n)
    pat = gm.dofmap_lagrange(degree).symbolic(major=lf.ROW_MAJOR)
    for alpha, gamma in ((lf.Coeff.const(1.0), lf.Coeff.const(0.0)), (lf.Coeff.const2x2([[2.0, 0.5], [-0.25, 1.0]]), lf.Coeff.const(0.75))):
        rows = pat.assemble_reaction_diffusion(degree, alpha, gamma, algo=lf.ALGO_FAN).to_host()
        gen = pat.assemble_reaction_diffusion(degree, alpha, gamma, algo=lf.ALGO_GATHER).to_host()
        assert rel_max_err(rows, gen) <= TOL


@pytest.mark.parametrize("degree", [2, 3])
@pytest.mark.parametrize("kind", ["tp_tria:24", "delaunay"])
def test_row_kernels_accumulate(ctx, lf, golden_meshes, degree, kind):
    # assembler.h:84-88: the matrix is not zeroed -- a second call adds; the P2 / P3 row kernels keep this on their fast path
    # (round 1 fell back to the item kernel for beta != 0)
    if kind == "delaunay":
        from scipy.spatial import Delaunay
        pts = np.random.default_rng(3).random((900, 2))
        tri = Delaunay(pts).simplices
        cn = np.full((tri.shape[0], 4), 0xFFFFFFFF, dtype=np.uint32)
        cn[:, :3] = tri
        om = lfo.Mesh.from_arrays(pts, cn)
        gm = ctx.mesh_upload(pts, cn)
        gm.build_topology()
    else:
        om = oracle_mesh(kind, golden_meshes)
        gm = gpu_mesh(ctx, kind, golden_meshes, om)
    pat = gm.dofmap_lagrange(degree).symbolic(major=lf.ROW_MAJOR)
    o = om.assemble_rd(degree, lfo.coeff.const(1.5), lfo.coeff.const(0.5), csr=True)
    v = pat.assemble_reaction_diffusion(degree, lf.Coeff.const(1.5), lf.Coeff.const(0.5), algo=lf.ALGO_FAN)
    assert rel_max_err(v.to_host(), o[2]) <= TOL
    pat.assemble_reaction_diffusion(degree, lf.Coeff.const(1.5), lf.Coeff.const(0.5), beta=1.0, out=v, algo=lf.ALGO_FAN)
    assert rel_max_err(v.to_host(), 2 * o[2]) <= TOL
    pat.assemble_reaction_diffusion(degree, lf.Coeff.const(1.5), lf.Coeff.const(0.5), beta=-0.5, out=v, algo=lf.ALGO_FAN)
    assert np.abs(v.to_host()).max() <= 1e-12 * np.abs(o[2]).max()
