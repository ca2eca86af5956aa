"""The known answers of lib/lf/uscalfe/test/full_gal_tests.cc:49-236 recomputed from the GPU matrix: energy v^T A v of the P1
interpolant on GenerateHybrid2DTestMesh(0, 1/3) after six regular refinements ON THE DEVICE, coefficients evaluated at the
quadrature points (PER_QP tables = MeshFunctionGlobal lambdas).  Checked against the reference's exact energies with the
reference's tolerances and against the oracle's matrix (tests/test_oracle_energies.py pins the oracle on the same cases).

Added after the round's GPU minutes were spent (uses only calls other GPU tests already exercise); sorts last on purpose.
"""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import lfo
from tests.helpers import rel_max_err, upload_oracle_mesh
from tests.test_oracle_energies import CASES, REFLEV

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


@pytest.fixture(scope="module")
def finest(ctx, lf, golden_meshes):
    om = lfo.Mesh.from_golden(golden_meshes["0"], 1.0 / 3.0)
    gm = upload_oracle_mesh(ctx, om)[0]
    gm.build_topology(cell_has_geometry=[c["coords"] is not None for c in golden_meshes["0"]["cells"]])
    for _ in range(REFLEV):
        om = om.refine_regular()
        gm = gm.refine_regular()
    assert gm.n_cells == om.n_cells == 9 * 4 ** REFLEV
    dm = gm.dofmap_lagrange(1)
    pat = dm.symbolic(major=lf.COL_MAJOR)
    stride = 4
    qp = gm.qp_coords(1, stride).to_host().reshape(gm.n_cells, stride, 2)
    nodes = gm.download()["node_coords"]
    return om, gm, pat, qp, nodes, stride


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_energy_from_gpu_matrix(ctx, lf, finest, case):
    _, v, alpha, gamma, expected, tol = case
    om, gm, pat, qp, nodes, stride = finest
    a_tab = np.ascontiguousarray(alpha(qp[..., 0], qp[..., 1]))
    g_tab = np.ascontiguousarray(gamma(qp[..., 0], qp[..., 1]))
    vals = pat.assemble_reaction_diffusion(1, lf.Coeff.per_qp(ctx.to_device(a_tab), stride), lf.Coeff.per_qp(ctx.to_device(g_tab), stride)).to_host()
    outer, inner = pat.download()
    n = outer.size - 1
    A = sp.csc_matrix((vals, inner, outer), shape=(n, n))
    vv = v(nodes[:, 0], nodes[:, 1])
    assert abs(vv @ (A @ vv) - expected) <= tol
    # and the oracle's matrix on its own refined mesh (same numbering: bit-exact pattern, values to 1e-12)
    oqp = om.qp_coords(2, 2)
    o_outer, o_inner, o_vals, _, _ = om.assemble_rd(1, lfo.coeff.table(alpha(oqp[..., 0], oqp[..., 1])), lfo.coeff.table(gamma(oqp[..., 0], oqp[..., 1])))
    assert np.array_equal(outer, o_outer) and np.array_equal(inner, o_inner)
    assert rel_max_err(vals, o_vals) <= 1e-12
