"""LFGPU_COEFF_NODAL: a coefficient given by its values at the mesh nodes, interpolated at the quadrature
points with the cell's vertex shape functions -- lf::fe::MeshFunctionFE of a FeSpaceLagrangeO1 function.  For an affine function
the interpolation is exact on triangles and on (bilinear) quadrilaterals, so the oracle evaluating the function itself at the
quadrature points is the checker."""
import numpy as np
import pytest

from oracle import lfo
from tests.helpers import rel_max_err
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


@pytest.mark.parametrize("kind", ["tp_tria:7", "hybrid:8", "tp_quad:5", "golden0"])
@pytest.mark.parametrize("degree", [1, 2, 3])
def test_nodal_coefficients_matrix_and_load(ctx, lf, golden_meshes, kind, degree):
    om = oracle_mesh(kind, golden_meshes)
    gm = gpu_mesh(ctx, kind, golden_meshes, om)
    xy = gm.download()["node_coords"]
    fa = lambda x, y: 1.0 + 2.0 * x + 3.0 * y      # noqa: E731
    fg = lambda x, y: 2.0 - 0.5 * x + 0.25 * y     # noqa: E731
    na, ng = ctx.to_device(fa(xy[:, 0], xy[:, 1])), ctx.to_device(fg(xy[:, 0], xy[:, 1]))
    dm = gm.dofmap_lagrange(degree)
    for major in (lf.ROW_MAJOR, lf.COL_MAJOR):
        pat = dm.symbolic(major=major)
        o = om.assemble_rd(degree, lfo.coeff.callback(fa), lfo.coeff.callback(fg), csr=(major == lf.ROW_MAJOR))
        for algo in (lf.ALGO_AUTO, lf.ALGO_GATHER, lf.ALGO_ATOMIC):
            v = pat.assemble_reaction_diffusion(degree, lf.Coeff.nodal(na), lf.Coeff.nodal(ng), algo=algo).to_host()
            assert rel_max_err(v, o[2]) <= TOL
        # one nodal, one constant
        o2 = om.assemble_rd(degree, lfo.coeff.const(1.5), lfo.coeff.callback(fg), csr=(major == lf.ROW_MAJOR))
        v2 = pat.assemble_reaction_diffusion(degree, lf.Coeff.const(1.5), lf.Coeff.nodal(ng)).to_host()
        assert rel_max_err(v2, o2[2]) <= TOL
    ov, _ = om.assemble_load(degree, lfo.coeff.callback(fa))
    for algo in (lf.ALGO_AUTO, lf.ALGO_GATHER):
        assert rel_max_err(dm.assemble_load(degree, lf.Coeff.nodal(na), algo=algo).to_host(), ov) <= TOL
