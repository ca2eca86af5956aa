"""Shared helpers of the parity tests (test infrastructure; may use the oracle)."""
import numpy as np

from oracle import lfo

BUILTIN = {
    1: lambda x, y: 1.0 + x * x + y * y,
    2: lambda x, y: 1.0 / (1.0 + x * x + y * y),
    3: lambda x, y: np.sin(2 * np.pi * x) * np.sin(2 * np.pi * y),
    4: lambda x, y: x * y,
    5: lambda x, y: x,
    6: lambda x, y: y,
    7: lambda x, y: x * x - y * y,
    8: lambda x, y: x * x + y * y,
}


def tensor100(x, y):
    """[1 x; y xy] (lagr_fe_tests.cc:818-820), row-major last axis of 4."""
    return np.stack([np.ones_like(x), x, y, x * y], axis=-1)


def rel_max_err(a, b):
    scale = np.abs(b).max()
    return np.abs(a - b).max() / (scale if scale > 0 else 1.0)


def upload_oracle_mesh(ctx, om):
    """Flatten an oracle mesh into the C-ABI upload call (the shim's flatten step, done by the test)."""
    ex = om.export()
    same = True
    nv = np.where(ex["cell_type"] == 3, 3, 4)
    for k in range(4):
        sel = nv > k
        idx = ex["cell_nodes"][sel, k]
        if not np.array_equal(ex["node_coords"][idx], ex["cell_coords"][sel, k]):
            same = False
    return ctx.mesh_upload(ex["node_coords"], ex["cell_nodes"], None if same else ex["cell_coords"]), ex


def per_qp_scalar(ctx, gmesh, degree, fid, qr_tria=None, qr_quad=None, stride=None):
    """Evaluate builtin scalar function `fid` at every quadrature point (host numpy) -> PER_QP device table."""
    import lehrfempp_b200 as lf
    if stride is None:
        nt = (qr_tria.weights.size if qr_tria else lf.default_quad_rule(3, 2 * degree).weights.size)
        nq = (qr_quad.weights.size if qr_quad else lf.default_quad_rule(4, 2 * degree).weights.size)
        stride = max(nt, nq)
    xy = gmesh.qp_coords(degree, stride, qr_tria, qr_quad).to_host().reshape(gmesh.n_cells, stride, 2)
    vals = BUILTIN[fid](xy[..., 0], xy[..., 1])
    dev = ctx.to_device(np.ascontiguousarray(vals))
    return lf.Coeff.per_qp(dev, stride), stride


def per_qp_tensor100(ctx, gmesh, degree, qr_tria=None, qr_quad=None):
    import lehrfempp_b200 as lf
    nt = (qr_tria.weights.size if qr_tria else lf.default_quad_rule(3, 2 * degree).weights.size)
    nq = (qr_quad.weights.size if qr_quad else lf.default_quad_rule(4, 2 * degree).weights.size)
    stride = max(nt, nq)
    xy = gmesh.qp_coords(degree, stride, qr_tria, qr_quad).to_host().reshape(gmesh.n_cells, stride, 2)
    vals = tensor100(xy[..., 0], xy[..., 1])
    dev = ctx.to_device(np.ascontiguousarray(vals))
    return lf.Coeff.per_qp_2x2(dev, stride)
