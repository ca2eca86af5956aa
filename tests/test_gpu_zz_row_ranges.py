"""Row ranges of the P2 / P3 row kernels (the share of one GPU in the row-block partition, lehrfempp_b200/distributed.py mode
"owner_rows"): assembling [0, N) in uneven pieces gives bitwise the values of one full pass.

Written after the round's GPU minutes were spent: the file name sorts last so that its first run on a B200 (the driver's
round-end run) cannot hide the results of the other files.
"""
import numpy as np
import pytest

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


@pytest.mark.parametrize("degree", [2, 3])
@pytest.mark.parametrize("gamma", [0.0, 0.5])
def test_ranges_tile_the_full_pass(ctx, lf, degree, gamma):
    gm = ctx.mesh_tp_tria(41, 29, 0.0, 0.0, 2.0, 1.0)
    pat = gm.dofmap_lagrange(degree).symbolic(major=lf.ROW_MAJOR)
    outer, _ = pat.download()
    n = outer.size - 1
    a, g = lf.Coeff.const(1.5), lf.Coeff.const(gamma)
    full = pat.assemble_reaction_diffusion(degree, a, g, algo=lf.ALGO_FAN).to_host()
    # cuts inside the vertex rows, inside the edge rows, (P3) inside the cell rows, and one-row pieces
    cuts = sorted({0, 1, 17, gm.n_nodes // 2 + 3, gm.n_nodes, gm.n_nodes + 5, (gm.n_nodes + n) // 2 + 1, n - gm.n_cells // 3, n - 1, n})
    out = ctx.to_device(np.full(pat.nnz, np.nan))
    for r0, r1 in zip(cuts[:-1], cuts[1:]):
        pat.assemble_reaction_diffusion_range(degree, a, g, r0, r1 - r0, out=out, algo=lf.ALGO_FAN)
        part = out.to_host()
        assert not np.isnan(part[outer[r0]:outer[r1]]).any()
    assert np.array_equal(out.to_host(), full)


def test_range_leaves_other_rows_alone(ctx, lf):
    gm = ctx.mesh_tp_tria(20, 20)
    pat = gm.dofmap_lagrange(2).symbolic(major=lf.ROW_MAJOR)
    outer, _ = pat.download()
    n = outer.size - 1
    r0, r1 = n // 3, 2 * n // 3
    out = ctx.to_device(np.full(pat.nnz, 7.0))
    pat.assemble_reaction_diffusion_range(2, lf.Coeff.const(1.0), lf.Coeff.const(0.0), r0, r1 - r0, out=out)
    v = out.to_host()
    assert np.all(v[:outer[r0]] == 7.0) and np.all(v[outer[r1]:] == 7.0)
    full = pat.assemble_reaction_diffusion(2, lf.Coeff.const(1.0), lf.Coeff.const(0.0)).to_host()
    assert np.array_equal(v[outer[r0]:outer[r1]], full[outer[r0]:outer[r1]])
