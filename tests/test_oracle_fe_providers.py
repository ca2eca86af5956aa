"""lf::fe::DiffusionElementMatrixProvider / MassElementMatrixProvider (lib/lf/fe/loc_comp_ellbvp.h:76-226, 256-384) restated in
the oracle (oracle/lfo_uscalfe.h, namespace lfo::fe) for the Lagrange spaces: they are the uscalfe reaction-diffusion provider
with the other coefficient zero -- bitwise in the reference's arithmetic -- which is how the shim maps them onto the device
path (include/lf_gpu_shim.hpp, namespace lfgpu::fe)."""
import numpy as np
import pytest

from oracle import lfo


@pytest.mark.parametrize("degree", [1, 2, 3])
@pytest.mark.parametrize("sel", ["0", "1", "6"])
def test_fe_providers_equal_reaction_diffusion_parts(golden_meshes, degree, sel):
    m = lfo.Mesh.from_golden(golden_meshes[sel])
    alpha, gamma = lfo.coeff.builtin(1), lfo.coeff.builtin(2)  # 1 + x^2 + y^2, 1 / (1 + x^2 + y^2)
    D = m.fe_element_matrices(degree, "diffusion", alpha)
    M = m.fe_element_matrices(degree, "mass", gamma)
    assert np.array_equal(D, m.element_matrices(degree, alpha, lfo.coeff.const(0.0)))
    assert np.array_equal(M, m.element_matrices(degree, lfo.coeff.const(0.0), gamma))
    RD = m.element_matrices(degree, alpha, gamma)
    assert np.abs(D + M - RD).max() <= 1e-14 * np.abs(RD).max()


def test_fe_diffusion_with_tensor_coefficient(golden_meshes):
    m = lfo.Mesh.from_golden(golden_meshes["0"])
    A = [[2.0, 0.5], [-0.25, 1.5]]
    D = m.fe_element_matrices(2, "diffusion", lfo.coeff.const2x2(A))
    assert np.array_equal(D, m.element_matrices(2, lfo.coeff.const2x2(A), lfo.coeff.const(0.0)))
    assert np.abs(D - D.transpose(0, 2, 1)).max() > 1e-3  # a non-symmetric tensor gives non-symmetric element matrices
