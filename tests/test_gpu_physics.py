"""Known-answer physics on the CUDA path that does not go through the oracle: viscous decay of a shear wave.

u = (A sin(pi y), 0[, 0]), rho and p constant, is a solution of the compressible Navier-Stokes equations up to O(A^2) (viscous
heating); its amplitude decays like exp(-nu k^2 t), nu = mu / rho, k = pi.  This pins sign and magnitude of the BR1 / BR2 viscous
terms (ViscousFlux.cpp:59-153, SpatialDiscrete.cpp:844-968) against an analytic number instead of the CPU restatement."""
import numpy as np
import pytest

from subrosadg_b200 import mesh as M
from subrosadg_b200.solver import Solver

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dim,cells,visc,tol", [(2, 8, 2, 2e-4), (2, 8, 1, 2e-4), (3, 6, 2, 1e-3)])
def test_shear_wave_decay(built, dim, cells, visc, tol):
    mu, rho, A, k = 0.02, 1.4, 1e-4, np.pi
    mesh = M.periodic_box_fast(dim, cells) if dim == 3 else M.periodic_box(2, cells)
    S = Solver(dict(p=3, model=1, transport=1, mu=mu, conv_flux=2, visc_flux=visc, rk=2), mesh, device=0)

    def ic(x):
        one = np.ones(x.shape[:-1])
        cols = [rho * one, A * np.sin(k * x[..., 1])] + [0 * one] * (dim - 1) + [one]
        return np.stack(cols, axis=-1)

    S.initializeSolver(ic)
    t = S.types[0]
    xq = S.quadrature_coordinates(t)
    s = np.sin(k * xq[..., 1])

    def amplitude():
        q = S.state_at_quadrature(t)
        return float(np.sum(q[..., 1] / q[..., 0] * s) / np.sum(s * s))

    a0 = amplitude()
    assert abs(a0 / A - 1.0) < 1e-6
    dt = 0.2 * S.calculateDeltaTime(1.0)
    nsteps = int(round(1.0 / dt))
    err = S.stepSolver(dt, nsteps)
    assert np.all(np.isfinite(err))
    got = amplitude() / a0
    want = np.exp(-(mu / rho) * k * k * dt * nsteps)
    assert abs(got / want - 1.0) < tol, (got, want)   # discretisation error of P3 on this grid (3-D: 6 cells per wavelength: 6e-4 at 4 cells)
    # the wave really decayed by a measurable amount (the check is not vacuous)
    assert want < 0.9
