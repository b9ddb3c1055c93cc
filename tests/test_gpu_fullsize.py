"""Size-independent properties of the CUDA path at (or near) BASELINE.json's full sizes, where the CPU oracle is too slow to be the
checker: discrete conservation of the residual on periodic meshes (telescoping face terms, SpatialDiscrete.cpp:738-744), the exact
travelling density wave of examples/periodic_3d_ceuler.cpp, set/get round trips, and the chunked face lists at scale."""
import numpy as np
import pytest

import cases
from subrosadg_b200 import mesh as M
from subrosadg_b200.solver import Solver

pytestmark = pytest.mark.gpu

HEX = M.HEXAHEDRON


def quad_weights(S, n_cells, length=2.0):
    """detJ w of the uniform periodic cube: (h/2)^3 times the tensor Gauss weights (Geometry.cpp:44-67)"""
    w1 = S.debug_plan(8)
    h = length / n_cells
    return (h / 2.0) ** 3 * np.einsum("i,j,k->ijk", w1, w1, w1).ravel()


@pytest.mark.parametrize("cfg,cells", [
    (dict(p=3, conv_flux=2, rk=2), 96),                                                         # config 4 family (884,736 elements)
    (dict(p=3, model=1, transport=1, mu=1.4 * 0.2 / 200.0, conv_flux=2, visc_flux=2, rk=2), 64),  # north_star NS target family
])
def test_residual_is_conservative_at_scale(built, cfg, cells):
    mesh = M.periodic_box_fast(3, cells)
    S = Solver(cfg, mesh, device=0)
    S.initializeSolver(cases.ic_density_wave([0.5, 0.3, 0.2]))
    _, q = S.residual()[HEX]                 # dU/dt at the quadrature points [n][Nq][Nv]
    jw = quad_weights(S, cells)
    total = np.einsum("q,eqv->v", jw, q)
    scale = np.einsum("q,eqv->v", jw, np.abs(q)) + 1e-300
    assert np.all(np.isfinite(q))
    assert np.abs(total / scale).max() < 1e-11, (total, scale)   # sums of ~6e7 terms: round-off grows like sqrt(n) eps


def test_travelling_wave_128cube(built):
    """BASELINE configs[3] at full size (128^3 P3 hexahedra, 671 M scalar DOF): the density wave is advected unchanged; P3 on
    h = 1/64 leaves a discretisation error far below 1e-6, and the solution stays on the exact one over 20 steps."""
    cells = 128
    vel = np.array([0.5, 0.3, 0.2])
    mesh = M.periodic_box_fast(3, cells)
    S = Solver(dict(p=3, conv_flux=2, rk=2), mesh, device=0)
    S.initializeSolver(cases.ic_density_wave(vel))
    dt = S.calculateDeltaTime(1.0)
    nsteps = 20
    err = S.stepSolver(dt, nsteps)
    assert np.all(np.isfinite(err)) and np.all(err > 0)
    x = S.quadrature_coordinates(HEX)
    rho = S.state_at_quadrature(HEX)[..., 0]
    exact = 1.0 + 0.2 * np.sin(np.pi * (x.sum(axis=-1) - vel.sum() * dt * nsteps))
    e = float(np.sqrt(np.mean((rho - exact) ** 2)))
    assert e < 1e-7, e


def test_modal_roundtrip_at_scale(built):
    cells = 64
    mesh = M.periodic_box_fast(3, cells)
    S = Solver(dict(p=3), mesh, device=0)
    S.initializeSolver(cases.ic_density_wave([0.5, 0.3, 0.2]))
    U = S.get_state(HEX)
    S.set_state(HEX, U)
    V = S.get_state(HEX)
    assert cases.rel_l2(V, U) < 1e-12   # modal -> collocation -> modal through Phi / Phi^-1 (cond ~ 1e2)
    # every face of the mesh appears in the chunk lists of both of its parents: 3 faces per element, listed twice unless both
    # parents share a 2x2x2 brick (12 of a brick's 36 faces)
    off = S.debug_plan(11)
    assert off[-1] == (cells ** 3 // 8) * 36
