"""ViewVariable::get (src/Solver/VariableConvertor.cpp:754-872) on the device (sdg_get_view_variable) against the oracle's restatement, at
the volume quadrature points: every ViewVariableEnum entry that exists for the case, plus the fall-through entries of the switch."""
import numpy as np
import pytest

import cases
from subrosadg_b200 import mesh as M

pytestmark = pytest.mark.gpu

NAMES = {0: "Density", 1: "Velocity", 2: "Temperature", 3: "Pressure", 4: "SoundSpeed", 5: "MachNumber", 6: "Entropy", 7: "Vorticity", 8: "HeatFlux",
         9: "ArtificialViscosity", 10: "VelocityX", 11: "VelocityY", 12: "VelocityZ", 13: "MachNumberX", 14: "MachNumberY", 15: "MachNumberZ",
         16: "VorticityX", 17: "VorticityY", 18: "VorticityZ", 19: "HeatFluxX", 20: "HeatFluxY", 21: "HeatFluxZ"}


def check(O, S, variables, label, tol=1e-11):
    for t in S.types:
        S.set_state(t, O.get_state(t))
    O.residual()      # fills the oracle's gradient coefficients / artificial viscosity for the current state
    S.update_artificial_viscosity()
    for v in variables:
        for t in S.types:
            a, b = S.view_variable(t, v), O.view_variable(t, v)
            scale = max(np.abs(b).max(), 1e-300)
            assert np.abs(a - b).max() <= tol * max(scale, 1.0), f"{label}: {NAMES[v]} type {t}: max abs diff {np.abs(a - b).max():.3e} (scale {scale:.3e})"


def test_view_variables_ns_3d(built):
    mesh = M.box(3, (3, 3, 3), 0.0, 1.0, geom_order=2, warp=lambda x: x + 0.03 * np.sin(np.pi * np.roll(x, 1, axis=1)),
                 phys_bc={k: M.RIEMANN_FARFIELD for k in range(1, 7)})
    ic = cases.ic_perturbed_freestream(0.4, 3.0, 3, amp=0.05)
    cfg = dict(p=3, model=1, transport=2, mu=0.01, conv_flux=2, visc_flux=2, rk=2)
    O, S = cases.make_pair(cfg, mesh, ic, cases.bc_freestream(0.4, 3.0, 3, wall_phys=()))
    check(O, S, list(range(0, 8)) + list(range(10, 22)) + [8, 9], "ns 3d")
    assert np.abs(S.view_variable(S.types[0], 7)).max() > 1e-3      # the vorticity magnitude of the perturbed flow is not trivially zero


def test_view_variables_ns_hybrid_2d(built):
    mesh = M.annulus(4, 12, r0=0.5, r1=3.0, geom_order=3, tri_rings=2, phys_bc={1: M.RIEMANN_FARFIELD, 2: M.ADIABATIC_NONSLIP_WALL})
    ic = cases.ic_perturbed_freestream(0.3, 0.0, 2, amp=0.02)
    cfg = dict(p=3, model=1, transport=1, mu=0.01, conv_flux=2, visc_flux=1, rk=2)
    O, S = cases.make_pair(cfg, mesh, ic, cases.bc_freestream(0.3, 0.0, 2, wall_phys=(2,)))
    check(O, S, [0, 1, 2, 3, 4, 5, 6, 7, 10, 11, 13, 14, 18, 19, 20], "ns hybrid 2d")
    with pytest.raises(RuntimeError):
        S.view_variable(S.types[0], 12)      # VelocityZ does not exist in two dimensions


def test_view_variables_euler_with_artificial_viscosity(built):
    """Euler + shock capturing: ArtificialViscosity is the nodal-basis interpolation of the corner values; Vorticity falls through to it"""
    mesh = M.box(2, (10, 8), 0.0, 1.0, phys_bc={1: M.RIEMANN_FARFIELD, 2: M.RIEMANN_FARFIELD, 3: M.ADIABATIC_SLIP_WALL, 4: M.ADIABATIC_SLIP_WALL})
    d = np.ones(2) / np.sqrt(2)

    def ic(x):
        s = np.tanh(((x * d).sum(axis=-1) - 0.5 * d.sum()) / 0.03)
        rho, p = 0.5625 - 0.4375 * s, 0.55 - 0.45 * s
        return np.stack([rho, 0.1 * rho, 0 * rho, 1.4 * p / rho], axis=-1)
    cfg = dict(p=3, conv_flux=2, rk=2, av_tolerance=1.0, av_factor=2.0)
    O, S = cases.make_pair(cfg, mesh, ic, lambda x, phys, time=None: ic(x))
    check(O, S, [0, 1, 2, 3, 4, 5, 6, 9, 7, 10, 11, 13, 14, 16, 19], "euler av 2d")
    t = S.types[0]
    assert S.view_variable(t, 9).max() > 0.0 and np.array_equal(S.view_variable(t, 7), S.view_variable(t, 9))
    assert np.all(S.view_variable(t, 16) == 0.0)


def test_view_variables_weakly_compressible(built):
    """incompressible model: SoundSpeed is the reference sound speed, Entropy falls through to Vorticity"""
    mesh = M.box(2, (5, 4), 0.0, 1.0, phys_bc={k: M.RIEMANN_FARFIELD for k in (1, 2, 3, 4)})
    cfg = dict(p=2, model=3, eos=1, transport=1, mu=0.01, conv_flux=4, visc_flux=2, c0=10.0, rho0=1.0, rk=2)
    ic = cases.ic_perturbed_freestream(0.5, 1.0, 2, amp=0.01)
    O, S = cases.make_pair(cfg, mesh, ic, cases.bc_freestream(0.5, 1.0, 2, wall_phys=()))
    check(O, S, [0, 1, 2, 3, 4, 5, 6, 7, 18], "weakly compressible")
    t = S.types[0]
    assert np.all(S.view_variable(t, 4) == 10.0) and np.array_equal(S.view_variable(t, 6), S.view_variable(t, 7))
