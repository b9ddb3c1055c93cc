"""GPU parity tests of the remaining rows of SURVEY.md 8a on the tensor (quadrangle / hexahedron) path: weakly compressible
EOS + Exact / Lax-Friedrichs / Central flux (P4), Boussinesq source (P8), velocity-inflow / pressure-outflow boundaries (P7),
time-varying boundary values re-uploaded every step (row N, BoundaryCondition.cpp:29-74) — CUDA path vs CPU oracle, BASELINE.json
tolerances."""
import numpy as np
import pytest

import cases
from subrosadg_b200 import mesh as M
from test_gpu_parity import compare, compare_ns

pytestmark = pytest.mark.gpu

WC = dict(eos=1, c0=10.0, rho0=1.0)   # EquationOfState<WeakCompressibleFluid>, PhysicalModel.cpp:57-78


def ic_wc(dim):
    """weakly compressible start: rho = 1 + small wave, smooth velocity, T varying (incompressible examples' variable set)"""
    def f(x):
        s = np.sin(np.pi * x.sum(axis=-1))
        c = np.cos(np.pi * x[..., 0])
        vel = [0.3 + 0.05 * s, -0.2 + 0.05 * c] + ([0.1 + 0.02 * s] if dim == 3 else [])
        return np.stack([1.0 + 1e-3 * s] + vel + [1.0 + 0.1 * c], axis=-1)
    return f


@pytest.mark.parametrize("flux", [0, 1, 4])
def test_incompressible_euler_periodic_2d(built, flux):
    """IncompresibleEuler + WeakCompressibleFluid, Central / Lax-Friedrichs / Exact flux (ConvectiveFlux.cpp:94-134,353-414)"""
    mesh = M.periodic_box(2, 6)
    O, S = cases.make_pair(dict(WC, p=3, model=2, conv_flux=flux, rk=2), mesh, ic_wc(2))
    dt = O.compute_dt(0.5)
    assert abs(S.calculateDeltaTime(0.5) - dt) <= 1e-13 * dt
    compare(O, S, dt, 4, label=f"incompressible euler flux {flux}")


def test_incompressible_euler_periodic_3d(built):
    mesh = M.periodic_box_fast(3, 4)
    O, S = cases.make_pair(dict(WC, p=2, model=2, conv_flux=4, rk=2), mesh, ic_wc(3))
    compare(O, S, 0.3 * O.compute_dt(1.0), 3, label="incompressible euler 3d")


@pytest.mark.parametrize("visc", [1, 2])
def test_incompressible_ns_boussinesq_2d(built, visc):
    """IncompresibleNS + Boussinesq source (examples/rayleighbenard-style): isothermal walls bottom / top, periodic in x"""
    mesh = M.box(2, (6, 5), 0.0, 1.0, periodic_axes=(0,), phys_bc={3: M.ISOTHERMAL_NONSLIP_WALL, 4: M.ISOTHERMAL_NONSLIP_WALL})
    cfg = dict(WC, p=3, model=3, transport=1, mu=0.01, conv_flux=4, visc_flux=visc, source=1, beta=0.5, t_ref=1.0)

    def ic(x):
        s = 1e-2 * np.sin(2 * np.pi * x[..., 0]) * np.sin(np.pi * x[..., 1])
        return np.stack([1.0 + 0.1 * s, s, -s, 1.0 + 0.5 * (0.5 - x[..., 1]) + s], axis=-1)

    def bc(x, phys, time=None):
        one = np.ones(x.shape[:-1])
        return np.stack([one, 0 * one, 0 * one, np.where(phys == 3, 1.25, 0.75) * one], axis=-1)

    O, S = cases.make_pair(cfg, mesh, ic, bc)
    compare_ns(O, S, 0.3 * O.compute_dt(1.0), 3, f"incompressible ns boussinesq visc{visc}")


def test_compressible_ns_boussinesq_3d(built):
    mesh = M.periodic_box_fast(3, 4)
    cfg = dict(p=2, model=1, transport=1, mu=0.01, conv_flux=2, visc_flux=2, source=1, beta=0.3, t_ref=1.0)
    O, S = cases.make_pair(cfg, mesh, cases.ic_density_wave([0.5, 0.3, 0.2]))
    compare_ns(O, S, 5e-4, 3, "compressible ns + boussinesq 3d")


def channel_bc(u_in, p_out_T):
    def f(x, phys, time=None):
        one = np.ones(x.shape[:-1])
        ramp = 1.0 if time is None else 1.0 + 0.5 * np.sin(40.0 * time)
        return np.stack([1.4 * one, u_in * ramp * one, 0 * one, p_out_T * one], axis=-1)
    return f


@pytest.mark.parametrize("model", [0, 1])
def test_inflow_outflow_channel_2d(built, model):
    """velocity inflow (x = 0), pressure outflow (x = 1), slip walls (BoundaryCondition.cpp:313-379,458-471); Euler and NS"""
    mesh = M.box(2, (6, 4), 0.0, 1.0, phys_bc={1: M.VELOCITY_INFLOW, 2: M.PRESSURE_OUTFLOW, 3: M.ADIABATIC_SLIP_WALL, 4: M.ADIABATIC_SLIP_WALL})
    cfg = dict(p=3, model=model, conv_flux=2, rk=2)
    if model == 1:
        cfg.update(transport=1, mu=0.005, visc_flux=2)
    ic = cases.ic_perturbed_freestream(0.3, 0.0, 2)
    O, S = cases.make_pair(cfg, mesh, ic, channel_bc(0.3, 1.0))
    dt = 0.3 * O.compute_dt(1.0)
    if model == 1:
        compare_ns(O, S, dt, 4, f"channel model {model}")
    else:
        compare(O, S, dt, 4, label=f"channel model {model}")


def test_inflow_outflow_channel_3d(built):
    mesh = M.box(3, (4, 3, 3), 0.0, 1.0, phys_bc={1: M.VELOCITY_INFLOW, 2: M.PRESSURE_OUTFLOW, 3: M.ADIABATIC_SLIP_WALL, 4: M.ADIABATIC_SLIP_WALL,
                                                5: M.ADIABATIC_NONSLIP_WALL, 6: M.ISOTHERMAL_NONSLIP_WALL})
    cfg = dict(p=2, model=1, transport=2, mu=0.005, visc_flux=2, conv_flux=3)

    def bc(x, phys, time=None):
        one = np.ones(x.shape[:-1])
        wall = phys >= 5
        return np.stack([1.4 * one, np.where(wall, 0.0, 0.3) * one, 0 * one, 0 * one, np.where(phys == 6, 1.1, 1.0) * one], axis=-1)

    O, S = cases.make_pair(cfg, mesh, cases.ic_perturbed_freestream(0.3, 0.0, 3), bc)
    compare_ns(O, S, 0.3 * O.compute_dt(1.0), 3, "channel 3d")


def test_time_varying_boundary(built):
    """BoundaryTimeEnum::TimeVarying: Solver::updateBoundaryVariable re-evaluates the user callback at t = step * dt before
    every step (BoundaryCondition.cpp:29-74, SystemControl.cpp:175)"""
    mesh = M.box(2, (6, 4), 0.0, 1.0, phys_bc={1: M.VELOCITY_INFLOW, 2: M.PRESSURE_OUTFLOW, 3: M.ADIABATIC_SLIP_WALL, 4: M.ADIABATIC_SLIP_WALL})
    bc = channel_bc(0.3, 1.0)
    O, S = cases.make_pair(dict(p=3, conv_flux=2, rk=2), mesh, cases.ic_perturbed_freestream(0.3, 0.0, 2), bc)
    t = S.types[0]
    S.set_state(t, O.get_state(t))
    dt = 0.3 * O.compute_dt(1.0)
    for step in range(1, 5):
        O.update_boundary(bc, step * dt); S.updateBoundaryVariable(bc, step * dt)
        eo = O.step(dt, 1); es = S.stepSolver(dt, 1)
        assert np.allclose(es, eo, rtol=1e-9, atol=1e-300)
    assert cases.rel_l2(S.state_at_quadrature(t), O.state_at_quadrature(t)) < 1e-10
    # the boundary values did change the solution: a frozen-boundary run differs
    O2, S2 = cases.make_pair(dict(p=3, conv_flux=2, rk=2), mesh, cases.ic_perturbed_freestream(0.3, 0.0, 2), bc)
    S2.set_state(t, O2.get_state(t)); S2.stepSolver(dt, 4)
    assert cases.rel_l2(S2.state_at_quadrature(t), S.state_at_quadrature(t)) > 1e-6


# ---- PolynomialOrderEnum P4 / P5 (src/Utils/Enum.cpp): tensor path, quadrangles and hexahedra ---------------------------------
@pytest.mark.parametrize("p", [4, 5])
def test_high_order_euler_2d(built, p):
    mesh = M.box(2, (5, 4), 0.0, 1.0, geom_order=2, warp=lambda x: x + 0.03 * np.sin(np.pi * x[:, ::-1]),
                 phys_bc={1: M.RIEMANN_FARFIELD, 2: M.RIEMANN_FARFIELD, 3: M.ADIABATIC_SLIP_WALL, 4: M.RIEMANN_FARFIELD})
    ic = cases.ic_perturbed_freestream(0.5, 2.0, 2)
    O, S = cases.make_pair(dict(p=p, conv_flux=3, rk=2), mesh, ic, cases.bc_freestream(0.5, 2.0, 2, wall_phys=(3,)))
    dt = O.compute_dt(0.5)
    assert abs(S.calculateDeltaTime(0.5) - dt) <= 1e-13 * dt
    compare(O, S, dt, 4, label=f"high order euler 2d p{p}")


@pytest.mark.parametrize("p", [4, 5])
def test_high_order_periodic_3d(built, p):
    mesh = M.periodic_box_fast(3, 3)
    O, S = cases.make_pair(dict(p=p, conv_flux=2, rk=2), mesh, cases.ic_density_wave([0.5, 0.3, 0.2]))
    # the modal <-> collocation map of 216 Lobatto products is worse conditioned than at P3: 1.7e-12 on the IC coefficients at P5
    compare(O, S, 5e-4, 3, label=f"high order euler 3d p{p}", tol_ic=5e-12)


@pytest.mark.parametrize("p,dim", [(4, 2), (5, 2), (4, 3), (5, 3)])
def test_high_order_cns(built, p, dim):
    mesh = M.periodic_box(2, 4) if dim == 2 else M.periodic_box_fast(3, 3)
    vel = [0.7, 0.3] if dim == 2 else [0.5, 0.3, 0.2]
    cfg = dict(p=p, model=1, transport=1, mu=0.01, conv_flux=2, visc_flux=2, rk=2)
    O, S = cases.make_pair(cfg, mesh, cases.ic_density_wave(vel))
    compare_ns(O, S, 2e-4, 3, f"high order cns p{p} dim{dim}")


# ---- DimensionEnum::D1: line elements (LineTrait), point faces ---------------------------------------------------------------------
def ic_wave_1d(x):
    rho = 1.0 + 0.2 * np.sin(np.pi * x[..., 0])
    return np.stack([rho, 0.5 + 0 * rho, 1.4 / rho], axis=-1)


@pytest.mark.parametrize("p", [1, 2, 3, 4, 5])
def test_periodic_1d_ceuler(built, p):
    mesh = M.box(1, (12,), 0.0, 2.0, periodic_axes=(0,))
    O, S = cases.make_pair(dict(p=p, conv_flux=2, rk=2), mesh, ic_wave_1d)
    dt = O.compute_dt(0.5)
    assert abs(S.calculateDeltaTime(0.5) - dt) <= 1e-13 * dt
    compare(O, S, dt, 6, label=f"periodic_1d p{p}", tol_ic=5e-12)


@pytest.mark.parametrize("flux", [0, 1, 3])
def test_fluxes_1d_curved(built, flux):
    mesh = M.box(1, (9,), 0.0, 2.0, periodic_axes=(0,), geom_order=3, warp=lambda x: x + 0.05 * np.sin(np.pi * x))
    O, S = cases.make_pair(dict(p=3, conv_flux=flux, rk=1), mesh, ic_wave_1d)
    compare(O, S, 1e-3, 4, label=f"1d curved flux {flux}")


@pytest.mark.parametrize("visc", [1, 2])
def test_shock_tube_like_1d_cns(built, visc):
    """1-D Navier-Stokes with far-field ends (sod_1d-style domain without the artificial viscosity the shipped example adds)"""
    mesh = M.box(1, (10,), 0.0, 1.0)

    def ic(x):
        s = np.tanh((x[..., 0] - 0.5) / 0.1)
        return np.stack([1.0 - 0.3 * s, 0.1 + 0 * s, 1.0 - 0.1 * s], axis=-1)

    def bc(x, phys, time=None):
        left = x[..., 0] < 0.5
        return np.stack([np.where(left, 1.3, 0.7), 0.1 + 0 * x[..., 0], np.where(left, 1.1, 0.9)], axis=-1)

    cfg = dict(p=3, model=1, transport=1, mu=0.01, conv_flux=2, visc_flux=visc, rk=2)
    O, S = cases.make_pair(cfg, mesh, ic, bc)
    compare_ns(O, S, 0.2 * O.compute_dt(1.0), 4, f"1d cns visc{visc}")
