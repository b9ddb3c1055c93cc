"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical meshes / initial conditions.
Tolerances are BASELINE.json's: per-stage residual rel-L2 <= 1e-12, conserved fields after N steps <= 1e-10 (fp64)."""
import numpy as np
import pytest

import cases
from subrosadg_b200 import mesh as M

pytestmark = pytest.mark.gpu

TOL_RES = 1e-12     # per-stage residual (variable_residual_, SpatialDiscrete.cpp:1016-1032), BASELINE.json
TOL_STATE = 1e-10   # conserved fields after N steps, BASELINE.json
# dU/dt = R M^-1 seen at the quadrature points.  The reference algorithm (and the oracle) applies a dense modal M^-1 whose
# condition number (1e3..1e4 for P3 hexahedra) multiplies the round-off of R, so this quantity is only reproducible to
# cond(M) * eps between ANY two fp64 implementations; the CUDA path's M is diagonal (no amplification).  See DESIGN.md.
TOL_RHS = 2e-11


def compare(O, S, dt, nsteps, tol_res=TOL_RES, tol_state=TOL_STATE, label="", tol_ic=1e-12, cond=1.0):
    """cond >= 1 loosens the residual tolerances for ill-conditioned set-ups: cases.conditioning() measures how far the ORACLE's own
    residual moves when its inputs move by one unit round-off; no two fp64 implementations can agree better than a few times that."""
    t = S.types[0]
    tol_res = tol_res * cond
    # Solver::initializeSolver parity: modal coefficients of the IC projection (same H1Legendre convention on both sides)
    assert cases.rel_l2(S.get_state(t), O.get_state(t)) < tol_ic, label
    # identical inputs from here on: hand the oracle's modal coefficients to the CUDA path through the seam
    S.set_state(t, O.get_state(t))
    Ro, qo = O.residual()[t]
    Rs, qs = S.residual()[t]
    e_q, e_R = cases.rel_l2(qs, qo), cases.rel_l2(Rs, Ro)
    assert e_R < tol_res, f"{label}: modal residual rel-L2 {e_R:.3e}"
    assert e_q < TOL_RHS * cond, f"{label}: dU/dt at quadrature points rel-L2 {e_q:.3e}"
    err_o = O.step(dt, nsteps)
    err_s = S.stepSolver(dt, nsteps)
    e_u = cases.rel_l2(S.state_at_quadrature(t), O.state_at_quadrature(t))
    assert e_u < tol_state, f"{label}: state after {nsteps} steps rel-L2 {e_u:.3e}"
    assert cases.rel_l2(S.get_state(t), O.get_state(t)) < tol_state, label
    assert np.allclose(err_s, err_o, rtol=1e-9, atol=1e-300), f"{label}: relative_error_ {err_s} vs {err_o}"
    return e_q, e_R, e_u


@pytest.mark.parametrize("p", [1, 2, 3])
def test_periodic_2d_ceuler(built, p):
    """config 1: examples/periodic_2d_ceuler.cpp (10x10 quads on [0,2]^2, HLLC, SSPRK3, dt = 1e-3)"""
    mesh = M.periodic_box(2, 10)
    O, S = cases.make_pair(dict(p=p, conv_flux=2, rk=2), mesh, cases.ic_density_wave([0.7, 0.3]))
    assert abs(S.calculateDeltaTime(1.0) - O.compute_dt(1.0)) <= 1e-14 * O.compute_dt(1.0)
    compare(O, S, 1e-3, 10, label=f"periodic_2d p{p}")


@pytest.mark.parametrize("p", [1, 2, 3])
def test_periodic_3d_ceuler(built, p):
    """config 4 (scaled down): periodic hex box, HLLC, SSPRK3"""
    mesh = M.periodic_box_fast(3, 6)
    O, S = cases.make_pair(dict(p=p, conv_flux=2, rk=2), mesh, cases.ic_density_wave([0.5, 0.3, 0.2]))
    compare(O, S, 1e-3, 5, label=f"periodic_3d p{p}")


@pytest.mark.parametrize("flux", [0, 1, 2, 3])
@pytest.mark.parametrize("rk", [0, 1, 2])
def test_fluxes_and_rk(built, flux, rk):
    mesh = M.periodic_box(2, 6)
    O, S = cases.make_pair(dict(p=2, conv_flux=flux, rk=rk), mesh, cases.ic_density_wave([0.7, 0.3]))
    compare(O, S, 5e-4, 4, label=f"flux {flux} rk {rk}")


def test_curved_box_farfield_2d(built):
    warp = lambda x: x + 0.04 * np.sin(np.pi * x[:, ::-1])
    mesh = M.box(2, (6, 5), 0.0, 1.0, geom_order=3, warp=warp)
    ic = cases.ic_perturbed_freestream(0.63, 2.0, 2)
    O, S = cases.make_pair(dict(p=3, conv_flux=2, rk=2), mesh, ic, cases.bc_freestream(0.63, 2.0, 2, wall_phys=()))
    compare(O, S, 1e-3, 5, label="curved box farfield")


def test_naca0012_2d_ceuler(built):
    """config 2 (scaled down): curved P3 quads, HLLC, Riemann far field + slip wall, M = 0.63, alpha = 2 deg"""
    mesh = M.naca0012(nr=8, nt=24)
    ic = cases.ic_perturbed_freestream(0.63, 2.0, 2, amp=1e-3)
    O, S = cases.make_pair(dict(p=3, conv_flux=2, rk=2), mesh, ic, cases.bc_freestream(0.63, 2.0, 2))
    dt = O.compute_dt(0.5)
    assert abs(S.calculateDeltaTime(0.5) - dt) <= 1e-13 * dt
    compare(O, S, dt, 5, label="naca0012")


def test_curved_hex_box_farfield_3d(built):
    warp = lambda x: x + 0.03 * np.sin(np.pi * np.roll(x, 1, axis=1))
    mesh = M.box(3, (3, 4, 3), 0.0, 1.0, geom_order=2, warp=warp)
    ic = cases.ic_perturbed_freestream(0.5, 3.0, 3)
    O, S = cases.make_pair(dict(p=2, conv_flux=3, rk=2), mesh, ic, cases.bc_freestream(0.5, 3.0, 3, wall_phys=()))
    compare(O, S, 1e-3, 4, label="curved hex box")


def test_modal_state_roundtrip(built):
    mesh = M.periodic_box(2, 6)
    O, S = cases.make_pair(dict(p=3), mesh, cases.ic_density_wave([0.7, 0.3]))
    t = S.types[0]
    U = O.get_state(t)
    S.set_state(t, U * 1.25)
    assert cases.rel_l2(S.get_state(t), U * 1.25) < 1e-14


# ---- Navier-Stokes (BR1 / BR2) ----------------------------------------------------------------------------------------------
NS = dict(model=1, transport=1, mu=1.4 * 0.2 / 200.0)   # examples/sphere_3d_cns.cpp:57-60 (Re = 200)


def compare_ns(O, S, dt, nsteps, label, cond=1.0):
    t = S.types[0]
    S.set_state(t, O.get_state(t))
    Ro, qo = O.residual()[t]
    go = O.gradient_at_quadrature(t)
    gs = S.gradient_at_quadrature(t)
    Rs, qs = S.residual()[t]
    e_g = cases.rel_l2(gs, go)
    assert e_g < 1e-11 * cond, f"{label}: total gradient at quadrature points rel-L2 {e_g:.3e}"
    e_m = cases.rel_l2(S.gradient_state(t), O.gradient_state(t))    # the RawBinary gradient blocks (RawBinary.cpp:75-154)
    # nodal -> modal goes through the inverse Vandermonde matrix of the H1-Legendre basis: the nodal error is amplified by up to its
    # 2-norm condition number (59 for P3 quadrangles, 458 / 1633 / 3591 for P3 / P4 / P5 hexahedra; measured errors are 7e-14..3e-13 times it)
    tol_m = max(1e-11, 1e-12 * np.linalg.cond(O.table(t, 0))) * cond
    assert e_m < tol_m, f"{label}: modal gradient coefficients rel-L2 {e_m:.3e}"
    bo, bs = O.boundary_gradient_state(), S.boundary_gradient_state()
    if bo.size:
        # a handful of elements (in 1-D: two) instead of the whole field: the same absolute error against a smaller norm
        assert cases.rel_l2(bs, bo) < 10 * tol_m, f"{label}: boundary-parent gradient blocks rel-L2 {cases.rel_l2(bs, bo):.3e}"
    e_R, e_q = cases.rel_l2(Rs, Ro), cases.rel_l2(qs, qo)
    assert e_R < TOL_RES * cond, f"{label}: modal residual rel-L2 {e_R:.3e}"
    assert e_q < TOL_RHS * cond, f"{label}: dU/dt rel-L2 {e_q:.3e}"
    err_o = O.step(dt, nsteps)
    err_s = S.stepSolver(dt, nsteps)
    e_u = cases.rel_l2(S.state_at_quadrature(t), O.state_at_quadrature(t))
    assert e_u < TOL_STATE, f"{label}: state after {nsteps} steps rel-L2 {e_u:.3e}"
    assert np.allclose(err_s, err_o, rtol=1e-8, atol=1e-300), f"{label}: relative_error_ {err_s} vs {err_o}"


@pytest.mark.parametrize("visc,transport", [(2, 1), (1, 1), (2, 2)])
@pytest.mark.parametrize("p", [2, 3])
def test_periodic_2d_cns(built, visc, transport, p):
    mesh = M.periodic_box(2, 6)
    cfg = dict(NS, p=p, visc_flux=visc, transport=transport, mu=0.01)
    O, S = cases.make_pair(cfg, mesh, cases.ic_density_wave([0.7, 0.3]))
    compare_ns(O, S, 5e-4, 4, f"periodic_2d_cns visc{visc} transport{transport} p{p}")


@pytest.mark.parametrize("visc", [1, 2])
@pytest.mark.parametrize("p", [1, 2, 3])
def test_periodic_3d_cns(built, visc, p):
    mesh = M.periodic_box_fast(3, 4)
    cfg = dict(NS, p=p, visc_flux=visc, mu=0.01)
    O, S = cases.make_pair(cfg, mesh, cases.ic_density_wave([0.5, 0.3, 0.2]))
    compare_ns(O, S, 5e-4, 3, f"periodic_3d_cns visc{visc} p{p}")


@pytest.mark.parametrize("wall", [M.ADIABATIC_NONSLIP_WALL, M.ISOTHERMAL_NONSLIP_WALL, M.ADIABATIC_SLIP_WALL])
def test_walls_2d_cns(built, wall):
    """flat-plate style box: wall on the bottom (physical 3), far field elsewhere (examples/blasius_2d_cns.cpp family)"""
    mesh = M.box(2, (6, 5), 0.0, 1.0, phys_bc={1: M.RIEMANN_FARFIELD, 2: M.RIEMANN_FARFIELD, 3: wall, 4: M.RIEMANN_FARFIELD})
    cfg = dict(NS, p=3, visc_flux=2, mu=0.005)
    ic = cases.ic_perturbed_freestream(0.3, 0.0, 2)
    O, S = cases.make_pair(cfg, mesh, ic, cases.bc_freestream(0.3, 0.0, 2, wall_phys=(3,)))
    compare_ns(O, S, 2e-4, 4, f"walls wall{wall}")


def test_karman_like_quads_2d_cns(built):
    """config 3 without its triangles: curved P3 quad ring around a cylinder, BR2, Sutherland, no-slip wall + far field"""
    mesh = M.annulus(5, 16, r0=0.5, r1=4.0, geom_order=3, stretch=1.5, phys_bc={1: M.RIEMANN_FARFIELD, 2: M.ADIABATIC_NONSLIP_WALL})
    cfg = dict(NS, p=3, visc_flux=2, transport=2)
    ic = cases.ic_perturbed_freestream(0.2, 0.0, 2, amp=1e-3)
    O, S = cases.make_pair(cfg, mesh, ic, cases.bc_freestream(0.2, 0.0, 2, wall_phys=(2,)))
    dt = 0.3 * O.compute_dt(1.0)
    compare_ns(O, S, dt, 3, "karman quads")


def test_sphere_3d_cns(built):
    """config 5 (scaled down): curved P3 hexahedra around a sphere, BR2, constant viscosity, no-slip wall + far field"""
    mesh = M.cubed_sphere_shell(3, 3, r0=0.5, r1=4.0, geom_order=3)
    cfg = dict(NS, p=3, visc_flux=2)
    ic = cases.ic_perturbed_freestream(0.2, 0.0, 3, amp=1e-3)
    O, S = cases.make_pair(cfg, mesh, ic, cases.bc_freestream(0.2, 0.0, 3, wall_phys=(2,)))
    dt = 0.3 * O.compute_dt(1.0)
    compare_ns(O, S, dt, 2, "sphere_3d_cns")


# ---- element-block partitions (multi-rank path on one device) -----------------------------------------------------------------
@pytest.mark.parametrize("dim,n,world,model", [(2, 8, 2, {}), (3, 6, 3, {}), (3, 4, 2, dict(NS, visc_flux=2, mu=0.01)), (2, 8, 4, dict(NS, visc_flux=1, mu=0.01)),
                                               (3, 4, 2, dict(NS, visc_flux=2, mu=0.01, p=3)), (3, 6, 3, dict(NS, visc_flux=1, mu=0.01, p=3)), (3, 4, 2, dict(p=3))])
def test_two_contexts_match_single_context(built, dim, n, world, model):
    """`world` contexts with ghost elements, pack kernel and part-0 / part-1 launches (the N>1 path of bench.py, with
    device copies in place of NCCL) must reproduce the single-context run element for element."""
    from subrosadg_b200.parallel import InProcessCluster
    from subrosadg_b200.solver import Solver
    mesh = M.periodic_box_fast(dim, n)
    cfg = dict(p=3 if dim == 2 else 2, conv_flux=2, rk=2)
    cfg.update(model)   # p = 3 in 3-D: eulerLineKernel / the trace-based NS line kernels (halo = face-trace rows)
    vel = [0.7, 0.3] if dim == 2 else [0.5, 0.3, 0.2]
    ic = cases.ic_density_wave(vel)
    S = Solver(dict(cfg), mesh, device=0)
    S.initializeSolver(ic)
    C = InProcessCluster(dict(cfg), mesh, world, device=0)
    C.initializeSolver(ic)
    t = S.types[0]
    assert abs(C.calculateDeltaTime(1.0) - S.calculateDeltaTime(1.0)) == 0.0
    err_s = S.stepSolver(5e-4, 3)
    err_c = C.stepSolver(5e-4, 3)
    a, b = C.state_at_quadrature(), S.state_at_quadrature(t)
    assert cases.rel_l2(a, b) < 1e-14, f"partitioned vs single context: {cases.rel_l2(a, b):.3e}"
    assert np.allclose(err_c, err_s, rtol=1e-11, atol=1e-300)
