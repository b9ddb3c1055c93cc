"""GPU parity tests of the dense-operator path (triangle blocks and hybrid triangle / quadrangle meshes) through the C ABI
against the CPU oracle: BASELINE configs[1] hybrid variant (naca0012-style Euler, Roe / HLLC, far field + slip wall) and
configs[2] (karmanvortex_2d_cns: tri + quad, HLLC, BR2, Sutherland, far field + adiabatic no-slip wall).
Tolerances are BASELINE.json's: per-stage residual rel-L2 <= 1e-12, conserved fields after N steps <= 1e-10 (fp64)."""
import numpy as np
import pytest

import cases
import oracle
from subrosadg_b200 import mesh as M
from subrosadg_b200.solver import Solver

pytestmark = pytest.mark.gpu

TOL_RES, TOL_STATE, TOL_RHS = 1e-12, 1e-10, 2e-11
NS = dict(model=1, transport=1, mu=1.4 * 0.2 / 200.0)


def pair(cfg, mesh, ic, bc=None, dense=False):
    O = oracle.Oracle(dict(cfg), mesh)
    S = Solver(dict(cfg, chunk=-1) if dense else dict(cfg), mesh, device=0)
    O.initialize(ic, bc)
    S.initializeSolver(ic, bc)
    return O, S


def all_types(fn_s, fn_o, types):
    a = np.concatenate([np.ravel(fn_s(t)) for t in types])
    b = np.concatenate([np.ravel(fn_o(t)) for t in types])
    return cases.rel_l2(a, b)


def compare(O, S, dt, nsteps, label, ns=False, cond=1.0):
    T = S.types
    assert all_types(S.get_state, O.get_state, T) < 1e-12, f"{label}: IC projection"
    for t in T:
        S.set_state(t, O.get_state(t))
    Ro = O.residual()   # also fills the oracle's gradient coefficients
    if ns:
        e_g = all_types(S.gradient_at_quadrature, O.gradient_at_quadrature, T)
        assert e_g < 1e-11 * cond, f"{label}: total gradient at quadrature points rel-L2 {e_g:.3e}"
        e_m = all_types(S.gradient_state, O.gradient_state, T)    # the RawBinary gradient blocks (RawBinary.cpp:75-154)
        # nodal -> modal goes through the inverse Vandermonde matrix of the H1-Legendre basis (2-norm condition number 59 for P3 quadrangles,
        # 458 for P3 hexahedra): the nodal tolerance above, times ten
        assert e_m < 1e-10 * cond, f"{label}: modal gradient coefficients rel-L2 {e_m:.3e}"
        bo, bs = O.boundary_gradient_state(), S.boundary_gradient_state()
        if bo.size:
            assert cases.rel_l2(bs, bo) < 1e-10 * cond, f"{label}: boundary-parent gradient blocks rel-L2 {cases.rel_l2(bs, bo):.3e}"
    Rs = S.residual()
    e_R = all_types(lambda t: Rs[t][0], lambda t: Ro[t][0], T)
    e_q = all_types(lambda t: Rs[t][1], lambda t: Ro[t][1], T)
    assert e_R < TOL_RES * cond, f"{label}: modal residual rel-L2 {e_R:.3e}"
    assert e_q < TOL_RHS * cond, f"{label}: dU/dt at quadrature points rel-L2 {e_q:.3e}"
    err_o = O.step(dt, nsteps)
    err_s = S.stepSolver(dt, nsteps)
    e_u = all_types(S.state_at_quadrature, O.state_at_quadrature, T)
    assert e_u < TOL_STATE, f"{label}: state after {nsteps} steps rel-L2 {e_u:.3e}"
    assert all_types(S.get_state, O.get_state, T) < TOL_STATE, label
    assert np.allclose(err_s, err_o, rtol=1e-8, atol=1e-300), f"{label}: relative_error_ {err_s} vs {err_o}"


@pytest.mark.parametrize("p", [1, 2, 3])
def test_dense_path_on_quads_matches_oracle(built, p):
    """the dense-operator kernels on the quad mesh of config 1 (cfg.chunk = -1 forces them)"""
    mesh = M.periodic_box(2, 8)
    O, S = pair(dict(p=p, conv_flux=2, rk=2), mesh, cases.ic_density_wave([0.7, 0.3]), dense=True)
    assert abs(S.calculateDeltaTime(1.0) - O.compute_dt(1.0)) <= 1e-13 * O.compute_dt(1.0)
    compare(O, S, 1e-3, 5, f"dense quads p{p}")


def test_dense_path_matches_tensor_path(built):
    """two device paths, one discrete operator: collocation tensor kernels vs dense modal kernels on the same quads"""
    mesh = M.box(2, (6, 5), 0.0, 1.0, geom_order=3, warp=lambda x: x + 0.04 * np.sin(np.pi * x[:, ::-1]))
    ic = cases.ic_perturbed_freestream(0.4, 2.0, 2)
    bc = cases.bc_freestream(0.4, 2.0, 2, wall_phys=())
    cfg = dict(NS, p=3, visc_flux=2, mu=0.01)
    A = Solver(dict(cfg), mesh, device=0); A.initializeSolver(ic, bc)
    B = Solver(dict(cfg, chunk=-1), mesh, device=0); B.initializeSolver(ic, bc)
    t = M.QUADRANGLE
    assert cases.rel_l2(A.get_state(t), B.get_state(t)) < 1e-12
    B.set_state(t, A.get_state(t))
    ea = A.stepSolver(2e-4, 4); eb = B.stepSolver(2e-4, 4)
    assert cases.rel_l2(A.state_at_quadrature(t), B.state_at_quadrature(t)) < 1e-11
    assert np.allclose(ea, eb, rtol=1e-8)


@pytest.mark.parametrize("p", [1, 2, 3])
def test_triangles_euler(built, p):
    """triangle-only O-mesh, far field + slip wall"""
    mesh = M.annulus(3, 12, r0=0.5, r1=3.0, geom_order=1, tri_rings=3)
    assert sorted(mesh.blocks) == [M.TRIANGLE]
    ic = cases.ic_perturbed_freestream(0.4, 2.0, 2)
    O, S = pair(dict(p=p, conv_flux=2, rk=2), mesh, ic, cases.bc_freestream(0.4, 2.0, 2))
    dt = O.compute_dt(0.5)
    assert abs(S.calculateDeltaTime(0.5) - dt) <= 1e-13 * dt
    compare(O, S, dt, 5, f"triangles p{p}")


@pytest.mark.parametrize("flux", [1, 2, 3])
@pytest.mark.parametrize("rk", [0, 2])
def test_hybrid_ceuler(built, flux, rk):
    """config 2, hybrid variant BASELINE.json names: curved P3 quads at the wall + triangles outside, Lax-Friedrichs / HLLC / Roe,
    Riemann far field + slip wall, M = 0.63, alpha = 2 deg (examples/naca0012_2d_ceuler.cpp:19-33)"""
    mesh = M.annulus(5, 16, r0=0.5, r1=4.0, geom_order=3, stretch=1.5, tri_rings=2)
    assert sorted(mesh.blocks) == [M.TRIANGLE, M.QUADRANGLE]
    ic = cases.ic_perturbed_freestream(0.63, 2.0, 2, amp=1e-3)
    O, S = pair(dict(p=3, conv_flux=flux, rk=rk), mesh, ic, cases.bc_freestream(0.63, 2.0, 2))
    dt = O.compute_dt(0.5)
    assert abs(S.calculateDeltaTime(0.5) - dt) <= 1e-13 * dt
    compare(O, S, dt, 4, f"hybrid euler flux{flux} rk{rk}")


@pytest.mark.parametrize("visc,transport,wall", [(2, 2, M.ADIABATIC_NONSLIP_WALL), (1, 1, M.ADIABATIC_NONSLIP_WALL), (2, 1, M.ISOTHERMAL_NONSLIP_WALL)])
def test_karmanvortex_2d_cns(built, visc, transport, wall):
    """config 3 (scaled down): P3 tri + quad, HLLC, BR2 (and BR1), Sutherland (and constant) viscosity, Re = 200,
    far field + no-slip cylinder wall (examples/karmanvortex_2d_cns.cpp:19-28,56-60)"""
    mesh = M.annulus(5, 16, r0=0.5, r1=4.0, geom_order=3, stretch=1.5, tri_rings=2,
                     phys_bc={1: M.RIEMANN_FARFIELD, 2: wall})
    cfg = dict(NS, p=3, visc_flux=visc, transport=transport)
    ic = cases.ic_perturbed_freestream(0.2, 0.0, 2, amp=1e-3)
    O, S = pair(cfg, mesh, ic, cases.bc_freestream(0.2, 0.0, 2, wall_phys=(2,)))
    dt = 0.3 * O.compute_dt(1.0)
    compare(O, S, dt, 3, f"karman hybrid visc{visc} transport{transport} wall{wall}", ns=True)


def test_hybrid_incompressible_boussinesq(built):
    """weakly compressible EOS + Exact flux + Boussinesq source on a hybrid mesh (P4, P8 rows of SURVEY 8a)"""
    mesh = M.annulus(4, 12, r0=0.5, r1=3.0, geom_order=2, tri_rings=2, phys_bc={1: M.RIEMANN_FARFIELD, 2: M.ISOTHERMAL_NONSLIP_WALL})
    cfg = dict(p=2, model=3, eos=1, transport=1, mu=0.01, conv_flux=4, visc_flux=2, source=1, c0=10.0, rho0=1.0, beta=0.5, t_ref=1.0)

    def ic(x):
        s = 1e-2 * np.sin(np.pi * x[..., 0]) * np.cos(np.pi * x[..., 1])
        return np.stack([1.0 + 0.1 * s, 0.1 + s, 0.05 - s, 1.0 + s], axis=-1)

    def bc(x, phys, time=None):
        one = np.ones(x.shape[:-1])
        wall = phys == 2
        return np.stack([one, np.where(wall, 0.0, 0.1), np.where(wall, 0.0, 0.05), np.where(wall, 1.2, 1.0) * one], axis=-1)

    O, S = pair(cfg, mesh, ic, bc)
    dt = 0.3 * O.compute_dt(1.0)
    compare(O, S, dt, 3, "hybrid incompressible boussinesq", ns=True)


def test_mixed_modal_state_roundtrip_and_launches(built):
    mesh = M.annulus(4, 12, tri_rings=2)
    S = Solver(dict(p=3), mesh, device=0)
    S.initializeSolver(cases.ic_perturbed_freestream(0.4, 0.0, 2), cases.bc_freestream(0.4, 0.0, 2))
    for t in S.types:
        U = S.get_state(t)
        S.set_state(t, U * 1.25)
        assert cases.rel_l2(S.get_state(t), U * 1.25) == 0.0
    n0 = S.launch_count
    S.stepSolver(1e-4, 2)
    assert S.launch_count - n0 == 2 * 3 * 3 + 1   # per stage: face kernel + one element kernel per type; + norm reduction
