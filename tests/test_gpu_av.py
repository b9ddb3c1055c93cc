"""ShockCapturingEnum::ArtificialViscosity on the CUDA path against the oracle (SpatialDiscrete.cpp:37-192 once per step; eps * grad(U) in
the volume and face fluxes of every stage, :210-253, 694-746, 786-819).  Cases after the reference's shock examples: sod_1d_ceuler (line,
far-field ends), sedovblast / khinstability (quadrangles), plus a hexahedral box."""
import numpy as np
import pytest

import cases
from subrosadg_b200 import mesh as M
from test_gpu_parity import compare, TOL_RES

pytestmark = pytest.mark.gpu


def check_viscosity(O, S, label):
    """the once-per-step part: element corner values and node values of the CURRENT state"""
    O.update_artificial_viscosity(); S.update_artificial_viscosity()
    no, ns = O.node_artificial_viscosity(), S.node_artificial_viscosity()
    assert no.max() > 0.0, f"{label}: the case does not switch the viscosity on"
    if O.cfg["p"] > 1:   # P1: every mode counts as "high", the indicator is log10(1) = 0 = the P1 threshold: half the full value everywhere
        assert (no == 0.0).any(), f"{label}: the case has no smooth region"
    ramp = (no > 0) & (no < no.max() * (1 - 1e-12))
    assert np.array_equal(no == 0.0, ns == 0.0), f"{label}: different elements are flagged"
    assert cases.rel_l2(ns, no) < 1e-10, f"{label}: node_artificial_viscosity_ rel-L2 {cases.rel_l2(ns, no):.3e}"
    for t in S.types:
        assert cases.rel_l2(S.element_artificial_viscosity(t), O.element_artificial_viscosity(t)) < 1e-10
    return int(ramp.sum())


def jump_ic(dim, centre=0.5, width=0.05):
    """a steep but resolved density / pressure jump across an oblique plane, fluid at rest (Sod-like)"""
    d = np.ones(dim) / np.sqrt(dim)

    def f(x):
        s = np.tanh(((x * d).sum(axis=-1) - centre * d.sum()) / width)
        rho = 0.5625 - 0.4375 * s
        p = 0.55 - 0.45 * s
        return np.stack([rho] + [np.zeros_like(rho)] * dim + [1.4 * p / rho], axis=-1)
    return f


@pytest.mark.parametrize("p", [1, 2, 3])
def test_sod_like_1d(built, p):
    mesh = M.box(1, (24,), 0.0, 1.0)
    ic = jump_ic(1, width=0.02 if p < 3 else 0.005)
    bc = lambda x, phys, time=None: ic(x)
    cfg = dict(p=p, conv_flux=2, rk=2, av_tolerance=0.5, av_factor=1.0)
    O, S = cases.make_pair(cfg, mesh, ic, bc)
    check_viscosity(O, S, f"sod 1d p{p}")
    compare(O, S, 0.2 * O.compute_dt(1.0), 6, label=f"sod 1d p{p}", cond=cases.conditioning(cfg, mesh, ic, bc, TOL_RES))
    assert S.node_artificial_viscosity().max() > 0.0     # the stepping path evaluated it as well


@pytest.mark.parametrize("p,warp", [(2, False), (3, False), (3, True)])
def test_oblique_jump_2d_quads(built, p, warp):
    w = (lambda x: x + 0.02 * np.sin(np.pi * x[:, ::-1])) if warp else None
    mesh = M.box(2, (10, 8), 0.0, 1.0, geom_order=2 if warp else 1, warp=w,
                 phys_bc={1: M.RIEMANN_FARFIELD, 2: M.RIEMANN_FARFIELD, 3: M.ADIABATIC_SLIP_WALL, 4: M.ADIABATIC_SLIP_WALL})
    ic = jump_ic(2, width=0.03)
    bc = lambda x, phys, time=None: ic(x)
    cfg = dict(p=p, conv_flux=2, rk=2, av_tolerance=1.0, av_factor=2.0)
    O, S = cases.make_pair(cfg, mesh, ic, bc)
    check_viscosity(O, S, f"jump 2d p{p} warp{warp}")
    compare(O, S, 0.2 * O.compute_dt(1.0), 4, label=f"jump 2d p{p} warp{warp}", cond=cases.conditioning(cfg, mesh, ic, bc, TOL_RES))


@pytest.mark.parametrize("p,flux", [(2, 3), (3, 2)])
def test_oblique_jump_3d_hexes(built, p, flux):
    """P3 hexahedra with artificial viscosity run on the node-per-thread kernels (the line kernels carry no eps terms)"""
    mesh = M.box(3, (5, 4, 4), 0.0, 1.0, periodic_axes=(2,), phys_bc={1: M.RIEMANN_FARFIELD, 2: M.RIEMANN_FARFIELD, 3: M.ADIABATIC_SLIP_WALL, 4: M.ADIABATIC_SLIP_WALL})
    ic = jump_ic(3, width=0.05)
    bc = lambda x, phys, time=None: ic(x)
    cfg = dict(p=p, conv_flux=flux, rk=2, av_tolerance=1.0, av_factor=1.0)
    O, S = cases.make_pair(cfg, mesh, ic, bc)
    check_viscosity(O, S, f"jump 3d p{p}")
    compare(O, S, 0.2 * O.compute_dt(1.0), 3, label=f"jump 3d p{p}", cond=cases.conditioning(cfg, mesh, ic, bc, TOL_RES))


def test_smooth_flow_is_untouched(built):
    """below the indicator threshold the viscosity is exactly zero and the run equals the plain Euler run (other kernels, same numbers)"""
    mesh = M.periodic_box(2, 8)
    ic = cases.ic_density_wave([0.7, 0.3])
    O, S = cases.make_pair(dict(p=3, conv_flux=2, rk=2, av_tolerance=0.2), mesh, ic)
    _, E = cases.make_pair(dict(p=3, conv_flux=2, rk=2), mesh, ic)
    S.stepSolver(1e-3, 5); E.stepSolver(1e-3, 5); O.step(1e-3, 5)
    assert S.node_artificial_viscosity().max() == 0.0 and O.node_artificial_viscosity().max() == 0.0
    t = S.types[0]
    assert cases.rel_l2(S.state_at_quadrature(t), E.state_at_quadrature(t)) < 1e-13
    assert cases.rel_l2(S.state_at_quadrature(t), O.state_at_quadrature(t)) < 1e-10


# ---- dense-operator path: triangles and hybrid triangle / quadrangle meshes (explosion_2d_ceuler, cylinder_2d_ceuler) --------------------
def radial_jump_ic(r_jump=1.5, width=0.08):
    """a steep radial density / pressure jump around the inner boundary of the O-mesh (explosion-like), fluid at rest"""
    def f(x):
        s = np.tanh((np.linalg.norm(x, axis=-1) - r_jump) / width)
        rho = 0.75 - 0.25 * s          # 2 : 1 jumps: the face traces of the under-resolved profile stay positive
        p = 0.75 - 0.25 * s
        return np.stack([rho, np.zeros_like(rho), np.zeros_like(rho), 1.4 * p / rho], axis=-1)
    return f


def compare_all_types(O, S, dt, nsteps, label):
    T = S.types
    for t in T:
        S.set_state(t, O.get_state(t))
    Ro, Rs = O.residual(), S.residual()
    for t in T:
        assert cases.rel_l2(Rs[t][0], Ro[t][0]) < 1e-12, f"{label} type {t}: modal residual rel-L2 {cases.rel_l2(Rs[t][0], Ro[t][0]):.3e}"
        assert cases.rel_l2(Rs[t][1], Ro[t][1]) < 2e-11, f"{label} type {t}: dU/dt rel-L2 {cases.rel_l2(Rs[t][1], Ro[t][1]):.3e}"
    eo, es = O.step(dt, nsteps), S.stepSolver(dt, nsteps)
    for t in T:
        assert cases.rel_l2(S.state_at_quadrature(t), O.state_at_quadrature(t)) < 1e-10, f"{label} type {t}: state"
    assert np.allclose(es, eo, rtol=1e-8, atol=1e-300)


@pytest.mark.parametrize("p", [2, 3])
def test_triangles(built, p):
    """explosion_2d_ceuler-style: MeshModelEnum::Triangle"""
    mesh = M.annulus(6, 16, r0=0.5, r1=3.0, geom_order=1, tri_rings=6)
    assert sorted(mesh.blocks) == [M.TRIANGLE]
    ic = radial_jump_ic(width=0.08 if p == 2 else 0.04)
    bc = lambda x, phys, time=None: ic(x)
    cfg = dict(p=p, conv_flux=2, rk=2, av_tolerance=1.0, av_factor=1.0)
    O, S = cases.make_pair(cfg, mesh, ic, bc)
    check_viscosity(O, S, f"triangles p{p}")
    compare_all_types(O, S, 0.05 * O.compute_dt(1.0), 4, f"triangles p{p}")   # explicit diffusion: well inside eps dt / h^2


def test_hybrid_cylinder(built):
    """cylinder_2d_ceuler-style: MeshModelEnum::TriangleQuadrangle, curved P3 quadrangles at the wall, triangles outside"""
    mesh = M.annulus(6, 16, r0=0.5, r1=3.0, geom_order=3, stretch=1.2, tri_rings=3)
    assert sorted(mesh.blocks) == [M.TRIANGLE, M.QUADRANGLE]
    ic = radial_jump_ic(r_jump=1.6, width=0.05)
    bc = lambda x, phys, time=None: ic(x)
    cfg = dict(p=3, conv_flux=2, rk=2, av_tolerance=1.0, av_factor=2.0)
    O, S = cases.make_pair(cfg, mesh, ic, bc)
    assert check_viscosity(O, S, "hybrid cylinder") >= 0
    compare_all_types(O, S, 0.2 * O.compute_dt(1.0), 3, "hybrid cylinder")
    # the node maximum crossed the element types: some quadrangle corner carries a value set by a triangle or vice versa
    nodes = S.node_artificial_viscosity()
    assert nodes.max() > 0.0


# ---- shock capturing across partitions: the node maximum is reduced over the ranks (the one collective of the path) -----------------------
@pytest.mark.parametrize("dim,shape,world,p,width", [(2, (12, 6), 2, 3, 0.04), (2, (9, 8), 3, 2, 0.04), (3, (6, 4, 3), 2, 2, 0.08), (3, (4, 3, 3), 2, 3, 0.04)])
def test_partitioned_shock_capturing_matches_single_context(built, dim, shape, world, p, width):
    """`world` contexts (ghost elements, part-0 / part-1 launches, element halo of U and of the volume gradient) with the node array
    max-reduced between them after sdg_step_begin (sdg_av_node_buffer / sdg_av_store) against the single-context run"""
    from subrosadg_b200.parallel import InProcessCluster
    from subrosadg_b200.solver import Solver
    far, slip = M.RIEMANN_FARFIELD, M.ADIABATIC_SLIP_WALL
    mesh = M.box(dim, shape, 0.0, 1.0, phys_bc={1: far, 2: far, 3: slip, 4: slip, 5: slip, 6: slip} if dim == 3 else {1: far, 2: far, 3: slip, 4: slip})
    cfg = dict(p=p, conv_flux=2, rk=2, av_tolerance=1.0, av_factor=2.0)
    ic = jump_ic(dim, width=width)
    bc = lambda x, phys, time=None: ic(x)
    S = Solver(dict(cfg), mesh, device=0)
    S.initializeSolver(ic, bc)
    C = InProcessCluster(dict(cfg), mesh, world, device=0)
    C.initializeSolver(ic, bc)
    t = S.types[0]
    dt = 0.2 * S.calculateDeltaTime(1.0)
    err_s = S.stepSolver(dt, 3)
    err_c = C.stepSolver(dt, 3)
    node = S.node_artificial_viscosity()
    assert node.max() > 0.0 and (node == 0.0).any()
    # every context ends with the complete node array: the maximum over ALL elements, wherever they live
    first = C.S[0].node_artificial_viscosity()
    for Sr in C.S:
        assert np.array_equal(Sr.node_artificial_viscosity(), first)
    assert np.array_equal(first == 0.0, node == 0.0) and cases.rel_l2(first, node) < 1e-12   # the states differ by round-off after two steps
    a, b = C.state_at_quadrature(), S.state_at_quadrature(t)
    assert cases.rel_l2(a, b) < 1e-14, f"partitioned vs single context: {cases.rel_l2(a, b):.3e}"
    assert np.allclose(err_c, err_s, rtol=1e-11, atol=1e-300)
