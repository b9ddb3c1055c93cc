"""sdg_step_host: one stepSolver on a state held in host memory, streamed (upload groups -> dependency levels of the thread-block chunks
-> downloads).  CPU: the levels of a plan-only context.  GPU: bit-identical to sdg_set_state -> sdg_step -> sdg_get_state."""
import os

import numpy as np
import pytest

import cases
from subrosadg_b200 import mesh as M
from subrosadg_b200.solver import Solver

HEX = M.HEXAHEDRON
NS = dict(model=1, transport=1, mu=1.4 * 0.2 / 200.0, visc_flux=2)   # CompresibleNS, constant viscosity, BR2 (the north-star kernel pair)


def _with_groups(g):
    os.environ["SDG_HOST_PIPE_GROUPS"] = str(g)


def test_levels_of_a_structured_cube_stream(built):
    """24^3 periodic cube, 12 upload groups of two element layers = one layer of 2 x 2 x 2 chunks each: a stage moves the dependency by
    one chunk layer, so group g goes back at level g + 3; the groups 0-2 next to the periodic wrap and the last ones wait for the end:
    5 of the 12 groups travel back while later groups are still arriving."""
    _with_groups(12)
    try:
        S = Solver(dict(p=3, conv_flux=2, rk=2), M.periodic_box_fast(3, 24), device=-1)
        groups, early = S.step_host_info()
        assert groups == 12
        assert abs(early - 5.0 / 12.0) < 1e-12, early
    finally:
        del os.environ["SDG_HOST_PIPE_GROUPS"]


def test_two_pass_levels(built):
    """Navier-Stokes: two launches per stage, so the dependency moves twice as far: group g of 12 goes back at level g + 6"""
    _with_groups(12)
    try:
        S = Solver(dict(p=3, conv_flux=2, rk=2, **NS), M.periodic_box_fast(3, 24), device=-1)
        groups, early = S.step_host_info()
        assert groups == 12 and early == 0.0     # 24 cells are too few for six chunk layers each way: everything waits for the wrap
        S = Solver(dict(p=3, conv_flux=2, rk=0, **NS), M.periodic_box_fast(3, 24), device=-1)   # forward Euler: two launches
        assert abs(S.step_host_info()[1] - 7.0 / 12.0) < 1e-12
        # a mesh with boundary faces: the virtual neighbour traces are launched per level too
        S = Solver(dict(p=3, conv_flux=2, rk=0, **NS), M.box(3, (24, 24, 24), 0.0, 2.0), device=-1)
        groups, early = S.step_host_info()
        assert groups == 12 and early > 0.5     # no periodic wrap: only the last groups wait
    finally:
        del os.environ["SDG_HOST_PIPE_GROUPS"]


def test_contexts_that_do_not_stream_report_zero_groups(built):
    S = Solver(dict(p=3, conv_flux=2, rk=2), M.periodic_box_fast(3, 8), device=-1)    # 512 elements: not worth streaming
    assert S.step_host_info()[0] == 0
    S = Solver(dict(p=2, conv_flux=2, rk=2, av_tolerance=1.0), M.periodic_box_fast(3, 24), device=-1)   # shock capturing: a pass over the whole mesh per step
    assert S.step_host_info()[0] == 0


def test_node_kernel_contexts_stream_too(built):
    """the chunks' face lists give the same dependency for the node-per-thread kernels: P2 hexahedra, P3 quadrangles"""
    _with_groups(8)
    try:
        S = Solver(dict(p=2, conv_flux=2, rk=2), M.periodic_box_fast(3, 24), device=-1)
        groups, early = S.step_host_info()
        assert groups == 8 and 0.0 < early < 1.0
        S = Solver(dict(p=3, conv_flux=2, rk=2), M.periodic_box_fast(2, 128), device=-1)
        groups, early = S.step_host_info()
        assert groups == 8 and 0.0 < early < 1.0
    finally:
        del os.environ["SDG_HOST_PIPE_GROUPS"]


@pytest.mark.gpu
@pytest.mark.parametrize("rk,groups,extra", [(2, 6, {}), (2, 1, {}), (1, 5, {}), (0, 4, {}), (2, 6, NS), (1, 3, dict(NS, visc_flux=1))])
def test_streamed_step_is_bit_identical(built, rk, groups, extra):
    _with_groups(groups)
    try:
        mesh = M.periodic_box_fast(3, 24)
        cfg = dict(p=3, conv_flux=2, rk=rk, **extra)
        S = Solver(cfg, mesh, device=0)
        S.initializeSolver(cases.ic_density_wave([0.5, 0.3, 0.2]))
        assert S.step_host_info()[0] == groups
        dt = S.calculateDeltaTime(1.0)
        U0 = S.get_state(HEX).copy()
        # phase after phase
        S.set_state(HEX, U0)
        e_ref = S.stepSolver(dt, 1).copy()
        U_ref = S.get_state(HEX).copy()
        S.set_state(HEX, U_ref)                                   # the host round trip is part of what is compared
        e_ref2 = S.stepSolver(dt, 1).copy()
        U_ref2 = S.get_state(HEX).copy()
        # streamed, twice in a row, the second time in place
        buf, e1 = S.step_host(HEX, U0, dt)
        assert np.array_equal(buf, U_ref) and np.array_equal(e1, e_ref)
        assert np.array_equal(S.get_state(HEX), U_ref)            # the device state is the new state as well
        _, e2 = S.step_host(HEX, buf, dt, out=buf)
        assert np.array_equal(buf, U_ref2) and np.array_equal(e2, e_ref2)
        # and the library carries on from there
        e3 = S.stepSolver(dt, 1)                                   # from the resident (nodal) state
        S.set_state(HEX, U_ref2)                                   # from its modal image: equal up to the round-off of the transform
        assert np.allclose(S.stepSolver(dt, 1), e3, rtol=1e-9, atol=0.0)
    finally:
        del os.environ["SDG_HOST_PIPE_GROUPS"]


@pytest.mark.gpu
@pytest.mark.parametrize("dim,cells,cfg", [(3, 24, dict(p=2, conv_flux=2, rk=2)), (2, 128, dict(p=3, conv_flux=3, rk=2)), (2, 100, dict(p=4, conv_flux=2, rk=1, **NS)),
                                           (3, 22, dict(p=2, conv_flux=2, rk=2, **NS))])
def test_streamed_step_on_the_node_kernels(built, dim, cells, cfg):
    _with_groups(5)
    try:
        mesh = M.periodic_box_fast(dim, cells)
        S = Solver(cfg, mesh, device=0)
        t = S.types[0]
        S.initializeSolver(cases.ic_density_wave([0.5, 0.3, 0.2][:dim]))
        assert S.step_host_info()[0] == 5
        dt = S.calculateDeltaTime(1.0)
        U0 = S.get_state(t).copy()
        S.set_state(t, U0)
        e_ref = S.stepSolver(dt, 1).copy()
        U_ref = S.get_state(t).copy()
        buf, e1 = S.step_host(t, U0, dt)
        assert np.array_equal(buf, U_ref) and np.array_equal(e1, e_ref)
    finally:
        del os.environ["SDG_HOST_PIPE_GROUPS"]


@pytest.mark.gpu
def test_streamed_step_with_boundaries(built):
    """a box with far-field and wall faces (boundary faces carry no dependency), and a small mesh that runs the phases one after the other"""
    _with_groups(7)
    try:
        walls = {3: M.RIEMANN_FARFIELD, 4: M.RIEMANN_FARFIELD, 5: M.ADIABATIC_SLIP_WALL, 6: M.ADIABATIC_SLIP_WALL}
        ns_walls = {3: M.RIEMANN_FARFIELD, 4: M.ISOTHERMAL_NONSLIP_WALL, 5: M.ADIABATIC_NONSLIP_WALL, 6: M.ADIABATIC_SLIP_WALL}
        for cfg, mesh in [(dict(p=3, conv_flux=2, rk=2), M.box(3, (24, 20, 22), 0.0, 2.0, periodic_axes=(0,), phys_bc=walls)),
                          (dict(p=3, conv_flux=2, rk=2, **NS), M.box(3, (24, 20, 22), 0.0, 2.0, periodic_axes=(0,), phys_bc=ns_walls)),
                          (dict(p=3, conv_flux=2, rk=0, **dict(NS, visc_flux=1)), M.box(3, (22, 20, 24), 0.0, 2.0)),   # far field on all six sides, BR1
                          (dict(p=2, conv_flux=2, rk=2), M.periodic_box_fast(3, 12))]:   # 1,728 elements: no streaming
            S = Solver(cfg, mesh, device=0)
            S.initializeSolver(cases.ic_density_wave([0.5, 0.3, 0.2]), cases.bc_freestream(0.4, 0.0, 3, wall_phys=(5, 6), vel=[0.5, 0.3, 0.2]) if mesh.faces["n_bnd"] else None)
            dt = S.calculateDeltaTime(0.5)
            U0 = S.get_state(HEX).copy()
            S.set_state(HEX, U0)
            e_ref = S.stepSolver(dt, 1).copy()
            U_ref = S.get_state(HEX).copy()
            buf, e1 = S.step_host(HEX, U0, dt)
            assert np.isfinite(U_ref).all() and np.isfinite(e_ref).all()
            assert np.array_equal(buf, U_ref) and np.array_equal(e1, e_ref)
    finally:
        del os.environ["SDG_HOST_PIPE_GROUPS"]
