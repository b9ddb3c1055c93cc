"""The reference's OWN solver — Solver<SC>::initializeSolver, calculateDeltaTime and stepSolver with all eight sweeps of SpatialDiscrete.cpp,
the RK update and the relative error, compiled from /root/reference/src (oracle/ref_sweeps.cpp) — against the oracle (CPU) and against
the CUDA path (GPU).  tests/golden/reference_sweeps.json was written by tests/golden/make_reference_sweeps.py in the development
container; here only the committed numbers are read.  Eight control types: line / quadrangle / triangle / hybrid / hexahedron meshes,
Euler and Navier-Stokes (BR1, BR2; constant and Sutherland viscosity), Lax-Friedrichs / HLLC / Roe, ForwardEuler / HeunRK2 / SSPRK3,
affine, curved and periodic meshes, far-field and wall boundaries; five more with ShockCapturingEnum::ArtificialViscosity (the reference's
calculateArtificialViscosity and its eps * grad(U) flux terms; the inner radius of the elements is an input on both sides)."""
import json
import os
import sys

import numpy as np
import pytest

import cases

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_reference_sweeps as gen   # noqa: E402  (the case list and the analytic fields; nothing reference-side is touched on import)

GOLD = {c["name"]: c for c in json.load(open(os.path.join(HERE, "golden", "reference_sweeps.json")))["cases"]}
CASES = gen.cases()


def _projection_tolerance(cfg, mesh):
    """initializeSolver multiplies by (Phi^T Phi)^-1 (InitialCondition.cpp:100-102), whose condition number is cond(Phi)^2 (2e5 for P3
    hexahedra): two fp64 inversions of it (here: Gauss-Jordan in the stand-in Eigen vs the oracle's extended-precision one) agree to
    about eps * cond(Phi)^2, and so do the projected coefficients."""
    import oracle
    O = oracle.Oracle(dict(cfg), mesh)
    return {t: max(1e-12, 2e-16 * np.linalg.cond(O.table(t, 0)) ** 2) for t in O.types}


def _check(solver_state, solver_initial, dt, relerr, gold, types, shapes, tol_state, tol_ic):
    assert abs(dt - gold["dt"]) <= 1e-12 * gold["dt"], f"delta_time {dt} vs {gold['dt']}"
    for t in types:
        ref0 = np.asarray(gold["initial"][str(t)]).reshape(shapes[t])
        ref = np.asarray(gold["state"][str(t)]).reshape(shapes[t])
        e0, e1 = cases.rel_l2(solver_initial[t], ref0), cases.rel_l2(solver_state[t], ref)
        assert e0 < tol_ic[t], f"type {t}: initializeSolver projection rel-L2 {e0:.3e} (tolerance {tol_ic[t]:.1e})"
        assert e1 < tol_state, f"type {t}: coefficients after {gold['steps']} steps rel-L2 {e1:.3e}"
    assert np.allclose(relerr, gold["relative_error"], rtol=1e-8, atol=1e-300), f"relative_error_ {relerr} vs {gold['relative_error']}"


def _check_viscosity(node_av, gold):
    """Solver::node_artificial_viscosity_ of the last step (calculateArtificialViscosity, SpatialDiscrete.cpp:124-192)"""
    ref = np.asarray(gold["node_artificial_viscosity"])
    assert ref.max() > 0.0 and np.array_equal(ref == 0.0, node_av == 0.0), "different elements are flagged"
    assert cases.rel_l2(node_av, ref) < 1e-9, f"node_artificial_viscosity_ rel-L2 {cases.rel_l2(node_av, ref):.3e}"


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_oracle_reproduces_the_reference_solver(case):
    import oracle
    name, _, cfg, mesh, vel, amp, steps, cfl = case
    gold = GOLD[name]
    for t, b in mesh.blocks.items():
        assert abs(float(np.asarray(b["coords"]).sum()) - gold["mesh_checksum"][str(t)]) < 1e-9, "the mesh producer changed: regenerate the golden file"
    ic, bc = gen.fields(mesh.dim, vel, amp, cfg)
    O = oracle.Oracle(dict(cfg, accurate=0), mesh)      # the reference's plain-double M^-1 apply
    O.initialize(ic, bc)
    initial = {t: O.get_state(t) for t in O.types}
    dt = O.compute_dt(cfl)
    if gen.time_varying(vel):   # Solver::updateBoundaryVariable at t = iteration_ * delta_time_ before every step (first step: t = 0)
        for i in range(steps):
            O.update_boundary(bc, i * dt)
            relerr = O.step(dt, 1)
    else:
        relerr = O.step(dt, steps)
    state = {t: O.get_state(t) for t in O.types}
    tol_ic = _projection_tolerance(cfg, mesh)
    _check(state, initial, dt, relerr, gold, O.types, {t: initial[t].shape for t in O.types}, max(1e-12, max(tol_ic.values())), tol_ic)
    if gold["node_artificial_viscosity"] is not None:
        _check_viscosity(O.node_artificial_viscosity(), gold)


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_cuda_path_reproduces_the_reference_solver(built, case):
    from subrosadg_b200.solver import Solver
    name, _, cfg, mesh, vel, amp, steps, cfl = case
    gold = GOLD[name]
    ic, bc = gen.fields(mesh.dim, vel, amp, cfg)
    S = Solver(dict(cfg), mesh, device=0)
    S.initializeSolver(ic, bc)
    initial = {t: S.get_state(t) for t in S.types}
    dt = S.calculateDeltaTime(cfl)
    if gen.time_varying(vel):
        for i in range(steps):
            S.updateBoundaryVariable(bc, i * dt)
            relerr = S.stepSolver(dt, 1)
    else:
        relerr = S.stepSolver(dt, steps)
    state = {t: S.get_state(t) for t in S.types}
    _check(state, initial, dt, relerr, gold, S.types, {t: initial[t].shape for t in S.types}, 1e-10, _projection_tolerance(cfg, mesh))   # 1e-10: BASELINE.json, fields after N steps
    if gold["node_artificial_viscosity"] is not None:
        _check_viscosity(S.node_artificial_viscosity(), gold)


# ---- raw files written by the reference's own Solver::writeRawBinary + the system's libzstd (RawBinary.cpp:42-57,75-191) ----------------------
RAW_CASES = [c for c in CASES if c[0] in gen.RAW_FILES]


def _reference_raw_file(case, sizes):
    """(capacity header, state, gradient blocks, boundary rows, viscosity tail) of tests/golden/reference_raw_<name>.zst, decoded by the
    restated reader of tests/test_raw_binary.py; the modal state in the file must be the golden state to the last bit."""
    from test_raw_binary import compress_bound, parse_payload, read_raw_binary
    name, _, cfg, mesh, *_ = case
    cap, buf = read_raw_binary(os.path.join(HERE, "golden", f"reference_raw_{name}.zst"))
    assert cap == compress_bound(len(buf))           # the header the reference writes: ZSTD_compressBound(payload size)
    ns = cfg.get("model", 0) in (1, 3)
    U, G, bnd, av = parse_payload(buf, mesh, {t: sizes[t].Nb for t in sizes}, ns)
    gold = GOLD[name]
    for t in sizes:
        assert np.array_equal(U[t].ravel(), np.asarray(gold["state"][str(t)])), "the file and the golden state come from the same reference run"
    if gold["node_artificial_viscosity"] is None:
        assert np.all(av == 0.0)
    else:
        assert np.array_equal(av, np.asarray(gold["node_artificial_viscosity"]))
    return U, G, bnd, av, ns


def _check_raw(file, state, gradient, boundary_gradient, node_av, types, tol):
    """what a solver would write (its state, gradient coefficients, boundary-parent rows, node viscosity) against the reference-written file"""
    U, G, bnd, av, ns = file
    for t in types:
        assert cases.rel_l2(state[t], U[t]) < tol
        if ns:
            assert cases.rel_l2(gradient[t], G[t]) < 10 * tol, f"type {t}: variable_gradient_basis_function_coefficient_ {cases.rel_l2(gradient[t], G[t]):.3e}"
    at = 0
    for (t, e, lf, u_row, g_row) in bnd:
        assert cases.rel_l2(state[t][e], u_row) < tol
        if ns:
            mine = boundary_gradient[at:at + g_row.size]
            at += g_row.size
            scale = max(np.abs(G[t]).max(), 1e-300)      # a row can be all but zero (uniform flow next to a far-field face)
            assert np.abs(mine - g_row).max() < 10 * tol * scale, f"boundary parent {e} face {lf}: {np.abs(mine - g_row).max() / scale:.3e}"
    if node_av is not None:
        assert cases.rel_l2(node_av, av) < 1e-9 if av.max() > 0 else np.all(node_av == 0.0)


@pytest.mark.parametrize("case", RAW_CASES, ids=[c[0] for c in RAW_CASES])
def test_oracle_matches_the_reference_written_raw_file(case):
    import oracle
    name, _, cfg, mesh, vel, amp, steps, cfl = case
    ic, bc = gen.fields(mesh.dim, vel, amp, cfg)
    O = oracle.Oracle(dict(cfg, accurate=0), mesh)
    O.initialize(ic, bc)
    O.step(O.compute_dt(cfl), steps)
    file = _reference_raw_file(case, {t: O.sizes(t) for t in O.types})
    ns = file[4]
    _check_raw(file, {t: O.get_state(t) for t in O.types}, {t: O.gradient_state(t) for t in O.types} if ns else None,
               O.boundary_gradient_state() if ns else None, O.node_artificial_viscosity() if cfg.get("av_tolerance") is not None else None, O.types, 1e-11)


@pytest.mark.gpu
@pytest.mark.parametrize("case", RAW_CASES, ids=[c[0] for c in RAW_CASES])
def test_cuda_path_matches_the_reference_written_raw_file(built, case):
    """the arrays Solver::writeRawBinary of the C++ host side streams out (sdg_get_state, sdg_get_gradient_state,
    sdg_get_boundary_gradient_state, sdg_get_node_artificial_viscosity) against the file the reference wrote for the same run"""
    from subrosadg_b200.solver import Solver
    name, _, cfg, mesh, vel, amp, steps, cfl = case
    ic, bc = gen.fields(mesh.dim, vel, amp, cfg)
    S = Solver(dict(cfg), mesh, device=0)
    S.initializeSolver(ic, bc)
    S.stepSolver(S.calculateDeltaTime(cfl), steps)
    file = _reference_raw_file(case, {t: S.sizes(t) for t in S.types})
    ns = file[4]
    _check_raw(file, {t: S.get_state(t) for t in S.types}, {t: S.gradient_state(t) for t in S.types} if ns else None,
               S.boundary_gradient_state() if ns else None, S.node_artificial_viscosity() if cfg.get("av_tolerance") is not None else None, S.types, 1e-10)
