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
    relerr = S.stepSolver(dt, steps)
    state = {t: S.get_state(t) for t in S.types}
    _check(state, initial, dt, relerr, gold, S.types, {t: initial[t].shape for t in S.types}, 1e-10, _projection_tolerance(cfg, mesh))   # 1e-10: BASELINE.json, fields after N steps
    if gold["node_artificial_viscosity"] is not None:
        _check_viscosity(S.node_artificial_viscosity(), gold)
