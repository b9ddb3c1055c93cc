"""Floating-point pins derived from the REFERENCE ITSELF: tests/golden/reference_physics.json holds inputs and outputs of the reference's
own pointwise physics (src/Solver/{VariableConvertor,ConvectiveFlux,ViscousFlux,BoundaryCondition,PhysicalModel,SourceTerm}.cpp compiled
by `make -C oracle ref`, generator tests/golden/make_reference_physics.py): Riemann fluxes (Central / Lax-Friedrichs / HLLC / Roe / Exact),
the six boundary conditions with their gradient states and modifyBoundaryVariable, primitive gradients, viscous fluxes (constant and
Sutherland transport), raw fluxes, variable conversions and the Boussinesq source, in 1, 2 and 3 dimensions.

CPU: the oracle's restatement (oracle/physics.hpp) must reproduce every vector to round-off (<= 1e-13 relative to the row's magnitude;
the reference arithmetic is Eigen expression order without FMA contraction, the oracle's is -march=x86-64-v3 with contraction).
GPU: the CUDA device functions of the product (physics.cuh) through sdg_debug_physics, <= 1e-12 (hardware rcp / rsqrt seeds + Newton)."""
import ctypes
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "reference_physics.json")) as f:
    GOLD = json.load(f)
PKEYS = ("cp", "cv", "mu", "c0", "rho0", "beta", "t_ref")
PARAMS = np.array([GOLD["params"][k] for k in PKEYS])
IDS = [f"{c['name']}-{c['dim']}d-what{c['what']}-bc{c['bc']}" for c in GOLD["cases"]]


def row_error(got, ref):
    """max over rows of |got - ref| / max |ref| of the row (a row = all outputs of one point: they share one magnitude)"""
    ref = np.asarray(ref); got = np.asarray(got)
    scale = np.maximum(np.abs(ref).max(axis=1, keepdims=True), 1e-300)
    return float((np.abs(got - ref) / scale).max())


def evaluate(fn, case):
    cfg = np.array(case["cfg"], dtype=np.int32)
    inp = np.ascontiguousarray(case["input"], dtype=np.float64)
    ref = np.asarray(case["output"])
    out = np.zeros_like(ref)
    dp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    rc = fn(cfg.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), dp(PARAMS), int(case["what"]), int(case["bc"]), inp.shape[0], dp(inp), dp(out))
    return rc, out, ref


def test_golden_file_covers_the_physics_rows():
    what = {(c["what"], c["cfg"][1], c["cfg"][4]) for c in GOLD["cases"]}
    for conv in (0, 1, 2, 3):
        assert (0, 0, conv) in what                      # P2-P4: Central, Lax-Friedrichs, HLLC, Roe on the ideal gas
    assert (0, 2, 4) in what                             # Exact flux on the weakly compressible fluid
    assert {c["bc"] for c in GOLD["cases"] if c["what"] == 1} == set(range(6))   # P7: all six BoundaryConditionImpl
    assert any(c["what"] == 2 and c["cfg"][3] == 2 for c in GOLD["cases"])         # P5: Sutherland
    assert any(c["what"] == 3 and c["cfg"][5] == 1 for c in GOLD["cases"])         # P8: Boussinesq
    assert {c["dim"] for c in GOLD["cases"]} == {1, 2, 3}
    view = [c for c in GOLD["cases"] if c["what"] == 4]              # ViewVariable::get: Euler and NS, ideal gas and weakly compressible, 1-3 D
    assert {c["cfg"][1] for c in view} == {0, 1, 2, 3} and {c["dim"] for c in view} == {1, 2, 3}


@pytest.mark.parametrize("case", GOLD["cases"], ids=IDS)
def test_oracle_reproduces_the_reference(built, case):
    import oracle
    rc, out, ref = evaluate(oracle.lib().orc_physics, case)
    assert rc == 0, oracle.lib().orc_last_error()
    assert row_error(out, ref) < 1e-13, (case["name"], row_error(out, ref))


@pytest.mark.gpu
@pytest.mark.parametrize("case", GOLD["cases"], ids=IDS)
def test_cuda_device_functions_reproduce_the_reference(built, case):
    from subrosadg_b200 import solver
    lib = solver.load_library()
    rc, out, ref = evaluate(lib.sdg_debug_physics, case)
    assert rc == 0, lib.sdg_last_error()
    assert row_error(out, ref) < 1e-12, (case["name"], row_error(out, ref))
