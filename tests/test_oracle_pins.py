"""The oracle reproduces its committed regression pins on the five (scaled-down) BASELINE configurations (CPU), and the CUDA path
reproduces the same numbers (GPU).  tests/golden/oracle_pins.json is written by tests/golden/make_oracle_pins.py."""
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_oracle_pins as P  # noqa: E402

PINS = json.load(open(os.path.join(HERE, "golden", "oracle_pins.json")))


def test_oracle_matches_its_pins(built):
    got = P.run()
    assert sorted(got) == sorted(PINS) and len(PINS) == 5
    for name, pin in PINS.items():
        g = got[name]
        assert g["elements"] == {int(k): v for k, v in pin["elements"].items()}
        assert np.allclose(g["relative_error"], pin["relative_error"], rtol=1e-9), name
        assert abs(g["state_sum"] - pin["state_sum"]) <= 1e-10 * pin["state_l2"], name
        assert abs(g["state_l2"] / pin["state_l2"] - 1.0) < 1e-12, name


@pytest.mark.gpu
def test_cuda_path_matches_the_pins(built):
    from subrosadg_b200.solver import Solver
    for name, cfg, mesh, ic, bc, dt, nsteps in P.configs():
        pin = PINS[name]
        S = Solver(dict(cfg), mesh, device=0)
        S.initializeSolver(ic, bc)
        err = S.stepSolver(pin["dt"], nsteps)
        q = np.concatenate([S.state_at_quadrature(t).ravel() for t in S.types])
        assert np.allclose(err, pin["relative_error"], rtol=1e-7), (name, err, pin["relative_error"])
        assert abs(np.sqrt((q * q).sum()) / pin["state_l2"] - 1.0) < 1e-10, name
        assert abs(q.sum() - pin["state_sum"]) <= 1e-9 * pin["state_l2"], name
