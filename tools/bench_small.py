#!/usr/bin/env python
"""Steps per second on the reference's small shipped configurations (launch-bound meshes): config 1 (periodic_2d_ceuler, 100 P3
quads) and config 2 (naca0012_2d_ceuler, 2,204 curved P3 quads), with the CUDA-graph replay of sdg_step and, for comparison,
with SDG_NO_GRAPH=1 (every launch issued from the host).  usage: python tools/bench_small.py [steps=3000]"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def run(steps):
    import numpy as np
    import cases
    from subrosadg_b200 import mesh as M
    from subrosadg_b200.solver import Solver
    out = {}
    for name, mesh, ic, bc in [
        ("periodic_2d_ceuler_100_quads", M.periodic_box(2, 10), cases.ic_density_wave([0.7, 0.3]), None),
        ("naca0012_2d_ceuler_2204_quads", M.naca0012(nr=19, nt=116), cases.ic_perturbed_freestream(0.63, 2.0, 2, amp=0.0), cases.bc_freestream(0.63, 2.0, 2)),
    ]:
        S = Solver(dict(p=3, conv_flux=2, rk=2), mesh, device=0)
        S.initializeSolver(ic, bc)
        dt = 0.2 * S.calculateDeltaTime(1.0)
        S.stepSolver(dt, 50)
        t0 = time.perf_counter()
        err = S.stepSolver(dt, steps)
        sec = time.perf_counter() - t0
        out[name] = {"elements": S.sizes(S.types[0]).n, "steps_per_s": steps / sec, "us_per_step": 1e6 * sec / steps, "finite": bool(np.all(np.isfinite(err)))}
    return out


if __name__ == "__main__":
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
    if os.environ.get("SDG_BENCH_SMALL_CHILD"):
        print(json.dumps(run(steps)))
    else:
        res = {}
        for label, env in (("graph", {}), ("no_graph", {"SDG_NO_GRAPH": "1"})):
            e = dict(os.environ, SDG_BENCH_SMALL_CHILD="1", **env)
            r = subprocess.run([sys.executable, __file__, str(steps)], env=e, capture_output=True, text=True)
            res[label] = json.loads(r.stdout.strip().splitlines()[-1]) if r.returncode == 0 else {"error": r.stderr[-500:]}
        print(json.dumps(res))
