#!/usr/bin/env python
"""Throughput of the dense-operator path on a config-3 style hybrid mesh (karmanvortex_2d_cns: P3 triangles + curved P3 quads,
HLLC, BR2, Sutherland) — a side measurement for profiles/, not the bench.py contract (that is the hexahedral tensor path).
usage: python tools/bench_hybrid.py [scale=4.0] [steps=20]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from subrosadg_b200 import mesh as M
from subrosadg_b200.solver import Solver

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
mesh = M.annulus(int(22 * scale), 4 * int(15 * scale), r0=0.5, r1=20.0, geom_order=3, stretch=1.3, tri_rings=int(11 * scale),
                 phys_bc={1: M.RIEMANN_FARFIELD, 2: M.ADIABATIC_NONSLIP_WALL})
cfg = dict(p=3, model=1, transport=2, mu=1.4 * 0.2 / 200.0, conv_flux=2, visc_flux=2, rk=2)
S = Solver(cfg, mesh, device=0)
one = lambda x: np.ones(x.shape[:-1])


def ic(x):   # smooth start (no-slip already satisfied at the cylinder): the impulsive start of the example is too stiff for a timing run
    f = np.tanh((np.hypot(x[..., 0], x[..., 1]) - 0.5) / 0.5)
    return np.stack([1.4 * one(x), 0.2 * f, 0 * one(x), one(x)], axis=-1)


S.initializeSolver(ic, lambda x, phys, time=None: np.stack([1.4 * one(x), np.where(phys == 2, 0.0, 0.2) * one(x), 0 * one(x), one(x)], axis=-1))
dt = 0.1 * S.calculateDeltaTime(1.0)   # the convective CFL formula of the reference ignores the viscous limit of the wall cells
S.step_timed(dt, 3)
l0 = S.launch_count
err, ms = S.step_timed(dt, steps)
dof = sum(S.sizes(t).n * S.sizes(t).Nb * S.sizes(t).Nv for t in S.types)
print(json.dumps({"workload": "karmanvortex_2d_cns style hybrid mesh, P3, HLLC, BR2, Sutherland (dense-operator path)",
                  "elements": {int(t): S.sizes(t).n for t in S.types}, "scalar_dof": dof, "steps": steps, "ms_per_stage": ms / (3 * steps),
                  "MDOF_stage_per_s": dof * 3 * steps / (ms * 1e-3) / 1e6, "launches_per_stage": (S.launch_count - l0 - 1) / (3 * steps),
                  "relative_error": [float(e) for e in err]}))
