"""Time line of one streamed host step (sdg_step_host) at 128^3 P3 hexahedra, Euler: when each upload group has arrived, when its
download is queued (last stage + transform done) and when it is back in host memory.  SDG_HOST_PIPE_TIMING=1 makes the library print it."""
import os
import sys
import time

os.environ["SDG_HOST_PIPE_TIMING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from subrosadg_b200 import mesh as M
from subrosadg_b200.solver import Solver

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 128
S = Solver(dict(bench.CFG), M.periodic_box_fast(3, cells), device=0)
S.initializeSolver(bench.ic_config4)
t = S.types[0]
sz = S.sizes(t)
dt = S.calculateDeltaTime(1.0)
U = torch.empty((sz.n, sz.Nb, sz.Nv), dtype=torch.float64).pin_memory().numpy()
U[...] = S.get_state(t)
S.step_host(t, U, dt, out=U)
sys.stderr.write("---- second call (everything allocated) ----\n")
t0 = time.perf_counter()
S.step_host(t, U, dt, out=U)
print(f"cells {cells}: groups {S.step_host_info()}, wall {1e3 * (time.perf_counter() - t0):.1f} ms")
