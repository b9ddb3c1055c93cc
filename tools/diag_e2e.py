"""Where the end-to-end step spends its time (128^3 P3 hexes): raw pinned copies, sdg_set_state, sdg_step, sdg_get_state."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from subrosadg_b200 import mesh as M
from subrosadg_b200.solver import Solver

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 128
mesh = M.periodic_box_fast(3, cells)
S = Solver(dict(p=3, conv_flux=2, rk=2), mesh, device=0)
t = S.types[0]
sz = S.sizes(t)
U = torch.empty((sz.n, sz.Nb, sz.Nv), dtype=torch.float64).pin_memory()
Un = U.numpy()
def ic(x):
    rho = 1.0 + 0.2 * np.sin(np.pi * x.sum(axis=-1))
    return np.stack([rho, np.full_like(rho, 0.5), np.full_like(rho, 0.3), np.full_like(rho, 0.2), 1.4 / rho], axis=-1)


S.initializeSolver(ic)
Un[...] = S.get_state(t)
d = torch.empty_like(U, device="cuda")
for name, fn in (("raw H2D", lambda: d.copy_(U, non_blocking=True)), ("raw D2H", lambda: U.copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"{name}: {U.numel() * 8 / dt / 1e9:.1f} GB/s ({dt * 1e3:.1f} ms)")
del d
dt_ = S.calculateDeltaTime(1.0)
for rep in range(3):
    t0 = time.perf_counter(); S.set_state(t, Un); t1 = time.perf_counter(); S.stepSolver(dt_, 1); t2 = time.perf_counter(); S.get_state(t, out=Un); t3 = time.perf_counter()
    print(f"set_state {1e3 * (t1 - t0):.1f} ms  step {1e3 * (t2 - t1):.1f} ms  get_state {1e3 * (t3 - t2):.1f} ms")
