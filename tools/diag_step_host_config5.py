"""sdg_step_host on BASELINE config 5 (sphere_3d_cns: 29,592 curved P3 hexahedra, Navier-Stokes BR2, far field + no-slip sphere): bit-identity
with the phases and the time of both on an unstructured block mesh with boundary faces (Morton order inside, caller order = block order)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
import cases
from subrosadg_b200 import mesh as M
from subrosadg_b200.solver import Solver
HEX = M.HEXAHEDRON
cfg = dict(p=3, model=1, transport=1, mu=1.4 * 0.2 / 200.0, conv_flux=2, visc_flux=2, rk=2)
vel = [0.2 * 0.9, 0.2 * 0.3, 0.2 * np.sqrt(1.0 - 0.81 - 0.09)]
mesh = M.sphere_in_box()
for G in (0, 4, 8):
    if G: os.environ["SDG_HOST_PIPE_GROUPS"] = str(G)
    S = Solver(cfg, mesh, device=0)
    S.initializeSolver(cases.ic_perturbed_freestream(0.2, 0.0, 3, amp=1e-3, vel=vel), cases.bc_freestream(0.2, 0.0, 3, wall_phys=(2,), vel=vel))
    dt = 0.3 * S.calculateDeltaTime(1.0)
    sz = S.sizes(HEX)
    U = torch.empty((sz.n, sz.Nb, sz.Nv), dtype=torch.float64).pin_memory().numpy()
    V = torch.empty((sz.n, sz.Nb, sz.Nv), dtype=torch.float64).pin_memory().numpy()
    U[...] = S.get_state(HEX); V[...] = U
    S.set_state(HEX, U); e_ref = S.stepSolver(dt, 1).copy(); ref = S.get_state(HEX).copy()
    out, e1 = S.step_host(HEX, V, dt, out=V)
    same = np.array_equal(out, ref) and np.array_equal(e1, e_ref) and bool(np.isfinite(ref).all())
    n = 20
    t0 = time.perf_counter()
    for _ in range(n):
        S.set_state(HEX, U); S.stepSolver(dt, 1); S.get_state(HEX, out=U)
    t_ph = (time.perf_counter() - t0) / n
    t0 = time.perf_counter()
    for _ in range(n):
        S.step_host(HEX, V, dt, out=V)
    t_st = (time.perf_counter() - t0) / n
    print(f"config 5, {sz.n} elements, groups {S.step_host_info()}: bit-identical {same}; phases {1e3 * t_ph:.2f} ms per step, streamed {1e3 * t_st:.2f} ms per step", flush=True)
    S.close()
