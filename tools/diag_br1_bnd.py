"""BR1 / BR2 Navier-Stokes on small P3-hexahedron boxes with boundary faces: CUDA path against the oracle (NaN hunt)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import cases
from subrosadg_b200 import mesh as M
HEX = M.HEXAHEDRON
NS = dict(model=1, transport=1, mu=1.4 * 0.2 / 200.0)
bc = cases.bc_freestream(0.4, 0.0, 3, wall_phys=(5, 6), vel=[0.5, 0.3, 0.2])
for shape in [(4, 3, 3), (22, 20, 24)]:
    for visc in (1, 2):
        for rk in (0, 1, 2):
            cfg = dict(p=3, conv_flux=2, rk=rk, visc_flux=visc, **NS)
            mesh = M.box(3, shape, 0.0, 2.0)
            if shape[0] < 10:
                O, S = cases.make_pair(cfg, mesh, cases.ic_density_wave([0.5, 0.3, 0.2]), bc)
            else:
                from subrosadg_b200.solver import Solver
                O = None; S = Solver(cfg, mesh, device=0); S.initializeSolver(cases.ic_density_wave([0.5, 0.3, 0.2]), bc)
            dt = S.calculateDeltaTime(0.5)
            e = S.stepSolver(dt, 1)
            q = S.state_at_quadrature(HEX)
            msg = f"shape {shape} visc {visc} rk {rk}: dt {dt:.3e} err {e} finite {np.isfinite(q).all()} nan elems {int((~np.isfinite(q)).any(axis=(1,2)).sum())}"
            if O is not None:
                O.step(dt, 1)
                msg += f" vs oracle {cases.rel_l2(q, O.state_at_quadrature(HEX)):.2e} oracle finite {np.isfinite(O.state_at_quadrature(HEX)).all()}"
            print(msg, flush=True)
            S.close()
