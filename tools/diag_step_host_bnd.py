"""sdg_step_host against the phases on Navier-Stokes boxes with boundary faces: size and place of a mismatch, by number of upload groups"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import cases
from subrosadg_b200 import mesh as M
from subrosadg_b200.solver import Solver
HEX = M.HEXAHEDRON
NS = dict(model=1, transport=1, mu=1.4 * 0.2 / 200.0, visc_flux=2)
ns_walls = {3: M.RIEMANN_FARFIELD, 4: M.ISOTHERMAL_NONSLIP_WALL, 5: M.ADIABATIC_NONSLIP_WALL, 6: M.ADIABATIC_SLIP_WALL}
runs = [("BR2 ssprk3 periodic-x", dict(p=3, conv_flux=2, rk=2, **NS), lambda: M.box(3, (24, 20, 22), 0.0, 2.0, periodic_axes=(0,), phys_bc=ns_walls)),
        ("BR1 heun all-bnd", dict(p=3, conv_flux=2, rk=1, **dict(NS, visc_flux=1)), lambda: M.box(3, (22, 20, 24), 0.0, 2.0, phys_bc={k: ns_walls.get(k, M.RIEMANN_FARFIELD) for k in range(1, 7)})),
        ("BR2 fe farfield only", dict(p=3, conv_flux=2, rk=0, **NS), lambda: M.box(3, (22, 20, 24), 0.0, 2.0))]
for G in (1, 2, 7):
    os.environ["SDG_HOST_PIPE_GROUPS"] = str(G)
    for name, cfg, mk in runs:
        mesh = mk()
        S = Solver(cfg, mesh, device=0)
        S.initializeSolver(cases.ic_density_wave([0.5, 0.3, 0.2]), cases.bc_freestream(0.4, 0.0, 3, wall_phys=(5, 6), vel=[0.5, 0.3, 0.2]))
        dt = S.calculateDeltaTime(0.5)
        U0 = S.get_state(HEX).copy()
        S.set_state(HEX, U0); e_ref = S.stepSolver(dt, 1).copy(); U_ref = S.get_state(HEX).copy()
        buf, e1 = S.step_host(HEX, U0, dt)
        d = np.abs(buf - U_ref)
        bad = np.argwhere(d.max(axis=(1, 2)) > 0).ravel()
        print(f"G={G} {name}: groups {S.step_host_info()} max diff {d.max():.3e} bad elements {bad.size} of {d.shape[0]} first {bad[:8]} err diff {np.abs(e1 - e_ref).max():.3e}", flush=True)
        S.close()
