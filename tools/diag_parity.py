"""Prints the oracle-vs-CUDA error table (no assertions): used to calibrate tolerances and to find roundoff sources."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases
from subrosadg_b200 import mesh as M


def row(label, cfg, mesh, ic, bc=None, dt=1e-3, n=5, same_input=True):
    O, S = cases.make_pair(cfg, mesh, ic, bc)
    t = S.types[0]
    e0 = cases.rel_l2(S.get_state(t), O.get_state(t))
    if same_input:
        S.set_state(t, O.get_state(t))
    Ro, qo = O.residual()[t]; Rs, qs = S.residual()[t]
    eq, eR = cases.rel_l2(qs, qo), cases.rel_l2(Rs, Ro)
    eo = O.step(dt, n); es = S.stepSolver(dt, n)
    eu = cases.rel_l2(S.state_at_quadrature(t), O.state_at_quadrature(t))
    en = np.max(np.abs(es - eo) / np.maximum(np.abs(eo), 1e-300))
    print(f"{label:34s} ic {e0:.1e} dUdt {eq:.2e} Rmodal {eR:.2e} state{n} {eu:.2e} relerr {en:.1e}", flush=True)


for same in (False, True):
    print("same_input", same)
    for p in (1, 2, 3):
        row(f"periodic2d p{p}", dict(p=p), M.periodic_box(2, 10), cases.ic_density_wave([0.7, 0.3]), same_input=same)
    for p in (1, 2, 3):
        row(f"periodic3d p{p}", dict(p=p), M.periodic_box_fast(3, 6), cases.ic_density_wave([0.5, 0.3, 0.2]), same_input=same)
    for fl in (0, 1, 2, 3):
        row(f"flux{fl} p2", dict(p=2, conv_flux=fl), M.periodic_box(2, 6), cases.ic_density_wave([0.7, 0.3]), same_input=same)
    warp = lambda x: x + 0.04 * np.sin(np.pi * x[:, ::-1])
    row("curved box farfield", dict(p=3), M.box(2, (6, 5), 0.0, 1.0, geom_order=3, warp=warp), cases.ic_perturbed_freestream(0.63, 2.0, 2),
        cases.bc_freestream(0.63, 2.0, 2, wall_phys=()), same_input=same)
    row("naca0012", dict(p=3), M.naca0012(nr=8, nt=24), cases.ic_perturbed_freestream(0.63, 2.0, 2, amp=1e-3), cases.bc_freestream(0.63, 2.0, 2), dt=1e-4, same_input=same)
    warp3 = lambda x: x + 0.03 * np.sin(np.pi * np.roll(x, 1, axis=1))
    row("curved hex box roe", dict(p=2, conv_flux=3), M.box(3, (3, 4, 3), 0.0, 1.0, geom_order=2, warp=warp3), cases.ic_perturbed_freestream(0.5, 3.0, 3),
        cases.bc_freestream(0.5, 3.0, 3, wall_phys=()), same_input=same)
    row("sphere shell euler", dict(p=3), M.cubed_sphere_shell(3, 3), cases.ic_perturbed_freestream(0.3, 0.0, 3, amp=1e-3), cases.bc_freestream(0.3, 0.0, 3), dt=1e-4, same_input=same)
