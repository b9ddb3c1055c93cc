"""GPU diagnostic: which elements / face-pair combinations of sphere_in_box disagree with the oracle."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases
from subrosadg_b200 import mesh as M

NS = dict(model=1, transport=1, mu=1.4 * 0.2 / 200.0)
mesh = M.sphere_in_box(3, 3, 2)
f = mesh.faces
ni = int(f["n_int"])
for name, cfg in [("euler p3", dict(p=3, conv_flux=2, rk=2)), ("euler p2", dict(p=2, conv_flux=2, rk=2)), ("ns br2 p3", dict(NS, p=3, visc_flux=2)), ("ns br1 p3", dict(NS, p=3, visc_flux=1)),
                  ("ns br2 p2", dict(NS, p=2, visc_flux=2))]:
    ic = cases.ic_perturbed_freestream(0.2, 0.0, 3, amp=1e-3)
    O, S = cases.make_pair(cfg, mesh, ic, cases.bc_freestream(0.2, 0.0, 3, wall_phys=(2,)))
    t = S.types[0]
    S.set_state(t, O.get_state(t))
    Ro, qo = O.residual()[t]
    Rs, qs = S.residual()[t]
    err = np.sqrt(((Rs - Ro) ** 2).sum(axis=(1, 2))) / np.sqrt((Ro ** 2).sum() / Ro.shape[0])
    bad = np.flatnonzero(err > 1e-9)
    print(f"== {name}: rel-L2 {cases.rel_l2(Rs, Ro):.3e}; {bad.size} of {err.size} elements off")
    if "ns" in name:
        go, gs = O.gradient_at_quadrature(t), S.gradient_at_quadrature(t)
        gerr = np.sqrt(((gs - go) ** 2).sum(axis=(1, 2))) / np.sqrt((go ** 2).sum() / go.shape[0])
        gbad = np.flatnonzero(gerr > 1e-9)
        print(f"   gradient rel-L2 {cases.rel_l2(gs, go):.3e}; {gbad.size} elements off")
        bad = np.union1d(bad, gbad)
    if bad.size:
        badset = set(bad.tolist())
        combos_bad, combos_all = {}, {}
        for i in range(len(f["le"])):
            key = (int(f["lf"][i]), int(f["rf"][i]), int(f["rot"][i]), int(f["bc"][i]) if i >= ni else -1)
            combos_all[key] = combos_all.get(key, 0) + 1
            touched = int(f["le"][i]) in badset or (i < ni and int(f["re"][i]) in badset)
            if touched:
                combos_bad[key] = combos_bad.get(key, 0) + 1
        for k in sorted(combos_all):
            print(f"   (lfL, lfR, rot, bc) {k}: {combos_bad.get(k, 0)} / {combos_all[k]} faces touch an off element")
        print("   worst elements:", bad[np.argsort(-err[bad])][:10], err[bad].max())
