"""Regression pins of the CPU oracle on scaled-down versions of the five BASELINE configurations: relative_error_ after a few
steps and a checksum of the conserved fields at the quadrature points.  They do not pin the oracle against the reference (nothing
can, see oracle/oracle.cpp) — they pin it against drift, so that a change to the checker is a visible event.
usage: python tests/golden/make_oracle_pins.py   (writes tests/golden/oracle_pins.json)"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402
import oracle  # noqa: E402
from subrosadg_b200 import mesh as M  # noqa: E402

NS = dict(model=1, transport=1, mu=1.4 * 0.2 / 200.0)


def configs():
    yield "periodic_2d_ceuler", dict(p=3, conv_flux=2, rk=2), M.periodic_box(2, 10), cases.ic_density_wave([0.7, 0.3]), None, 1e-3, 5
    yield "naca0012_2d_ceuler", dict(p=3, conv_flux=2, rk=2), M.naca0012(nr=6, nt=16), cases.ic_perturbed_freestream(0.63, 2.0, 2, amp=1e-3), cases.bc_freestream(0.63, 2.0, 2), None, 4
    yield "karmanvortex_2d_cns", dict(NS, p=3, visc_flux=2, transport=2), M.annulus(4, 12, r0=0.5, r1=4.0, geom_order=3, stretch=1.5, tri_rings=2, phys_bc={1: M.RIEMANN_FARFIELD, 2: M.ADIABATIC_NONSLIP_WALL}), \
        cases.ic_perturbed_freestream(0.2, 0.0, 2, amp=1e-3), cases.bc_freestream(0.2, 0.0, 2, wall_phys=(2,)), None, 3
    yield "periodic_3d_ceuler", dict(p=3, conv_flux=2, rk=2), M.periodic_box_fast(3, 4), cases.ic_density_wave([0.5, 0.3, 0.2]), None, 5e-4, 4
    yield "sphere_3d_cns", dict(NS, p=3, visc_flux=2), M.cubed_sphere_shell(2, 2, r0=0.5, r1=4.0, geom_order=3), cases.ic_perturbed_freestream(0.2, 0.0, 3, amp=1e-3), \
        cases.bc_freestream(0.2, 0.0, 3, wall_phys=(2,)), None, 2


def run():
    out = {}
    for name, cfg, mesh, ic, bc, dt, nsteps in configs():
        O = oracle.Oracle(dict(cfg), mesh)
        O.initialize(ic, bc)
        if dt is None:
            dt = 0.3 * O.compute_dt(1.0)
        err = O.step(dt, nsteps)
        q = np.concatenate([O.state_at_quadrature(t).ravel() for t in O.types])
        out[name] = dict(dt=float(dt), steps=nsteps, relative_error=[float(e) for e in err], state_sum=float(q.sum()), state_l2=float(np.sqrt((q * q).sum())),
                         elements={int(t): int(O.sizes(t).n) for t in O.types})
    return out


if __name__ == "__main__":
    with open(os.path.join(HERE, "oracle_pins.json"), "w") as f:
        json.dump(run(), f, indent=1)
    print("written")
