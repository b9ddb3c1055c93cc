#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun on the GPU box): DistributedSolver over NCCL against a single-context run
of the same mesh on rank 0.  Prints one line per case; exits non-zero on mismatch."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
from subrosadg_b200 import mesh as M  # noqa: E402
from subrosadg_b200.parallel import DistributedSolver  # noqa: E402
from subrosadg_b200.solver import Solver  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    ok = True
    NS = dict(model=1, transport=1, mu=0.01)
    for dim, n, cfg in [(3, 8, dict(p=3, conv_flux=2, rk=2)), (2, 16, dict(p=3, conv_flux=3, rk=2)), (3, 8, dict(NS, p=2, conv_flux=2, rk=2, visc_flux=2)),
                        (3, 6, dict(NS, p=3, conv_flux=2, rk=2, visc_flux=1)), (3, 8, dict(NS, p=3, conv_flux=2, rk=2, visc_flux=2))]:
        mesh = M.periodic_box_fast(dim, n)
        ic = cases.ic_density_wave([0.7, 0.3] if dim == 2 else [0.5, 0.3, 0.2])
        D = DistributedSolver(dict(cfg), mesh, device=local)
        D.initializeSolver(ic)
        dt = D.calculateDeltaTime(0.5)
        err = D.stepSolver(dt, 4)
        got = D.gather_state_at_quadrature()
        if rank == 0:
            S = Solver(dict(cfg), mesh, device=local)
            S.initializeSolver(ic)
            dt1 = S.calculateDeltaTime(0.5)
            err1 = S.stepSolver(dt1, 4)
            ref = S.state_at_quadrature(S.types[0])
            e = cases.rel_l2(got, ref)
            good = e < 1e-14 and dt == dt1 and np.allclose(err, err1, rtol=1e-11, atol=1e-300)
            ok = ok and good
            print(f"mgpu_check world={world} dim={dim} n={n} cfg={cfg}: state rel-L2 {e:.2e} dt {dt:.6e}/{dt1:.6e} relerr {err[0]:.6e}/{err1[0]:.6e} {'OK' if good else 'MISMATCH'}", flush=True)
        dist.barrier()
    # shock capturing across the ranks: the node maximum is all-reduced (ncclMax) after sdg_step_begin -- the one collective of the path
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_gpu_av import jump_ic
    far, slip = M.RIEMANN_FARFIELD, M.ADIABATIC_SLIP_WALL
    for dim, shape, p, width in [(2, (16, 8), 3, 0.04), (3, (8, 4, 3), 2, 0.08), (3, (8, 3, 3), 3, 0.04)]:
        mesh = M.box(dim, shape, 0.0, 1.0, phys_bc={1: far, 2: far, 3: slip, 4: slip, 5: slip, 6: slip} if dim == 3 else {1: far, 2: far, 3: slip, 4: slip})
        cfg = dict(p=p, conv_flux=2, rk=2, av_tolerance=1.0, av_factor=2.0)
        ic = jump_ic(dim, width=width)
        bc = lambda x, phys, time=None: ic(x)
        D = DistributedSolver(dict(cfg), mesh, device=local)
        D.initializeSolver(ic, bc)
        dt = 0.2 * D.calculateDeltaTime(1.0)
        err = D.stepSolver(dt, 3)
        got = D.gather_state_at_quadrature()
        node = D.S.node_artificial_viscosity()
        if rank == 0:
            S = Solver(dict(cfg), mesh, device=local)
            S.initializeSolver(ic, bc)
            dt1 = 0.2 * S.calculateDeltaTime(1.0)
            err1 = S.stepSolver(dt1, 3)
            ref = S.state_at_quadrature(S.types[0])
            node1 = S.node_artificial_viscosity()
            e = cases.rel_l2(got, ref)
            good = (e < 1e-14 and dt == dt1 and np.allclose(err, err1, rtol=1e-11, atol=1e-300) and node1.max() > 0 and np.array_equal(node == 0, node1 == 0)
                    and cases.rel_l2(node, node1) < 1e-12)
            ok = ok and good
            print(f"mgpu_check world={world} shock capturing dim={dim} shape={shape} p={p}: state rel-L2 {e:.2e} node viscosity rel-L2 {cases.rel_l2(node, node1):.2e} "
                  f"({int((node1 > 0).sum())} of {node1.size} nodes) relerr {err[0]:.6e}/{err1[0]:.6e} {'OK' if good else 'MISMATCH'}", flush=True)
        dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
