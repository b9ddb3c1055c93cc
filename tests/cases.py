"""Shared test cases: meshes, initial / boundary conditions of the five BASELINE configurations (scaled down) and
helpers to run the CPU oracle (oracle/, test infrastructure) next to the CUDA path (subrosadg_b200, the product)."""
import numpy as np

from subrosadg_b200 import mesh as M

GAMMA = 1.4


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    den = np.sqrt(np.sum(b * b))
    return float(np.sqrt(np.sum((a - b) ** 2)) / (den if den > 0 else 1.0))


def ic_density_wave(u):
    """examples/periodic_{2,3}d_ceuler.cpp:19-35: rho = 1 + 0.2 sin(pi sum x), constant velocity, p = 1 (T = 1.4/rho)."""
    u = np.asarray(u, dtype=float)

    def f(x):
        rho = 1.0 + 0.2 * np.sin(np.pi * x.sum(axis=-1))
        cols = [rho] + [np.full_like(rho, ui) for ui in u] + [1.4 / rho]
        return np.stack(cols, axis=-1)
    return f


def ic_perturbed_freestream(mach, alpha_deg, dim, amp=1e-2):
    """Uniform far-field state (examples/naca0012_2d_ceuler.cpp:29-33 style: rho=1.4, |u|=M, T=1) times a smooth
    perturbation so that residuals are non-trivial."""
    a = np.deg2rad(alpha_deg)
    vel = [mach * np.cos(a), mach * np.sin(a)] + ([0.0] if dim == 3 else [])

    def f(x):
        s = np.sin(np.pi * x[..., 0]) * np.cos(np.pi * x[..., 1])
        if dim == 3:
            s = s * np.cos(np.pi * x[..., 2])
        g = 1.0 + amp * s
        cols = [1.4 * g] + [v * g + 0.0 * s for v in vel] + [1.0 * g]
        return np.stack(cols, axis=-1)
    return f


def bc_freestream(mach, alpha_deg, dim, wall_phys=(2,)):
    a = np.deg2rad(alpha_deg)
    vel = [mach * np.cos(a), mach * np.sin(a)] + ([0.0] if dim == 3 else [])

    def f(x, phys, time=None):
        rho = np.full(x.shape[:-1], 1.4)
        wall = np.isin(phys, wall_phys)
        cols = [rho] + [np.where(wall, 0.0, v) + 0.0 * rho for v in vel] + [np.ones_like(rho)]
        return np.stack(cols, axis=-1)
    return f


def make_pair(cfg, mesh, ic, bc=None, threads=None):
    """(oracle, product) initialised identically."""
    import oracle
    from subrosadg_b200.solver import Solver
    O = oracle.Oracle(dict(cfg), mesh, threads=threads)
    S = Solver(dict(cfg), mesh, device=0)
    O.initialize(ic, bc)
    S.initializeSolver(ic, bc)
    return O, S
