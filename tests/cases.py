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


def ic_perturbed_freestream(mach, alpha_deg, dim, amp=1e-2, vel=None):
    """Uniform far-field state (examples/naca0012_2d_ceuler.cpp:29-33 style: rho=1.4, |u|=M, T=1) times a smooth
    perturbation so that residuals are non-trivial.  `vel` overrides the velocity vector: on a box-shaped far
    field a velocity parallel to a boundary face puts the Riemann far-field condition exactly ON its inflow / outflow switch
    (BoundaryCondition.cpp:82-285 branches on the sign of u.n), where parity between two fp64 implementations is undefined."""
    a = np.deg2rad(alpha_deg)
    vel = list(vel) if vel is not None else [mach * np.cos(a), mach * np.sin(a)] + ([0.0] if dim == 3 else [])

    def f(x):
        s = np.sin(np.pi * x[..., 0]) * np.cos(np.pi * x[..., 1])
        if dim == 3:
            s = s * np.cos(np.pi * x[..., 2])
        g = 1.0 + amp * s
        cols = [1.4 * g] + [v * g + 0.0 * s for v in vel] + [1.0 * g]
        return np.stack(cols, axis=-1)
    return f


def bc_freestream(mach, alpha_deg, dim, wall_phys=(2,), vel=None):
    a = np.deg2rad(alpha_deg)
    vel = list(vel) if vel is not None else [mach * np.cos(a), mach * np.sin(a)] + ([0.0] if dim == 3 else [])

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


def residual_sensitivity(cfg, mesh, ic, bc=None):
    """Conditioning probe for the parity tolerance: rel-L2 change of the ORACLE's own modal residual when every mesh coordinate moves by
    one unit round-off (deterministic +-1 ulp pattern, identical for the copies of a shared node) and so does every modal coefficient.  Two fp64 implementations that
    derive their metric terms from the same coordinates with different arithmetic (the oracle: Lagrange-derivative sums like gmsh's
    getJacobian; the product: host_plan.hpp) cannot agree better than this: coordinates of size |x| on cells of size h leave a relative
    error eps |x| / h in every metric term, and near-uniform flows amplify it again by |F| / (h |dF/dx|) through the cancellation of
    the constant part of the flux between the volume and the face integrals."""
    import copy
    import oracle
    m2 = copy.copy(mesh)
    m2.blocks = {}
    for t, b in mesh.blocks.items():
        x = np.asarray(b["coords"], dtype=np.float64)
        key = np.rint(x * 2.0 ** 20).astype(np.int64)               # same key for every copy of a shared node
        sgn = np.where(((key * 2654435761) >> 7) & 1, 1.0, -1.0)
        b2 = dict(b); b2["coords"] = x * (1.0 + sgn * np.finfo(np.float64).eps)
        m2.blocks[t] = b2
    A = oracle.Oracle(dict(cfg), mesh); A.initialize(ic, bc)
    B = oracle.Oracle(dict(cfg), m2); B.initialize(ic, bc)
    for t in A.types:   # ... and every modal coefficient by one unit round-off as well (conditioning with respect to the state)
        U = A.get_state(t)
        flip = np.where((np.arange(U.size).reshape(U.shape) * 2654435761 >> 5) & 1, 1.0, -1.0)
        B.set_state(t, U * (1.0 + flip * np.finfo(np.float64).eps))
    Ra, Rb = A.residual(), B.residual()
    a = np.concatenate([np.ravel(Ra[t][0]) for t in A.types]); b = np.concatenate([np.ravel(Rb[t][0]) for t in A.types])
    return rel_l2(b, a)


def conditioning(cfg, mesh, ic, bc=None, base=1e-12, safety=8.0):
    """Factor >= 1 by which BASELINE.json's residual tolerance (1e-12) has to be loosened for this set-up: `safety` unit round-offs of
    backward error in the inputs (coordinates, coefficients) times the oracle's measured sensitivity to one."""
    return max(1.0, safety * residual_sensitivity(cfg, mesh, ic, bc) / base)
