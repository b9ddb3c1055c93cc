"""Generates tests/golden/reference_physics.json from the REFERENCE'S OWN pointwise physics.

oracle/Makefile (target `ref`) compiles /root/reference/src/Solver/{VariableConvertor,ConvectiveFlux,ViscousFlux,BoundaryCondition,
PhysicalModel,SourceTerm}.cpp — where they lie, behind oracle/ref_physics.cpp and the declaration-level stand-ins of oracle/ref_shim/ —
into oracle/_ref/libref_physics.so; this script feeds it a fixed point set and stores inputs and outputs.  Run it in the development
container (it needs /root/reference); the committed JSON is what travels:

    python tests/golden/make_reference_physics.py

Consumers: tests/test_reference_physics.py (CPU: the oracle's restatement against these vectors; GPU: the CUDA device functions)."""
import ctypes
import json
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
PARAMS = dict(cp=2.5, cv=25.0 / 14.0, mu=1.4 * 0.2 / 200.0, c0=10.0, rho0=1.0, beta=0.5, t_ref=1.0)
PKEYS = ("cp", "cv", "mu", "c0", "rho0", "beta", "t_ref")

# (name, model, eos, transport, conv) — src/Utils/Enum.cpp values
EULER = [("ceuler_central", 0, 0, 0, 0), ("ceuler_lf", 0, 0, 0, 1), ("ceuler_hllc", 0, 0, 0, 2), ("ceuler_roe", 0, 0, 0, 3),
         ("ieuler_central", 2, 1, 0, 0), ("ieuler_lf", 2, 1, 0, 1), ("ieuler_exact", 2, 1, 0, 4)]
NS = [("cns_constant", 1, 0, 1, 2), ("cns_sutherland", 1, 0, 2, 2), ("ins_constant", 3, 1, 1, 4)]


def unit_normals(rng, n, dim):
    v = rng.normal(size=(n, dim))
    v[0] = np.eye(dim)[0]                     # an axis-aligned normal as well
    return v / np.linalg.norm(v, axis=1, keepdims=True)


def states(rng, n, dim, compressible, mach):
    """conserved states with the given normal-ish Mach numbers (compressible: ideal gas, gamma = 1.4; else weakly compressible c0 = 10)"""
    rho = 1.0 + 0.4 * rng.random(n) if compressible else 1.0 + 0.02 * rng.random(n)
    T = 0.8 + 0.5 * rng.random(n)
    c = np.sqrt(1.4 * 0.4 * PARAMS["cv"] * T) if compressible else np.full(n, PARAMS["c0"])
    direction = rng.normal(size=(n, dim)); direction /= np.linalg.norm(direction, axis=1, keepdims=True)
    vel = direction * (np.asarray(mach)[:n, None] * c[:, None])
    e = PARAMS["cv"] * T
    E = rho * (e + 0.5 * (vel ** 2).sum(axis=1)) if compressible else rho * e
    return np.concatenate([rho[:, None], rho[:, None] * vel, E[:, None]], axis=1), np.concatenate([rho[:, None], vel, T[:, None]], axis=1)


def main():
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], check=True)
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_physics.so"))
    dp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    params = np.array([PARAMS[k] for k in PKEYS])
    cases = []

    def call(dim, model, eos, transport, conv, source, what, bc, inp, n_out):
        cfg = np.array([dim, model, eos, transport, conv, source], dtype=np.int32)
        inp = np.ascontiguousarray(inp, dtype=np.float64)
        out = np.zeros((inp.shape[0], n_out))
        rc = lib.ref_physics(cfg.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), dp(params), what, bc, inp.shape[0], dp(inp), dp(out))
        assert rc == 0, (dim, model, eos, transport, conv, what, bc, rc)
        assert np.all(np.isfinite(out))
        return inp, out

    rng = np.random.default_rng(20261017)
    rng4 = np.random.default_rng(20261018)      # its own stream: the cases above keep their inputs
    n = 8
    mach = np.array([0.1, 0.5, 0.9, 1.6, 2.5, 0.3, 0.7, 0.05])
    for dim in (1, 2, 3):
        nv, nc = dim + 2, dim + 3
        for name, model, eos, transport, conv in EULER + NS:
            comp = eos == 0
            m = mach if comp else mach * 0.02
            nrm = unit_normals(rng, n, dim)
            L, _ = states(rng, n, dim, comp, m)
            R, primR = states(rng, n, dim, comp, m[::-1])
            # what 0: Riemann flux; the first half of the points use nearly equal states (smooth flow), the second half strong jumps, and two
            # points are aligned with the normal so that the supersonic early returns of HLLC (S_L >= 0, S_R <= 0) are hit in both directions
            Rw = R.copy(); Rw[: n // 2] = L[: n // 2] * (1.0 + 1e-3 * rng.normal(size=(n // 2, nv)))
            if comp:
                for k, sgn in ((3, 1.0), (4, -1.0)):
                    speed = np.linalg.norm(L[k, 1:1 + dim] / L[k, 0])
                    for S in (L, Rw):
                        rho = S[k, 0]; e_int = S[k, -1] / rho - 0.5 * np.sum((S[k, 1:1 + dim] / rho) ** 2)
                        S[k, 1:1 + dim] = rho * sgn * speed * nrm[k]
                        S[k, -1] = rho * (e_int + 0.5 * speed ** 2)
            inp, out = call(dim, model, eos, transport, conv, 0, 0, 0, np.concatenate([nrm, L, Rw], axis=1), nv)
            cases.append(dict(name=name, dim=dim, cfg=[dim, model, eos, transport, conv, 0], what=0, bc=0, input=inp.tolist(), output=out.tolist()))
            # what 1: the six boundary conditions; interior velocity mostly along +-normal so that sub- / supersonic in- / outflow all occur
            if conv in (2, 4):
                grad = 0.3 * rng.normal(size=(n, nv * dim))
                Lb = L.copy()
                sign = np.where(np.arange(n) % 2 == 0, 1.0, -1.0)
                speed = np.linalg.norm(L[:, 1:1 + dim] / L[:, :1], axis=1)
                vel = (sign * speed)[:, None] * nrm + 0.1 * speed[:, None] * rng.normal(size=(n, dim))
                if comp:
                    e_int = L[:, -1] / L[:, 0] - 0.5 * (L[:, 1:1 + dim] ** 2).sum(axis=1) / L[:, 0] ** 2
                    Lb[:, -1] = L[:, 0] * (e_int + 0.5 * (vel ** 2).sum(axis=1))
                Lb[:, 1:1 + dim] = L[:, :1] * vel
                for bc in range(6):
                    n_out = nc + 3 * nv + ((nc + nv) if model in (1, 3) else 0)
                    inp, out = call(dim, model, eos, transport, conv, 0, 1, bc, np.concatenate([nrm, Lb, primR, grad], axis=1), n_out)
                    cases.append(dict(name=name, dim=dim, cfg=[dim, model, eos, transport, conv, 0], what=1, bc=bc, input=inp.tolist(), output=out.tolist()))
            # what 2: viscous terms
            if model in (1, 3):
                grad = 0.5 * rng.normal(size=(n, nv * dim))
                inp, out = call(dim, model, eos, transport, conv, 0, 2, 0, np.concatenate([nrm, L, grad], axis=1), 2 * nv * dim + nv)
                cases.append(dict(name=name, dim=dim, cfg=[dim, model, eos, transport, conv, 0], what=2, bc=0, input=inp.tolist(), output=out.tolist()))
            # what 3: conversions, raw flux, Boussinesq source (2-D / 3-D)
            source = 1 if (dim >= 2 and model in (1, 3)) else 0
            inp, out = call(dim, model, eos, transport, conv, source, 3, 0, L, nc + nv + nv * dim + nv)
            cases.append(dict(name=name, dim=dim, cfg=[dim, model, eos, transport, conv, source], what=3, bc=0, input=inp.tolist(), output=out.tolist()))
            # what 4: ViewVariable::get, all 22 ViewVariableEnum values (the fall-through chain of its switch included)
            if conv in (2, 4):
                grad = 0.5 * rng4.normal(size=(n, nv * dim))
                eps = np.abs(rng4.normal(size=(n, 1))) * 0.01
                inp, out = call(dim, model, eos, transport, conv, 0, 4, 0, np.concatenate([L, grad, eps], axis=1), 22)
                cases.append(dict(name=name, dim=dim, cfg=[dim, model, eos, transport, conv, 0], what=4, bc=0, input=inp.tolist(), output=out.tolist()))
    doc = dict(source="reference sources compiled by oracle/Makefile target `ref` (oracle/ref_physics.cpp, oracle/ref_shim/, oracle/ref_patch.py); "
                      "g++ -std=c++23 -O2 -ffp-contract=off", params=PARAMS,
               layout={"0": "in: normal[D], consL[NV], consR[NV]; out: Riemann flux[NV]",
                       "1": "in: normal[D], consL[NV], user primitive[NV], conserved gradient[NV*D]; out: boundary comp[D+3], volCons[NV], intCons[NV], "
                            "convective boundary flux[NV], NS: interior comp after modifyBoundaryVariable[D+3], averaged viscous flux[NV]",
                       "2": "in: normal[D], cons[NV], conserved gradient[NV*D]; out: primitive gradient[NV*D], raw viscous flux[NV*D], normal viscous flux[NV]",
                       "3": "in: cons[NV]; out: comp[D+3], primitive[NV], raw convective flux[NV*D], source[NV]",
                       "4": "in: cons[NV], conserved gradient[NV*D], artificial viscosity; out: ViewVariable::get for the 22 ViewVariableEnum values "
                            "(0 where the variable names a direction the dimension lacks)"},
               cases=cases)
    path = os.path.join(HERE, "reference_physics.json")
    with open(path, "w") as f:
        json.dump(doc, f)
    print(f"{len(cases)} cases -> {path} ({os.path.getsize(path) / 1024:.0f} KB)")


if __name__ == "__main__":
    main()
