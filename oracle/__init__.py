"""ctypes front end of the CPU oracle (oracle/oracle.cpp).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package;
the product (subrosadg_b200) never does.  The oracle restates SubrosaDG's stepSolver path
(src/Solver/TimeIntegration.cpp:326-350 and everything below it); parity with the real reference binary is
UNPINNED (the reference cannot be built here and ships no golden vectors) — see oracle/oracle.cpp header.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")

POINT, LINE, TRIANGLE, QUADRANGLE, TETRAHEDRON, PYRAMID, HEXAHEDRON = range(7)


def build(force: bool = False) -> str:
    """Compile oracle/oracle.cpp with the committed Makefile (g++ -O3 -fopenmp)."""
    if force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB_PATH)
        for f in ("oracle.cpp", "tables.hpp", "physics.hpp", "Makefile")
    ):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _LIB_PATH


class _Config(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in
                ("dim", "p", "model", "eos", "transport", "conv_flux", "visc_flux", "source", "rk", "dead_gradient", "accurate")] + \
               [(n, ctypes.c_double) for n in ("cp", "cv", "mu", "c0", "rho0", "beta", "t_ref")]


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = _LIB_PATH if os.path.exists(_LIB_PATH) else build()
        _lib = ctypes.CDLL(path)
        _lib.orc_last_error.restype = ctypes.c_char_p
    return _lib


def _chk(rc):
    if rc != 0:
        raise RuntimeError(lib().orc_last_error().decode())


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _ip(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))


def reference_nodes(etype: int, order: int) -> np.ndarray:
    cnt = ctypes.c_int32(0)
    _chk(lib().orc_reference_nodes(etype, order, None, ctypes.byref(cnt)))
    out = np.zeros((cnt.value, 3))
    _chk(lib().orc_reference_nodes(etype, order, _dp(out), ctypes.byref(cnt)))
    return out


def face_sequence(ftype: int, p: int, rotation: int) -> np.ndarray:
    n = {POINT: 1, LINE: p + 1, QUADRANGLE: (p + 1) ** 2}[ftype]
    out = np.zeros(n, dtype=np.int32)
    _chk(lib().orc_face_sequence(ftype, p, rotation, _ip(out)))
    return out


@dataclass
class Sizes:
    n: int
    Nb: int
    Nq: int
    Nf: int
    Naq: int
    nn: int
    Nqf: int
    Nv: int


class Oracle:
    """Mirror of the subset of SubrosaDG::Solver<SC> exercised by System::solve (src/Utils/SystemControl.cpp:159-195)."""

    def __init__(self, cfg: dict, mesh, threads: int | None = None):
        c = _Config()
        defaults = dict(dim=mesh.dim, p=cfg["p"], model=0, eos=0, transport=0, conv_flux=2, visc_flux=0, source=0, rk=2,
                        dead_gradient=1, accurate=1, cp=2.5, cv=25.0 / 14.0, mu=0.0, c0=1.0, rho0=1.0, beta=0.0, t_ref=0.0)
        defaults.update(cfg)
        for k, v in defaults.items():
            setattr(c, k, v)
        self.cfg = defaults
        self.h = ctypes.c_void_p()
        if threads:
            lib().orc_set_threads(int(threads))
        _chk(lib().orc_create(ctypes.byref(c), ctypes.byref(self.h)))
        self.mesh = mesh
        self.types = sorted(mesh.blocks)
        for t in self.types:
            b = mesh.blocks[t]
            coords = np.ascontiguousarray(b["coords"], dtype=np.float64)
            _chk(lib().orc_add_elements(self.h, t, coords.shape[0], b["geom_order"], _dp(coords)))
        f = mesh.faces
        arrs = [np.ascontiguousarray(f[k], dtype=np.int32) for k in ("le", "lt", "lf", "re", "rt", "rf", "rot", "bc", "phys")]
        _chk(lib().orc_set_faces(self.h, int(f["n_int"]), int(f["n_bnd"]), *[_ip(a) for a in arrs]))
        _chk(lib().orc_finalize(self.h))
        self.dim = mesh.dim
        self.Nv = mesh.dim + 2

    def __del__(self):
        try:
            if self.h:
                lib().orc_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def sizes(self, t) -> Sizes:
        out = np.zeros(8, dtype=np.int32)
        _chk(lib().orc_sizes(self.h, t, _ip(out)))
        return Sizes(*[int(x) for x in out])

    def table(self, t, which):
        s = self.sizes(t)
        D = self.dim
        shape = {0: (s.Nb, s.Nq), 1: (s.Nb, s.Nq * D), 2: (s.Nb, s.Naq), 3: (s.Nb, s.Nb), 4: (s.Nq, 3), 5: (s.Nq,),
                 6: (s.Nqf, 3), 7: (s.Nqf,), 8: (s.Naq, 3)}[which]
        out = np.zeros(shape)
        _chk(lib().orc_get_table(self.h, t, which, _dp(out)))
        return out.T if which in (0, 1, 2, 3) else out  # column-major -> [row, col]

    def element_geometry(self, t, which):
        s = self.sizes(t)
        D = self.dim
        shape = {0: (s.n, s.Nq, D), 1: (s.n, s.Nq), 2: (s.n, s.Nq, D * D), 3: (s.n, s.Nb, s.Nb), 4: (s.n,)}[which]
        out = np.zeros(shape)
        _chk(lib().orc_get_element_geometry(self.h, t, which, _dp(out)))
        return out

    def face_geometry(self, which):
        f = self.mesh.faces
        nf = int(f["n_int"]) + int(f["n_bnd"])
        s = self.sizes(self.types[0])
        shape = (nf, s.Nqf, self.dim) if which in (0, 1) else (nf, s.Nqf)
        out = np.zeros(shape)
        _chk(lib().orc_get_face_geometry(self.h, which, _dp(out)))
        return out

    def check_face_match(self) -> float:
        v = ctypes.c_double(0)
        _chk(lib().orc_check_face_match(self.h, ctypes.byref(v)))
        return v.value

    def quadrature_coordinates(self, t):
        return self.element_geometry(t, 0)

    def boundary_quadrature_coordinates(self):
        return self.face_geometry(0)[int(self.mesh.faces["n_int"]):]

    # -- Solver<SC>::initializeSolver (InitialCondition.cpp:151-186) ----------------------------------------------------
    def initialize(self, ic, bc=None):
        """ic(x[..., D]) -> primitive[..., Nv]; bc(x, phys[...]) -> primitive[..., Nv] (user callbacks on the host)."""
        for t in self.types:
            x = self.quadrature_coordinates(t)
            prim = np.ascontiguousarray(ic(x), dtype=np.float64)
            _chk(lib().orc_set_state_from_primitive(self.h, t, _dp(prim)))
        if int(self.mesh.faces["n_bnd"]) > 0:
            self.update_boundary(bc, None)

    def update_boundary(self, bc, time):
        xb = self.boundary_quadrature_coordinates()
        phys = np.asarray(self.mesh.faces["phys"])[int(self.mesh.faces["n_int"]):]
        physb = np.broadcast_to(phys[:, None], xb.shape[:2])
        prim = bc(xb, physb) if time is None else bc(xb, physb, time)
        prim = np.ascontiguousarray(prim, dtype=np.float64)
        _chk(lib().orc_set_boundary_primitive(self.h, _dp(prim)))

    def get_state(self, t):
        s = self.sizes(t)
        out = np.zeros((s.n, s.Nb, s.Nv))
        _chk(lib().orc_get_state(self.h, t, _dp(out)))
        return out

    def set_state(self, t, U):
        U = np.ascontiguousarray(U, dtype=np.float64)
        _chk(lib().orc_set_state(self.h, t, _dp(U)))

    def state_at_quadrature(self, t):
        s = self.sizes(t)
        out = np.zeros((s.n, s.Nq, s.Nv))
        _chk(lib().orc_get_state_at_quadrature(self.h, t, _dp(out)))
        return out

    def gradient_at_quadrature(self, t):
        s = self.sizes(t)
        out = np.zeros((s.n, s.Nq, s.Nv * self.dim))
        _chk(lib().orc_get_gradient_at_quadrature(self.h, t, _dp(out)))
        return out

    def compute_dt(self, cfl) -> float:
        v = ctypes.c_double(0)
        _chk(lib().orc_compute_dt(self.h, ctypes.c_double(cfl), ctypes.byref(v)))
        return v.value

    def step(self, dt, nsteps=1):
        err = np.zeros(self.Nv)
        _chk(lib().orc_step(self.h, ctypes.c_double(dt), int(nsteps), _dp(err)))
        return err

    def residual(self):
        """One residual evaluation: {type: (R_modal[n,Nb,Nv], rhs_at_quadrature[n,Nq,Nv])}."""
        _chk(lib().orc_eval_residual(self.h))
        out = {}
        for t in self.types:
            s = self.sizes(t)
            R = np.zeros((s.n, s.Nb, s.Nv))
            q = np.zeros((s.n, s.Nq, s.Nv))
            _chk(lib().orc_fetch_residual(self.h, t, _dp(R), _dp(q)))
            out[t] = (R, q)
        return out


def max_threads() -> int:
    return int(lib().orc_max_threads())
