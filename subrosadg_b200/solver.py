"""Host-side mirror of SubrosaDG::Solver<SimulationControl> over the C ABI (include/subrosadg_b200.h).

The reference drives its solver from System<SC>::solve() (src/Utils/SystemControl.cpp:159-195):
initializeSolver -> calculateDeltaTime -> stepSolver per time step -> relative_error_.  `Solver` below exposes the same
calls with the same meaning; user IC/BC callbacks stay on the host exactly like the reference's template
specialisations `InitialCondition<SC>::calculatePrimitiveFromCoordinate` (InitialCondition.cpp:38-39) and
`BoundaryCondition<SC>::calculatePrimitiveFromCoordinate` (BoundaryCondition.cpp:575-579).

Everything numerical happens inside libsubrosadg_b200.so (hand-written sm_100a kernels).  There is no fallback: if the
library is missing or no CUDA device is usable, construction raises.
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SDG_LIB", os.path.join(_HERE, "libsubrosadg_b200.so"))  # SDG_LIB: A/B builds of the same ABI

POINT, LINE, TRIANGLE, QUADRANGLE, TETRAHEDRON, PYRAMID, HEXAHEDRON = range(7)

# enum values of src/Utils/Enum.cpp
EQUATION_MODEL = dict(CompresibleEuler=0, CompresibleNS=1, IncompresibleEuler=2, IncompresibleNS=3)
EQUATION_OF_STATE = dict(IdealGas=0, WeakCompressibleFluid=1)
TRANSPORT_MODEL = dict(**{"None": 0}, Constant=1, Sutherland=2)
CONVECTIVE_FLUX = dict(Central=0, LaxFriedrichs=1, HLLC=2, Roe=3, Exact=4)
VISCOUS_FLUX = dict(**{"None": 0}, BR1=1, BR2=2)
SOURCE_TERM = dict(**{"None": 0}, Boussinesq=1)
TIME_INTEGRATION = dict(ForwardEuler=0, HeunRK2=1, SSPRK3=2)


class SdgConfig(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in
                ("dim", "p", "model", "eos", "transport", "conv_flux", "visc_flux", "source", "rk", "device", "chunk", "reorder")] + \
               [(n, ctypes.c_double) for n in ("cp", "cv", "mu", "c0", "rho0", "beta", "t_ref")]


EXPORTS = [
    "sdg_last_error", "sdg_version", "sdg_create", "sdg_destroy", "sdg_add_elements", "sdg_set_faces", "sdg_finalize", "sdg_sizes",
    "sdg_get_quadrature_coordinates", "sdg_get_boundary_quadrature_coordinates", "sdg_set_state_from_primitive",
    "sdg_set_boundary_primitive", "sdg_set_state", "sdg_get_state", "sdg_get_state_at_quadrature", "sdg_get_gradient_at_quadrature",
    "sdg_compute_dt", "sdg_step", "sdg_step_timed", "sdg_step_host", "sdg_step_host_info", "sdg_residual", "sdg_set_halo_send", "sdg_halo_pack", "sdg_halo_buffers_device", "sdg_step_begin",
    "sdg_stage_pass", "sdg_step_end", "sdg_num_passes", "sdg_num_stages", "sdg_stream", "sdg_synchronize", "sdg_set_state_device",
    "sdg_get_state_device", "sdg_launch_count", "sdg_debug_plan", "sdg_ipc_export", "sdg_ipc_connect", "sdg_halo_push", "sdg_halo_wait",
    "sdg_halo_doubles_per_element", "sdg_debug_physics", "sdg_uses_trace_rows", "sdg_set_halo_rows", "sdg_halo_unpack",
    "sdg_ipc_set_destination_units", "sdg_get_gradient_state", "sdg_get_boundary_gradient_state", "sdg_set_artificial_viscosity",
    "sdg_set_element_nodes", "sdg_get_node_artificial_viscosity", "sdg_get_element_artificial_viscosity", "sdg_update_artificial_viscosity",
    "sdg_get_view_variable", "sdg_av_node_buffer", "sdg_av_store",
]

_lib = None


def load_library():
    """dlopen the in-tree CUDA library; raises if it has not been built (no CPU path exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(make -C subrosadg_b200/csrc). The product has no CPU fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        lib.sdg_last_error.restype = ctypes.c_char_p
        lib.sdg_stream.restype = ctypes.c_void_p
        lib.sdg_launch_count.restype = ctypes.c_int64
        lib.sdg_destroy.restype = None
        _lib = lib
    return _lib


def _chk(rc):
    if rc != 0:
        raise RuntimeError(load_library().sdg_last_error().decode())


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _ip(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))


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


class Solver:
    """SubrosaDG::Solver<SC> (src/Solver/SolveControl.cpp:327-436) on one B200.

    cfg keys mirror the SimulationControl template parameters as Enum.cpp integers: p, model, eos, transport, conv_flux,
    visc_flux, source, rk, plus the physical-model parameters cp, cv, mu, c0, rho0, beta, t_ref.
    `n_ghost`: {type: count} — trailing elements of each block that are halo copies (multi-GPU partitions).
    `node_data`: ({type: global node tags of the block's elements}, global node number, {type: inner radii}) for shock-capturing runs of a partition.
    `device=-1` builds a plan-only context (host flattening, no compute) used by CPU tests of the host logic.
    """

    def __init__(self, cfg: dict, mesh, device: int = 0, n_ghost: dict | None = None, reorder: int = 1, node_data=None):
        lib = load_library()
        c = SdgConfig()
        vals = dict(dim=mesh.dim, p=cfg["p"], model=0, eos=0, transport=0, conv_flux=2, visc_flux=0, source=0, rk=2, device=device,
                    chunk=0, reorder=reorder, cp=2.5, cv=25.0 / 14.0, mu=0.0, c0=1.0, rho0=1.0, beta=0.0, t_ref=0.0)
        vals.update({k: v for k, v in cfg.items() if k in vals})
        for k, v in vals.items():
            setattr(c, k, v)
        self.cfg = vals
        self.h = ctypes.c_void_p()
        _chk(lib.sdg_create(ctypes.byref(c), ctypes.byref(self.h)))
        self.mesh = mesh
        self.dim = mesh.dim
        self.Nv = mesh.dim + 2
        self.types = sorted(mesh.blocks)
        n_ghost = n_ghost or {}
        for t in self.types:
            b = mesh.blocks[t]
            coords = np.ascontiguousarray(b["coords"], dtype=np.float64)
            _chk(lib.sdg_add_elements(self.h, t, coords.shape[0], int(n_ghost.get(t, 0)), int(b["geom_order"]), _dp(coords)))
        f = mesh.faces
        arrs = [np.ascontiguousarray(f[k], dtype=np.int32) for k in ("le", "lt", "lf", "re", "rt", "rf", "rot", "bc", "phys")]
        _chk(lib.sdg_set_faces(self.h, int(f["n_int"]), int(f["n_bnd"]), *[_ip(a) for a in arrs]))
        if cfg.get("av_tolerance") is not None:   # System::setArtificialViscosity with ShockCapturingEnum::ArtificialViscosity
            from . import mesh as M
            # node_data = (tags, node_number, inner radii): a partition passes the GLOBAL node tags of its owned + ghost elements
            tags, self.n_nodes, radii = node_data if node_data is not None else (*M.node_tags(mesh), {t: M.inner_radius(mesh, t) for t in self.types})
            _chk(lib.sdg_set_artificial_viscosity(self.h, ctypes.c_double(cfg["av_tolerance"]), ctypes.c_double(cfg.get("av_factor", 1.0)), int(self.n_nodes)))
            for t in self.types:
                r = np.ascontiguousarray(radii[t], dtype=np.float64)
                _chk(lib.sdg_set_element_nodes(self.h, t, _ip(np.ascontiguousarray(tags[t], dtype=np.int32)), _dp(r)))
        _chk(lib.sdg_finalize(self.h))
        self.relative_error_ = np.zeros(self.Nv)  # Solver::relative_error_, SolveControl.cpp:300
        self.delta_time_ = 0.0

    def close(self):
        if getattr(self, "h", None):
            load_library().sdg_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- sizes / coordinates -------------------------------------------------------------------------------------------------
    def sizes(self, t) -> Sizes:
        out = np.zeros(8, dtype=np.int32)
        _chk(load_library().sdg_sizes(self.h, t, _ip(out)))
        return Sizes(*[int(x) for x in out])

    def quadrature_coordinates(self, t):
        s = self.sizes(t)
        out = np.zeros((s.n, s.Nq, self.dim))
        _chk(load_library().sdg_get_quadrature_coordinates(self.h, t, _dp(out)))
        return out

    def boundary_quadrature_coordinates(self):
        s = self.sizes(self.types[0])
        out = np.zeros((int(self.mesh.faces["n_bnd"]), s.Nqf, self.dim))
        _chk(load_library().sdg_get_boundary_quadrature_coordinates(self.h, _dp(out)))
        return out

    # -- Solver::initializeSolver (InitialCondition.cpp:151-186) ------------------------------------------------------------
    def initializeSolver(self, ic, bc=None):
        """ic(x[..., D]) -> primitive[..., Nv]; bc(x, phys) -> primitive[..., Nv]: host callbacks, evaluated at the
        quadrature points like the reference does."""
        for t in self.types:
            prim = np.ascontiguousarray(ic(self.quadrature_coordinates(t)), dtype=np.float64)
            _chk(load_library().sdg_set_state_from_primitive(self.h, t, _dp(prim)))
        if int(self.mesh.faces["n_bnd"]) > 0:
            self.updateBoundaryVariable(bc, None)

    initialize = initializeSolver

    # -- Solver::updateBoundaryVariable (BoundaryCondition.cpp:29-74) ---------------------------------------------------------
    def updateBoundaryVariable(self, bc, time=None):
        xb = self.boundary_quadrature_coordinates()
        phys = np.asarray(self.mesh.faces["phys"])[int(self.mesh.faces["n_int"]):]
        physb = np.broadcast_to(phys[:, None], xb.shape[:2])
        prim = bc(xb, physb) if time is None else bc(xb, physb, time)
        prim = np.ascontiguousarray(prim, dtype=np.float64)
        _chk(load_library().sdg_set_boundary_primitive(self.h, _dp(prim)))

    # -- state access --------------------------------------------------------------------------------------------------------
    def get_state(self, t, out=None):
        """Modal coefficients [n][Nb][Nv]; `out` may be a caller-owned (e.g. pinned) float64 array of that shape."""
        s = self.sizes(t)
        if out is None:
            out = np.zeros((s.n, s.Nb, s.Nv))
        elif out.dtype != np.float64 or not out.flags.c_contiguous or out.size != s.n * s.Nb * s.Nv:
            raise ValueError("out must be a C-contiguous float64 array with n*Nb*Nv entries")
        _chk(load_library().sdg_get_state(self.h, t, _dp(out)))
        return out

    def set_state(self, t, U):
        U = np.ascontiguousarray(U, dtype=np.float64)
        _chk(load_library().sdg_set_state(self.h, t, _dp(U)))

    def state_at_quadrature(self, t):
        s = self.sizes(t)
        out = np.zeros((s.n, s.Nq, s.Nv))
        _chk(load_library().sdg_get_state_at_quadrature(self.h, t, _dp(out)))
        return out

    def gradient_at_quadrature(self, t):
        s = self.sizes(t)
        out = np.zeros((s.n, s.Nq, s.Nv * self.dim))
        _chk(load_library().sdg_get_gradient_at_quadrature(self.h, t, _dp(out)))
        return out

    def gradient_state(self, t):
        """variable_gradient_basis_function_coefficient_ [n][Nb][Nv*D] of the current state (RawBinary.cpp:75-88)."""
        s = self.sizes(t)
        out = np.zeros((s.n, s.Nb, s.Nv * self.dim))
        _chk(load_library().sdg_get_gradient_state(self.h, t, _dp(out)))
        return out

    def boundary_gradient_state(self):
        """Per boundary face (face order) the parent's gradient block as RawBinary.cpp:89-154 writes it (BR1 total, BR2 volume + that
        face's lift): a flat array of Nb(parent type) x Nv*D rows."""
        f = self.mesh.faces
        lt = np.asarray(f["lt"])[int(f["n_int"]):int(f["n_int"]) + int(f["n_bnd"])]
        n = sum(self.sizes(int(t)).Nb for t in lt) * self.Nv * self.dim
        out = np.zeros(int(n))
        if n:
            _chk(load_library().sdg_get_boundary_gradient_state(self.h, _dp(out)))
        return out

    def view_variable(self, t, variable: int):
        """ViewVariable::get (VariableConvertor.cpp:754-872) at the volume quadrature points, [n][Nq]; variable = ViewVariableEnum value."""
        s = self.sizes(t)
        out = np.zeros((s.n, s.Nq))
        _chk(load_library().sdg_get_view_variable(self.h, t, int(variable), _dp(out)))
        return out

    def update_artificial_viscosity(self):
        _chk(load_library().sdg_update_artificial_viscosity(self.h))

    def node_artificial_viscosity(self):
        """Solver::node_artificial_viscosity_ [node_number] as of the last step (SpatialDiscrete.cpp:137-192)."""
        out = np.zeros(self.n_nodes)
        _chk(load_library().sdg_get_node_artificial_viscosity(self.h, _dp(out)))
        return out

    def element_artificial_viscosity(self, t):
        from . import mesh as M
        out = np.zeros((self.sizes(t).n, M.N_BASIC[t]))
        _chk(load_library().sdg_get_element_artificial_viscosity(self.h, t, _dp(out)))
        return out

    # -- Solver::calculateDeltaTime (TimeIntegration.cpp:133-179) -------------------------------------------------------------
    def calculateDeltaTime(self, cfl) -> float:
        v = ctypes.c_double(0)
        _chk(load_library().sdg_compute_dt(self.h, ctypes.c_double(cfl), ctypes.byref(v)))
        self.delta_time_ = v.value
        return v.value

    compute_dt = calculateDeltaTime

    # -- Solver::stepSolver (TimeIntegration.cpp:326-350) ---------------------------------------------------------------------
    def stepSolver(self, dt=None, nsteps=1, want_error=True):
        dt = self.delta_time_ if dt is None else dt
        err = np.zeros(self.Nv)
        _chk(load_library().sdg_step(self.h, ctypes.c_double(dt), int(nsteps), _dp(err) if want_error else None))
        if want_error:
            self.relative_error_ = err
        return err

    def step_timed(self, dt, nsteps=1):
        """stepSolver x nsteps; returns (relative_error_, device milliseconds measured with CUDA events on the stream)."""
        err = np.zeros(self.Nv)
        ms = ctypes.c_float(0)
        _chk(load_library().sdg_step_timed(self.h, ctypes.c_double(dt), int(nsteps), _dp(err), ctypes.byref(ms)))
        self.relative_error_ = err
        return err, float(ms.value)

    def step_host(self, t, U_in, dt=None, out=None):
        """One stepSolver on a state held in host memory (sdg_step_host): modal coefficients in, modal coefficients out (into `out`,
        which may be `U_in` itself), upload / stages / download streamed where the kernels allow it.  Returns (out, relative_error_)."""
        dt = self.delta_time_ if dt is None else dt
        s = self.sizes(t)
        if U_in.dtype != np.float64 or not U_in.flags.c_contiguous or U_in.size != s.n * s.Nb * s.Nv:
            raise ValueError("U_in must be a C-contiguous float64 array with n*Nb*Nv entries")
        if out is None:
            out = np.zeros((s.n, s.Nb, s.Nv))
        elif out.dtype != np.float64 or not out.flags.c_contiguous or out.size != s.n * s.Nb * s.Nv:
            raise ValueError("out must be a C-contiguous float64 array with n*Nb*Nv entries")
        err = np.zeros(self.Nv)
        _chk(load_library().sdg_step_host(self.h, t, ctypes.c_double(dt), _dp(U_in), _dp(out), _dp(err)))
        self.relative_error_ = err
        return out, err

    def step_host_info(self):
        g = ctypes.c_int32(0); f = ctypes.c_double(0)
        _chk(load_library().sdg_step_host_info(self.h, ctypes.byref(g), ctypes.byref(f)))
        return int(g.value), float(f.value)

    def step(self, dt, nsteps=1):
        return self.stepSolver(dt, nsteps)

    def residual(self):
        """{type: (R_modal[n,Nb,Nv], dU/dt at the quadrature points [n,Nq,Nv])} for the current state."""
        out = {}
        for t in self.types:
            s = self.sizes(t)
            R = np.zeros((s.n, s.Nb, s.Nv))
            q = np.zeros((s.n, s.Nq, s.Nv))
            _chk(load_library().sdg_residual(self.h, t, _dp(R), _dp(q)))
            out[t] = (R, q)
        return out

    def debug_plan(self, what):
        """Host-plan diagnostics (see sdg_debug_plan)."""
        cnt = ctypes.c_int64(0)
        lib = load_library()
        _chk(lib.sdg_debug_plan(self.h, what, None, None, ctypes.byref(cnt)))
        if what < 10 or what >= 90:   # doubles (>= 90: dense-operator path tables, see sdg_debug_plan)
            out = np.zeros(cnt.value)
            _chk(lib.sdg_debug_plan(self.h, what, _dp(out), None, ctypes.byref(cnt)))
        else:
            out = np.zeros(cnt.value, dtype=np.int32)
            _chk(lib.sdg_debug_plan(self.h, what, None, _ip(out), ctypes.byref(cnt)))
        return out

    def synchronize(self):
        _chk(load_library().sdg_synchronize(self.h))

    @property
    def launch_count(self) -> int:
        return int(load_library().sdg_launch_count(self.h))
