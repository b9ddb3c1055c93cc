"""Generates tests/golden/reference_sweeps.json from the REFERENCE'S OWN solver.

oracle/Makefile (target `ref`) compiles /root/reference/src/Solver/{InitialCondition,SpatialDiscrete,TimeIntegration,...}.cpp and
src/Mesh/{BasisFunction,Quadrature}.cpp — where they lie — behind oracle/ref_sweeps.cpp into oracle/_ref/libref_sweeps.so: the reference's
initializeSolver, calculateDeltaTime and stepSolver (all eight sweeps, RK update, relative error) run on Mesh<SC> objects filled from the
arrays below.  See the header of oracle/ref_sweeps.cpp for what is and what is not the reference's code.  Run in the development
container (needs /root/reference); the committed JSON travels:

    python tests/golden/make_reference_sweeps.py

Consumers: tests/test_reference_sweeps.py (CPU: the oracle against these numbers; GPU: the CUDA path)."""
import ctypes
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from subrosadg_b200 import mesh as M   # noqa: E402

FAR, SLIP, NOSLIP, ISO = M.RIEMANN_FARFIELD, M.ADIABATIC_SLIP_WALL, M.ADIABATIC_NONSLIP_WALL, M.ISOTHERMAL_NONSLIP_WALL
INFLOW, OUTFLOW = M.VELOCITY_INFLOW, M.PRESSURE_OUTFLOW
WC = dict(eos=1, c0=10.0, rho0=1.0)   # EquationOfState<WeakCompressibleFluid> of the incompressible examples
WARP2 = lambda x: x + 0.03 * np.sin(np.pi * x[:, ::-1])
WARP3 = lambda x: x + 0.02 * np.sin(np.pi * np.roll(x, 1, axis=1))


# cases whose final state the reference's Solver::writeRawBinary (RawBinary.cpp:42-57,75-191) also writes to a committed .zst file:
# BR2 (boundary parents carry volume + lift of that face), BR1, and a shock-capturing run on a hybrid mesh (two element types, viscosity tail)
RAW_FILES = ("quad_p3_ns_br2_sutherland", "quad_p2_ns_br1_roe_heun_periodic", "av_hybrid_p3_radial_jump", "hex_p2_ns_br2_constant", "hybrid_p3_ns_br2_sutherland",
             "hex_p3_ns_br2_constant_curved")


def cases():
    """(name, compiled control type in oracle/ref_sweeps.cpp, oracle / product configuration, mesh, free-stream velocity, amplitude, steps, cfl).
    Shock-capturing cases carry av_tolerance / av_factor in the configuration and jump = (width, radius) instead of a velocity."""
    mu = 0.01
    ann = dict(r0=0.5, r1=3.0)
    return [
        ("quad_p2_euler_hllc_ssprk3", 0, dict(p=2, conv_flux=2, rk=2),
         M.box(2, (4, 3), 0.0, 1.0, geom_order=2, warp=WARP2, phys_bc={1: FAR, 2: FAR, 3: SLIP, 4: FAR}), [0.5, 0.1, 0.0], 0.02, 3, 0.5),
        ("quad_p3_ns_br2_sutherland", 1, dict(p=3, model=1, transport=2, mu=mu, conv_flux=2, visc_flux=2, rk=2),
         M.box(2, (3, 3), 0.0, 1.0, geom_order=3, warp=WARP2, phys_bc={1: FAR, 2: FAR, 3: NOSLIP, 4: ISO}), [0.3, 0.05, 0.0], 0.02, 2, 0.3),
        ("quad_p2_ns_br1_roe_heun_periodic", 2, dict(p=2, model=1, transport=1, mu=mu, conv_flux=3, visc_flux=1, rk=1),
         M.periodic_box(2, 4), [0.4, 0.2, 0.0], 0.05, 3, 0.3),
        ("line_p3_euler_lf_forward_euler", 3, dict(p=3, conv_flux=1, rk=0),
         M.box(1, (8,), 0.0, 1.0, phys_bc={1: FAR, 2: FAR}), [0.4, 0.0, 0.0], 0.05, 4, 0.2),
        ("hex_p2_ns_br2_constant", 4, dict(p=2, model=1, transport=1, mu=mu, conv_flux=2, visc_flux=2, rk=2),
         M.box(3, (2, 2, 3), 0.0, 1.0, geom_order=2, warp=WARP3, periodic_axes=(2,), phys_bc={1: FAR, 2: FAR, 3: NOSLIP, 4: SLIP}), [0.3, 0.1, 0.05], 0.02, 2, 0.3),
        ("triangle_p2_euler_hllc", 5, dict(p=2, conv_flux=2, rk=2),
         M.annulus(2, 8, r0=0.5, r1=2.0, geom_order=1, tri_rings=2), [0.4, 0.05, 0.0], 0.02, 3, 0.5),
        ("hybrid_p3_ns_br2_sutherland", 6, dict(p=3, model=1, transport=2, mu=mu, conv_flux=2, visc_flux=2, rk=2),
         M.annulus(3, 8, r0=0.5, r1=2.5, geom_order=3, tri_rings=1, phys_bc={1: FAR, 2: NOSLIP}), [0.2, 0.0, 0.0], 0.01, 2, 0.3),
        ("hex_p3_euler_periodic", 7, dict(p=3, conv_flux=2, rk=2), M.periodic_box_fast(3, 3), [0.5, 0.3, 0.2], 0.05, 2, 0.5),
        # ShockCapturingEnum::ArtificialViscosity (sod_1d / sedovblast_2d / explosion_2d / cylinder_2d control types): jump = (width, radius)
        ("av_line_p2_sod", 8, dict(p=2, conv_flux=2, rk=2, av_tolerance=0.5, av_factor=1.0),
         M.box(1, (24,), 0.0, 1.0, phys_bc={1: FAR, 2: FAR}), (0.02, 0.0), 0.0, 4, 0.2),
        ("av_quad_p3_oblique_jump", 9, dict(p=3, conv_flux=2, rk=2, av_tolerance=1.0, av_factor=2.0),
         M.box(2, (8, 6), 0.0, 1.0, geom_order=2, warp=WARP2, phys_bc={1: FAR, 2: FAR, 3: SLIP, 4: SLIP}), (0.03, 0.0), 0.0, 3, 0.2),
        ("av_triangle_p2_radial_jump", 10, dict(p=2, conv_flux=2, rk=2, av_tolerance=1.0, av_factor=1.0),
         M.annulus(6, 16, geom_order=1, tri_rings=6, **ann), (0.08, 1.5), 0.0, 3, 0.05),
        ("av_hybrid_p3_radial_jump", 11, dict(p=3, conv_flux=2, rk=2, av_tolerance=1.0, av_factor=2.0),
         M.annulus(6, 16, geom_order=3, stretch=1.2, tri_rings=3, **ann), (0.05, 1.6), 0.0, 2, 0.2),
        ("av_hex_p2_oblique_jump_roe", 12, dict(p=2, conv_flux=3, rk=2, av_tolerance=1.0, av_factor=1.0),
         M.box(3, (4, 3, 3), 0.0, 1.0, periodic_axes=(2,), phys_bc={1: FAR, 2: FAR, 3: SLIP, 4: SLIP}), (0.05, 0.0), 0.0, 2, 0.2),
        # the incompressible (weakly compressible) control types of examples/*_incns.cpp / shearlayer_2d_inceuler.cpp
        ("inc_quad_p3_ns_lf_br2_inflow_outflow", 13, dict(WC, p=3, model=3, transport=1, mu=mu, conv_flux=1, visc_flux=2, rk=2),
         M.box(2, (5, 4), 0.0, 1.0, geom_order=2, warp=WARP2, phys_bc={1: INFLOW, 2: OUTFLOW, 3: ISO, 4: NOSLIP}), [0.3, 0.05, 0.0], 0.02, 3, 0.3),
        ("inc_quad_p1_ns_exact_boussinesq_cavity", 14, dict(WC, p=1, model=3, transport=1, mu=mu, conv_flux=4, visc_flux=2, source=1, beta=0.5, t_ref=1.0, rk=2),
         M.box(2, (6, 5), 0.0, 1.0, phys_bc={1: ISO, 2: ISO, 3: NOSLIP, 4: NOSLIP}), [0.1, 0.05, 0.0], 0.02, 4, 0.3),
        ("inc_quad_p1_euler_lf_periodic", 15, dict(WC, p=1, model=2, conv_flux=1, rk=2), M.periodic_box(2, 6), [0.3, -0.2, 0.0], 0.05, 4, 0.5),
        ("inc_hex_p1_ns_lf_br2", 16, dict(WC, p=1, model=3, transport=1, mu=mu, conv_flux=1, visc_flux=2, rk=2),
         M.box(3, (3, 3, 2), 0.0, 1.0, phys_bc={1: INFLOW, 2: OUTFLOW, 3: NOSLIP, 4: SLIP, 5: ISO, 6: NOSLIP}), [0.3, 0.05, 0.1], 0.02, 3, 0.3),
        ("inc_hex_p3_ns_exact_boussinesq", 17, dict(WC, p=3, model=3, transport=1, mu=mu, conv_flux=4, visc_flux=2, source=1, beta=0.5, t_ref=1.0, rk=2),
         M.box(3, (3, 3, 2), 0.0, 1.0, geom_order=2, warp=WARP3, periodic_axes=(1,), phys_bc={1: ISO, 2: ISO, 5: NOSLIP, 6: NOSLIP}), [0.1, 0.05, 0.1], 0.02, 2, 0.3),
        ("inc_quad_p4_ns_exact_periodic", 18, dict(WC, p=4, model=3, transport=1, mu=mu, conv_flux=4, visc_flux=2, rk=2), M.periodic_box(2, 4), [0.3, -0.2, 0.0], 0.05, 2, 0.3),
        # the remaining compressible control types of examples/ (sphere_3d_cns = the north-star kernel family; blasius_3d / delta_3d_cns; rae2822_2d_cns;
        # khinstability_2d_ceuler; sod_1d / shuosher_1d_ceuler)
        ("hex_p3_ns_br2_constant_curved", 19, dict(p=3, model=1, transport=1, mu=mu, conv_flux=2, visc_flux=2, rk=2),
         M.box(3, (3, 3, 2), 0.0, 1.0, geom_order=2, warp=WARP3, phys_bc={1: FAR, 2: FAR, 3: NOSLIP, 4: FAR, 5: ISO, 6: SLIP}), [0.3, 0.1, 0.05], 0.02, 2, 0.3),
        ("hex_p1_ns_br2_constant", 20, dict(p=1, model=1, transport=1, mu=mu, conv_flux=2, visc_flux=2, rk=2),
         M.box(3, (4, 3, 3), 0.0, 1.0, phys_bc={1: FAR, 2: FAR, 3: NOSLIP, 4: FAR, 5: FAR, 6: FAR}), [0.3, 0.1, 0.05], 0.02, 3, 0.3),
        ("quad_p5_ns_br2_sutherland", 21, dict(p=5, model=1, transport=2, mu=mu, conv_flux=2, visc_flux=2, rk=2),
         M.box(2, (3, 3), 0.0, 1.0, geom_order=2, warp=WARP2, phys_bc={1: FAR, 2: FAR, 3: NOSLIP, 4: FAR}), [0.5, 0.1, 0.0], 0.02, 2, 0.3),
        ("av_quad_p5_oblique_jump", 22, dict(p=5, conv_flux=2, rk=2, av_tolerance=1.0, av_factor=2.0),
         M.box(2, (6, 5), 0.0, 1.0, phys_bc={1: FAR, 2: FAR, 3: SLIP, 4: SLIP}), (0.04, 0.0), 0.0, 2, 0.2),
        # BoundaryTimeEnum::TimeVarying: vel[3] = growth rate of the boundary velocity, 1 + rate * t with t = iteration * dt (first step: t = 0)
        ("inc_quad_p2_ns_time_varying_inflow", 24, dict(WC, p=2, model=3, transport=1, mu=mu, conv_flux=1, visc_flux=2, rk=2),
         M.box(2, (5, 4), 0.0, 1.0, phys_bc={1: INFLOW, 2: OUTFLOW, 3: ISO, 4: NOSLIP}), [0.3, 0.05, 0.0, 400.0], 0.02, 4, 0.3),
        ("av_line_p3_sod", 23, dict(p=3, conv_flux=2, rk=2, av_tolerance=1.0, av_factor=1.0),
         M.box(1, (24,), 0.0, 1.0, phys_bc={1: FAR, 2: FAR}), (0.01, 0.0), 0.0, 4, 0.2),
    ]


def time_varying(vel):
    """True for the BoundaryTimeEnum::TimeVarying cases: the consumer re-evaluates the boundary callback at t = iteration * dt before every step"""
    return not isinstance(vel, tuple) and len(vel) > 3 and vel[3] != 0.0


def fields(dim, vel, amp, cfg=None):
    """the analytic fields compiled into oracle/ref_sweeps.cpp (fieldAt / jumpAt): initial condition and (amp = 0) boundary values"""
    weak = bool(cfg) and cfg.get("eos", 0) == 1
    if isinstance(vel, tuple):          # shock-capturing cases: (jump width, jump radius)
        width, radius = vel

        def jump(x, *_):
            if radius > 0.0:
                s = np.tanh((np.sqrt((x * x).sum(axis=-1)) - radius) / width)
                rho, p = 0.75 - 0.25 * s, 0.75 - 0.25 * s
            else:
                c = 1.0 / np.sqrt(float(dim))
                s = np.tanh(((x * c).sum(axis=-1) - 0.5 * c * dim) / width)
                rho, p = 0.5625 - 0.4375 * s, 0.55 - 0.45 * s
            return np.stack([rho] + [np.zeros_like(rho)] * dim + [1.4 * p / rho], axis=-1)
        return jump, jump

    rate = vel[3] if len(vel) > 3 else 0.0

    def make(a):
        def f(x, phys=None, time=None):
            grow = 1.0 + rate * (time or 0.0) if phys is not None else 1.0       # boundary callback only (BoundaryTimeEnum::TimeVarying)
            s = np.sin(np.pi * x[..., 0])
            if dim >= 2:
                s = s * np.cos(np.pi * x[..., 1])
            if dim >= 3:
                s = s * np.cos(np.pi * x[..., 2])
            g = 1.0 + a * s
            if weak:
                return np.stack([cfg["rho0"] * (1.0 + 0.01 * a * s)] + [vel[d] * g * grow for d in range(dim)] + [1.0 + 2.0 * a * s], axis=-1)
            return np.stack([1.4 * g] + [(vel[d] * g + 0.0 * s) * grow for d in range(dim)] + [1.0 * g], axis=-1)
        return f
    return make(amp), make(0.0)


def run_reference(lib, case_id, cfg, mesh, vel, amp, steps, cfl, raw_path=None):
    import oracle
    O = oracle.Oracle(dict(cfg), mesh)          # geometry factors only (never stepped here)
    types = O.types
    dp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    ip = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))
    geo = {t: [np.ascontiguousarray(O.element_geometry(t, w)) for w in range(5)] for t in types}
    sizes = {t: O.sizes(t) for t in types}
    coef = {t: np.zeros((sizes[t].n, sizes[t].Nb, mesh.dim + 2)) for t in types}
    PP = ctypes.POINTER(ctypes.c_double) * len(types)
    arr = lambda w: PP(*[dp(geo[t][w]) for t in types])
    f = mesh.faces
    fint = [np.ascontiguousarray(f[k], dtype=np.int32) for k in ("le", "lt", "lf", "re", "rt", "rf", "rot", "bc", "phys")]
    FI = (ctypes.POINTER(ctypes.c_int32) * 9)(*[ip(a) for a in fint])
    xf, nrm, fjw = (np.ascontiguousarray(O.face_geometry(w)) for w in range(3))
    shock = isinstance(vel, tuple)
    params = np.array([2.5, 25.0 / 14.0, cfg.get("mu", 0.0), amp] + ([0.0, 0.0, 0.0, vel[0], vel[1], cfg["av_tolerance"], cfg["av_factor"]] if shock
                      else [vel[0], vel[1], vel[2], 0.0, 0.0, 0.0, 1.0])
                      + [cfg.get("c0", 1.0), cfg.get("rho0", 1.0), cfg.get("beta", 0.0), cfg.get("t_ref", 0.0), 1.0 if cfg.get("eos", 0) == 1 else 0.0,
                         vel[3] if (not shock and len(vel) > 3) else 0.0])
    tags, n_nodes = M.node_tags(mesh) if (shock or raw_path) else ({}, 1)     # Mesh::node_number_ sizes the tail of the raw file
    radius = {t: np.ascontiguousarray(M.inner_radius(mesh, t)) for t in types} if shock else {}
    IP = ctypes.POINTER(ctypes.c_int32) * len(types)
    node_av = np.zeros(n_nodes)
    relerr = np.zeros(mesh.dim + 2)
    dt = ctypes.c_double(0)
    tarr = np.array(types, dtype=np.int32); narr = np.array([sizes[t].n for t in types], dtype=np.int32)
    lib.ref_sweeps.restype = ctypes.c_int
    lib.ref_sweeps_set_raw_path(raw_path.encode() if raw_path else None)   # the reference's Solver::writeRawBinary after the last step
    lib.ref_sweeps_error.restype = ctypes.c_char_p
    rc = lib.ref_sweeps(case_id, dp(params), len(types), ip(tarr), ip(narr), arr(0), arr(1), arr(2), arr(3), arr(4), int(f["n_int"]), int(f["n_bnd"]), FI,
                        dp(xf), dp(nrm), dp(fjw), int(steps), ctypes.c_double(cfl), ctypes.c_double(0.0), PP(*[dp(coef[t]) for t in types]), dp(relerr),
                        ctypes.byref(dt), int(n_nodes), IP(*[ip(tags[t]) for t in types]) if shock else None, PP(*[dp(radius[t]) for t in types]) if shock else None,
                        dp(node_av) if shock else None)
    assert rc == 0, lib.ref_sweeps_error().decode()
    return coef, relerr, dt.value, (node_av if shock else None)


def main():
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], check=True)
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_sweeps.so"))
    out = []
    for name, case_id, cfg, mesh, vel, amp, steps, cfl in cases():
        coef0, _, _, _ = run_reference(lib, case_id, cfg, mesh, vel, amp, 0, cfl)            # Solver::initializeSolver alone
        coef, relerr, dt, node_av = run_reference(lib, case_id, cfg, mesh, vel, amp, steps, cfl)
        assert all(np.isfinite(c).all() for c in coef.values()), name
        if node_av is not None:
            assert node_av.max() > 0.0 and (node_av == 0.0).any(), name      # the case switches the viscosity on somewhere, not everywhere
        if name in RAW_FILES:     # the reference's own writer and the system's libzstd: tests/golden/reference_raw_<name>.zst
            raw = os.path.join(HERE, f"reference_raw_{name}.zst")
            again, _, _, _ = run_reference(lib, case_id, cfg, mesh, vel, amp, steps, cfl, raw_path=raw)
            assert all(np.array_equal(again[t], coef[t]) for t in coef)
            print(f"  raw file {raw} ({os.path.getsize(raw)} bytes)")
        out.append(dict(name=name, steps=steps, cfl=cfl, dt=dt, relative_error=relerr.tolist(), node_artificial_viscosity=None if node_av is None else node_av.tolist(),
                        initial={str(t): c.ravel().tolist() for t, c in coef0.items()}, state={str(t): c.ravel().tolist() for t, c in coef.items()},
                        mesh_checksum={str(t): float(np.asarray(b["coords"]).sum()) for t, b in mesh.blocks.items()}))
        print(f"{name}: dt {dt:.6e} relative_error {relerr}")
    path = os.path.join(HERE, "reference_sweeps.json")
    json.dump(dict(source="the reference's Solver<SC>::initializeSolver / calculateDeltaTime / stepSolver compiled from /root/reference/src (oracle/ref_sweeps.cpp)",
                   cases=out), open(path, "w"))
    print(f"{len(out)} cases -> {path} ({os.path.getsize(path) // 1024} KB)")


if __name__ == "__main__":
    main()
