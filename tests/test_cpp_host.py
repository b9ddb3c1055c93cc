"""The header-only C++ host side (include/SubrosaDG_b200/SubrosaDG.hpp) and its example drivers: they must build with a
plain g++ (CPU check) and, on the GPU box, reproduce the ctypes path bit for bit and the exact travelling-wave solution."""
import os
import subprocess

import numpy as np
import pytest

import cases
from subrosadg_b200 import mesh as M

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EX = os.path.join(ROOT, "examples")


@pytest.fixture(scope="module")
def examples_built(built):
    r = subprocess.run(["make", "-C", EX], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return True


def test_examples_build_and_fail_loudly_without_gpu(examples_built, tmp_path):
    exe = os.path.join(EX, "_build", "periodic_2d_ceuler")
    assert os.path.exists(exe)
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    r = subprocess.run([exe, "1"], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode != 0 and "no usable CUDA device" in (r.stdout + r.stderr)


def test_cpp_periodic_box_equals_python_producer(examples_built, tmp_path):
    """makePeriodicBox (C++) and mesh.periodic_box_fast (Python) are the same integer maps / coordinates: checked through the
    flat mesh file format both sides share."""
    src = tmp_path / "dump.cpp"
    src.write_text('#include "SubrosaDG_b200/SubrosaDG.hpp"\n#include <iostream>\nint main(int c, char** v) { auto m = SubrosaDG::makePeriodicBox(std::atoi(v[1]), std::atoi(v[2]));'
                   ' auto r = SubrosaDG::MeshData::readFlat(v[3]); bool ok = m.dim == r.dim && m.n_int == r.n_int && m.n_bnd == r.n_bnd && m.le == r.le && m.lf == r.lf && m.re == r.re'
                   ' && m.rf == r.rf && m.rot == r.rot && m.lt == r.lt && m.rt == r.rt && m.blocks[0].coords == r.blocks[0].coords && m.blocks[0].type == r.blocks[0].type;'
                   ' std::cout << (ok ? "SAME" : "DIFFERENT") << std::endl; return ok ? 0 : 1; }\n')
    exe = tmp_path / "dump"
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    r = subprocess.run([cxx, "-std=c++20", "-O1", f"-I{ROOT}/include", str(src), "-o", str(exe), f"-L{ROOT}/subrosadg_b200", "-lsubrosadg_b200",
                        f"-Wl,-rpath,{ROOT}/subrosadg_b200"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    for dim, n in [(2, 10), (3, 5)]:
        path = tmp_path / f"m{dim}.sdgm"
        M.write_flat(M.periodic_box_fast(dim, n), path)
        r = subprocess.run([str(exe), str(dim), str(n), str(path)], capture_output=True, text=True)
        assert r.returncode == 0 and "SAME" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("name,dim,vel,dt", [("periodic_2d_ceuler", 2, [0.7, 0.3], 1e-3), ("periodic_3d_ceuler", 3, [0.5, 0.3, 0.2], 5e-4)])
def test_cpp_driver_matches_ctypes_path(examples_built, tmp_path, name, dim, vel, dt):
    from subrosadg_b200.solver import Solver
    out = tmp_path / "state.bin"
    n = 10 if dim == 2 else 6
    args = [os.path.join(EX, "_build", name), "20", str(out)] + ([str(n)] if dim == 3 else [])
    r = subprocess.run(args, capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 0, r.stdout + r.stderr
    err = float(r.stdout.strip().splitlines()[-1].split(":")[-1])
    assert err < 1e-4, r.stdout           # P3 on this grid: discretisation error of the travelling wave
    mesh = M.periodic_box_fast(dim, n)
    S = Solver(dict(p=3, conv_flux=2, rk=2), mesh, device=0)
    S.initializeSolver(cases.ic_density_wave(vel))
    S.stepSolver(dt, 20)
    ref = S.state_at_quadrature(S.types[0])
    got = np.fromfile(out, dtype=np.float64).reshape(ref.shape)
    assert np.array_equal(got, ref), f"C++ driver vs ctypes path: rel-L2 {cases.rel_l2(got, ref):.3e}"


@pytest.mark.gpu
@pytest.mark.parametrize("name,producer,scale,cfg,vel", [
    ("naca0012_2d_ceuler", "naca0012", 0.3, dict(p=3, conv_flux=2, rk=2), [0.63 * np.cos(np.deg2rad(2.0)), 0.63 * np.sin(np.deg2rad(2.0))]),
    ("karmanvortex_2d_cns", "karmanvortex", 0.3, dict(p=3, model=1, transport=2, mu=1.4 * 0.2 / 200.0, conv_flux=2, visc_flux=2, rk=2), [0.2, 0.0]),
    ("sphere_3d_cns", "sphere", 0.3, dict(p=3, model=1, transport=1, mu=1.4 * 0.2 / 200.0, conv_flux=2, visc_flux=2, rk=2), [0.0, 0.2, 0.0]),
])
def test_cpp_config_drivers_match_ctypes_path(examples_built, tmp_path, name, producer, scale, cfg, vel):
    """examples/{naca0012_2d_ceuler,karmanvortex_2d_cns,sphere_3d_cns}.cpp (configs 2, 3, 5 with the reference's IC / BC values)
    over a flat mesh file: same result as the Python host mirror, bit for bit, on every element type."""
    from subrosadg_b200.solver import Solver
    mesh = M.EXAMPLE_MESHES[producer](scale)
    path = tmp_path / "mesh.sdgm"
    M.write_flat(mesh, path)
    out = tmp_path / "state"
    r = subprocess.run([os.path.join(EX, "_build", name), str(path), "3", str(out)], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 0, r.stdout + r.stderr
    dim = mesh.dim

    def ic(x):
        one = np.ones(x.shape[:-1])
        return np.stack([1.4 * one] + [v * one for v in vel] + [one], axis=-1)

    def bc(x, phys, time=None):
        one = np.ones(x.shape[:-1])
        return np.stack([1.4 * one] + [np.where(phys == 2, 0.0, v) * one for v in vel] + [one], axis=-1)

    S = Solver(cfg, mesh, device=0)
    S.initializeSolver(ic, bc)
    dt = S.calculateDeltaTime(1.0)
    assert abs(float(r.stdout.strip().splitlines()[-1].split()[-1]) - dt) <= 1e-5 * dt   # printed with 6 significant digits
    S.stepSolver(dt, 3)
    assert len(S.types) == (2 if producer == "karmanvortex" else 1)
    for t in S.types:
        ref = S.state_at_quadrature(t)
        got = np.fromfile(f"{out}.{t}.bin", dtype=np.float64).reshape(ref.shape)
        assert np.isfinite(ref).all()
        assert np.array_equal(got, ref), f"{name} type {t}: C++ driver vs ctypes path rel-L2 {cases.rel_l2(got, ref):.3e}"
    assert dim == len(vel)


def _sod_ic(x):
    left = x[..., 0] <= 0.5
    return np.stack([np.where(left, 1.0, 0.125), np.where(left, 0.75, 0.0), np.where(left, 1.4, 0.8 * 1.4)], axis=-1)


def _sod_bc(x, phys, time=None):
    return np.stack([np.where(phys == 1, 1.0, 0.125), np.where(phys == 1, 0.75, 0.0), np.where(phys == 1, 1.4, 0.8 * 1.4)], axis=-1)


def _cavity(temperature, lid):
    """(ic, bc) of the two cavity examples: fluid at rest, density 1; `temperature` = {physical index: wall value, None: interior}, lid = physical index
    of the moving wall (u = 1) or None"""
    def ic(x):
        one = np.ones(x.shape[:-1])
        return np.stack([one, 0 * one, 0 * one, temperature[None] * one], axis=-1)

    def bc(x, phys, time=None):
        one = np.ones(x.shape[:-1])
        T = sum(np.where(phys == k, v, 0.0) for k, v in temperature.items() if k is not None)
        return np.stack([one, np.where(phys == lid, 1.0, 0.0) * one if lid else 0 * one, 0 * one, T * one], axis=-1)
    return ic, bc


WC_CAVITY = dict(model=3, eos=1, rho0=1.0, transport=1, visc_flux=2, rk=2, cp=1.0, cv=1.0)
MORE_DRIVERS = [
    # examples/sod_1d_ceuler.cpp: ShockCapturingEnum::ArtificialViscosity on lines, setArtificialViscosity(0.5), CFL 0.001
    ("sod_1d_ceuler", "sod_1d", 0.3, dict(p=3, conv_flux=2, rk=2, av_tolerance=0.5, av_factor=1.0), 0.001, (_sod_ic, _sod_bc)),
    # examples/lidcavity_2d_incns.cpp: IncompresibleNS, WeakCompressibleFluid(10, 1), mu = 1 / 5000, Lax-Friedrichs + BR2, moving lid
    ("lidcavity_2d_incns", "lidcavity_2d", 0.3, dict(WC_CAVITY, p=3, c0=10.0, mu=1.0 / 5000.0, conv_flux=1), 1.0, _cavity({None: 1.0, 1: 1.0, 2: 1.0}, 2)),
    # examples/thermalcavity_2d_incns.cpp: P1 quadrangles (seven-point rule, dense-operator path), Exact flux, Boussinesq(1, 0.5), hot / cold walls
    ("thermalcavity_2d_incns", "thermalcavity_2d", 0.1, dict(WC_CAVITY, p=1, c0=3.0, mu=float(np.sqrt(0.71 / 1e6)), conv_flux=4, source=1, beta=1.0, t_ref=0.5), 0.5,
     _cavity({None: 0.5, 1: 0.5, 2: 0.0, 3: 1.0}, None)),
]


@pytest.mark.gpu
@pytest.mark.parametrize("name,producer,scale,cfg,cfl,fields", MORE_DRIVERS, ids=[d[0] for d in MORE_DRIVERS])
def test_cpp_drivers_of_more_reference_examples_match_ctypes_path(examples_built, tmp_path, name, producer, scale, cfg, cfl, fields):
    """three more drivers that mirror examples of the reference (shock capturing in 1-D, weakly compressible Navier-Stokes with a moving wall,
    Boussinesq convection on P1 quadrangles) over flat mesh files: the same numbers as the Python host mirror, bit for bit"""
    from subrosadg_b200.solver import Solver
    mesh = M.EXAMPLE_MESHES[producer](scale)
    path = tmp_path / "mesh.sdgm"
    M.write_flat(mesh, path)
    out = tmp_path / "state"
    r = subprocess.run([os.path.join(EX, "_build", name), str(path), "4", str(out)], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 0, r.stdout + r.stderr
    ic, bc = fields
    S = Solver(cfg, mesh, device=0)
    S.initializeSolver(ic, bc)
    dt = S.calculateDeltaTime(cfl)
    assert abs(float(r.stdout.strip().splitlines()[-1].split()[-1]) - dt) <= 1e-5 * dt   # printed with 6 significant digits
    S.stepSolver(dt, 4)
    for t in S.types:
        ref = S.state_at_quadrature(t)
        got = np.fromfile(f"{out}.{t}.bin", dtype=np.float64).reshape(ref.shape)
        assert np.isfinite(ref).all()
        assert np.array_equal(got, ref), f"{name} type {t}: C++ driver vs ctypes path rel-L2 {cases.rel_l2(got, ref):.3e}"
    # something happened in four steps: the Sod jump moved, the lid / the hot wall set the fluid in motion
    U0 = Solver(cfg, mesh, device=0)
    U0.initializeSolver(ic, bc)
    assert cases.rel_l2(S.state_at_quadrature(S.types[0]), U0.state_at_quadrature(S.types[0])) > 1e-7


def _compile(src, exe):
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    r = subprocess.run([cxx, "-std=c++20", "-O1", "-Wall", "-Wextra", f"-I{ROOT}/include", str(src), "-o", str(exe), f"-L{ROOT}/subrosadg_b200",
                        "-lsubrosadg_b200", f"-Wl,-rpath,{ROOT}/subrosadg_b200"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return str(exe)


def test_boundary_time_driver_compiles(built, tmp_path):
    _compile(os.path.join(ROOT, "tests", "cpp", "boundary_time_driver.cpp"), tmp_path / "bt")


@pytest.mark.gpu
def test_time_varying_boundary_sees_the_reference_times(built, tmp_path):
    """System::solve assigns iteration_ AFTER stepSolver (SystemControl.cpp:175-177): the TimeVarying callback of step i is evaluated at
    t = (i - 1) dt, the first step at t = 0."""
    from subrosadg_b200.solver import Solver
    exe = _compile(os.path.join(ROOT, "tests", "cpp", "boundary_time_driver.cpp"), tmp_path / "bt")
    mesh = M.box(2, (5, 4), 0.0, 1.0, phys_bc={k: M.RIEMANN_FARFIELD for k in (1, 2, 3, 4)})
    M.write_flat(mesh, tmp_path / "mesh.sdgm")
    r = subprocess.run([exe, str(tmp_path / "mesh.sdgm"), str(tmp_path / "out"), "3"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    dt = 2.0e-3
    times = [float(x) for x in r.stdout.strip().splitlines()[-1].split()[1:]]
    assert np.allclose(times, [0.0, dt, 2 * dt], rtol=1e-5, atol=0), times

    def ic(x):
        one = np.ones(x.shape[:-1])
        return np.stack([1.4 * one, 0.3 * one, 0.1 * one, one], axis=-1)

    def bc(x, phys, time=0.0):
        one = np.ones(x.shape[:-1])
        return np.stack([1.4 * one, 0.3 * (1.0 + 5.0 * time) * one, 0.1 * one, one], axis=-1)

    S = Solver(dict(p=2, conv_flux=2, rk=2), mesh, device=0)
    S.initializeSolver(ic, bc)
    for i in range(1, 4):
        S.updateBoundaryVariable(bc, (i - 1) * dt)
        S.stepSolver(dt, 1)
    ref = S.state_at_quadrature(S.types[0])
    got = np.fromfile(tmp_path / "out" / "state.bin", dtype=np.float64).reshape(ref.shape)
    assert np.array_equal(got, ref), f"rel-L2 {cases.rel_l2(got, ref):.3e}"
    S2 = Solver(dict(p=2, conv_flux=2, rk=2), mesh, device=0)      # one step late (t = i dt) is a different solution
    S2.initializeSolver(ic, bc)
    for i in range(1, 4):
        S2.updateBoundaryVariable(bc, i * dt)
        S2.stepSolver(dt, 1)
    assert cases.rel_l2(S2.state_at_quadrature(S.types[0]), ref) > 1e-9


@pytest.mark.gpu
def test_step_solver_host_of_the_cpp_shim(built, tmp_path):
    """Solver<SC>::stepSolverHost (sdg_step_host behind the C++ mirror): one more step on coefficients held in host memory, in place,
    after the three steps of the driver; the TimeVarying callback sees t = iteration_ * delta_time_ = 3 dt."""
    from subrosadg_b200.solver import Solver
    exe = _compile(os.path.join(ROOT, "tests", "cpp", "boundary_time_driver.cpp"), tmp_path / "bt")
    mesh = M.box(2, (5, 4), 0.0, 1.0, phys_bc={k: M.RIEMANN_FARFIELD for k in (1, 2, 3, 4)})
    M.write_flat(mesh, tmp_path / "mesh.sdgm")
    r = subprocess.run([exe, str(tmp_path / "mesh.sdgm"), str(tmp_path / "out"), "3", "host"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    dt = 2.0e-3

    def ic(x):
        one = np.ones(x.shape[:-1])
        return np.stack([1.4 * one, 0.3 * one, 0.1 * one, one], axis=-1)

    def bc(x, phys, time=0.0):
        one = np.ones(x.shape[:-1])
        return np.stack([1.4 * one, 0.3 * (1.0 + 5.0 * time) * one, 0.1 * one, one], axis=-1)

    S = Solver(dict(p=2, conv_flux=2, rk=2), mesh, device=0)
    S.initializeSolver(ic, bc)
    for i in range(1, 4):
        S.updateBoundaryVariable(bc, (i - 1) * dt)
        S.stepSolver(dt, 1)
    t = S.types[0]
    U = S.get_state(t)
    S.updateBoundaryVariable(bc, 3 * dt)
    ref, _ = S.step_host(t, U, dt)
    got = np.fromfile(tmp_path / "out" / "coefficient_host.bin", dtype=np.float64).reshape(ref.shape)
    assert np.isfinite(ref).all() and cases.rel_l2(ref, U) > 1e-9
    assert np.array_equal(got, ref), f"rel-L2 {cases.rel_l2(got, ref):.3e}"
