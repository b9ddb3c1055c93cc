"""Solver::writeRawBinary of the C++ host side (include/SubrosaDG_b200/SubrosaDG.hpp): the reference's .zst container
(src/View/RawBinary.cpp:42-74) and payload order (:75-191), read back by a restatement of the reference's readers
(RawBinaryCompress::read, :59-73; InitialCondition.cpp:41-80) written here in numpy over libzstd / pyarrow's zstd codec."""
import ctypes
import os
import struct
import subprocess

import numpy as np
import pytest

import cases
from subrosadg_b200 import mesh as M

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def compile_cpp(src, exe):
    r = subprocess.run([CXX, "-std=c++20", "-O1", "-Wall", "-Wextra", f"-I{ROOT}/include", str(src), "-o", str(exe), f"-L{ROOT}/subrosadg_b200",
                        "-lsubrosadg_b200", f"-Wl,-rpath,{ROOT}/subrosadg_b200"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def _libzstd():
    try:
        z = ctypes.CDLL("libzstd.so.1")
    except OSError:
        return None
    z.ZSTD_compressBound.restype = ctypes.c_size_t
    z.ZSTD_compressBound.argtypes = [ctypes.c_size_t]
    z.ZSTD_decompress.restype = ctypes.c_size_t
    z.ZSTD_decompress.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t]
    z.ZSTD_isError.restype = ctypes.c_uint
    z.ZSTD_isError.argtypes = [ctypes.c_size_t]
    return z


def compress_bound(n):
    """ZSTD_COMPRESSBOUND (zstd.h)"""
    return n + (n >> 8) + (((128 << 10) - n) >> 11 if n < (128 << 10) else 0)


def read_raw_binary(path):
    """RawBinaryCompress::read (RawBinary.cpp:59-73): 8-byte destination capacity, then one zstd frame.  Decoded with every decoder
    available (libzstd through ctypes, pyarrow's bundled zstd); they must agree."""
    blob = open(path, "rb").read()
    (capacity,) = struct.unpack("<Q", blob[:8])
    frame = blob[8:]
    outs = []
    z = _libzstd()
    if z is not None:
        buf = ctypes.create_string_buffer(max(capacity, 1))
        n = z.ZSTD_decompress(buf, capacity, frame, len(frame))
        assert not z.ZSTD_isError(n), f"{path}: libzstd cannot decode the frame"
        outs.append(buf.raw[:n])
    try:
        import pyarrow as pa
        if pa.Codec.is_available("zstd"):
            # the frame header carries the content size (both writers set it): bytes 6..13 of the raw writer's header; ask zstd itself
            size = len(outs[0]) if outs else None
            if size is None:
                size = struct.unpack("<Q", frame[6:14])[0]
            outs.append(pa.Codec("zstd").decompress(frame, decompressed_size=size).to_pybytes() if size else b"")
    except ImportError:
        pass
    assert outs, "no zstd decoder available"
    assert all(o == outs[0] for o in outs)
    assert capacity >= len(outs[0])
    return capacity, outs[0]


def pattern(n):
    """the payload of tests/cpp/raw_binary_container.cpp"""
    out = bytearray(n)
    x = 0x9E3779B97F4A7C15
    mask = (1 << 64) - 1
    for i in range(n):
        if i % 4096 < 1024:
            out[i] = (i // 4096) & 0xFF
            continue
        x ^= (x << 13) & mask
        x ^= x >> 7
        x ^= (x << 17) & mask
        out[i] = x & 0xFF
    return bytes(out)


def test_container_both_writers_both_readers(built, tmp_path):
    exe = compile_cpp(os.path.join(ROOT, "tests", "cpp", "raw_binary_container.cpp"), tmp_path / "container")
    sizes = [0, 1, 1000, 131072, 131073, 300001]
    r = subprocess.run([str(exe), str(tmp_path)] + [str(s) for s in sizes], capture_output=True, text=True)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout + r.stderr
    have_lib = "libzstd found" in r.stdout
    for n in sizes:
        want = pattern(n)
        for w in (["lib"] if have_lib else []) + ["raw"]:
            cap, got = read_raw_binary(tmp_path / f"{w}_{n}.zst")
            assert got == want, f"{w} writer, {n} bytes"
            assert cap == compress_bound(n)          # the header the reference writes: ZSTD_compressBound(payload size)
            z = _libzstd()
            if z is not None:
                assert cap == z.ZSTD_compressBound(n)
        if have_lib and n >= 1000:   # level 1 of the real compressor squeezes the constant stretches of the pattern; the raw writer cannot
            assert os.path.getsize(tmp_path / f"lib_{n}.zst") < os.path.getsize(tmp_path / f"raw_{n}.zst")


def test_host_reader_decodes_files_written_by_the_reference(built, tmp_path):
    """RawBinaryCompress::read of the C++ host side (what InitialConditionEnum::LastStep / SpecificFile call) on the .zst files the REFERENCE'S
    own Solver::writeRawBinary wrote with libzstd (tests/golden/reference_raw_*.zst, tests/golden/make_reference_sweeps.py)"""
    if _libzstd() is None:
        pytest.skip("compressed blocks need libzstd.so.1, as in the reference")
    exe = compile_cpp(os.path.join(ROOT, "tests", "cpp", "raw_binary_container.cpp"), tmp_path / "container")
    files = sorted(f for f in os.listdir(os.path.join(ROOT, "tests", "golden")) if f.startswith("reference_raw_") and f.endswith(".zst"))
    assert len(files) >= 4
    for f in files:
        src = os.path.join(ROOT, "tests", "golden", f)
        r = subprocess.run([str(exe), "decode", src, str(tmp_path / "payload.bin")], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        cap, want = read_raw_binary(src)
        assert open(tmp_path / "payload.bin", "rb").read() == want and len(want) % 8 == 0 and cap == compress_bound(len(want))
        assert os.path.getsize(src) - 8 < len(want)        # really compressed: the reference's ZSTD_compress at level 1


# ---- payload (RawBinary.cpp:75-191) -------------------------------------------------------------------------------------------------------
def node_number(mesh):
    """Mesh::node_number_ as the C++ host side counts it: distinct coordinate tuples over all blocks"""
    pts = np.concatenate([np.asarray(b["coords"], dtype=np.float64).reshape(-1, mesh.dim) for b in mesh.blocks.values()])
    return np.unique(pts, axis=0).shape[0]


def parse_payload(buf, mesh, nb, ns):
    """Restated reader of Solver::writeRawBinary's stream: per element type (ascending ElementEnum) and element [Nb][Nv] (+ [Nb][Nv*D] for
    Navier-Stokes, as InitialCondition.cpp:47-56 skips it and ElementViewSolver::calcluateElementViewVariable reads it,
    RawBinary.cpp:193-240); per boundary face the same blocks of its parent (RawBinary.cpp:89-154, read at :242-296); node_number_ reals
    at the tail (ViewSolver::calcluateViewVariable seeks them first, :352-356).  nb = {type: Nb}.  Returns (U, G, [(type, parent, local face, U row, G row)], node artificial viscosity)."""
    a = np.frombuffer(buf, dtype=np.float64)
    D, Nv = mesh.dim, mesh.dim + 2
    pos = 0
    U, G = {}, {}
    for t in sorted(mesh.blocks):
        n = np.asarray(mesh.blocks[t]["coords"]).shape[0]
        row, grow = nb[t] * Nv, (nb[t] * Nv * D if ns else 0)
        blk = a[pos:pos + n * (row + grow)].reshape(n, row + grow)
        pos += n * (row + grow)
        U[t] = blk[:, :row].reshape(n, nb[t], Nv)
        if ns:
            G[t] = blk[:, row:].reshape(n, nb[t], Nv * D)
    f = mesh.faces
    bnd = []
    for i in range(int(f["n_int"]), int(f["n_int"]) + int(f["n_bnd"])):
        t = int(f["lt"][i])
        row, grow = nb[t] * Nv, (nb[t] * Nv * D if ns else 0)
        bnd.append((t, int(f["le"][i]), int(f["lf"][i]), a[pos:pos + row].reshape(nb[t], Nv), a[pos + row:pos + row + grow]))
        pos += row + grow
    nn = node_number(mesh)
    av = a[pos:pos + nn]
    pos += nn
    assert pos == a.size, f"payload has {a.size} reals, the mesh accounts for {pos}"
    return U, G, bnd, av


@pytest.mark.gpu
@pytest.mark.parametrize("name,producer,cfg,vel", [
    ("naca0012_2d_ceuler", "naca0012", dict(p=3, conv_flux=2, rk=2), [0.63 * np.cos(np.deg2rad(2.0)), 0.63 * np.sin(np.deg2rad(2.0))]),
    ("karmanvortex_2d_cns", "karmanvortex", dict(p=3, model=1, transport=2, mu=1.4 * 0.2 / 200.0, conv_flux=2, visc_flux=2, rk=2), [0.2, 0.0]),
    ("sphere_3d_cns", "sphere", dict(p=3, model=1, transport=1, mu=1.4 * 0.2 / 200.0, conv_flux=2, visc_flux=2, rk=2), [0.0, 0.2, 0.0]),
])
def test_raw_binary_of_the_config_drivers(built, tmp_path, name, producer, cfg, vel):
    """raw/<name>_0.zst and raw/<name>_3.zst written by examples/<name>.cpp (System::solve, SystemControl.cpp:166-183) hold what the
    ctypes path reports for the same run: modal state, gradient coefficients, the boundary parents' blocks (BR2: volume part + the
    lift of that face), zeros for the node artificial viscosity."""
    from subrosadg_b200.solver import Solver
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "examples")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    mesh = M.EXAMPLE_MESHES[producer](0.3)
    path = tmp_path / "mesh.sdgm"
    M.write_flat(mesh, path)
    r = subprocess.run([os.path.join(ROOT, "examples", "_build", name), str(path), "3"], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 0, r.stdout + r.stderr
    ns = cfg.get("model", 0) == 1

    def ic(x):
        one = np.ones(x.shape[:-1])
        return np.stack([1.4 * one] + [v * one for v in vel] + [one], axis=-1)

    def bc(x, phys, time=None):
        one = np.ones(x.shape[:-1])
        return np.stack([1.4 * one] + [np.where(phys == 2, 0.0, v) * one for v in vel] + [one], axis=-1)

    S = Solver(cfg, mesh, device=0)
    S.initializeSolver(ic, bc)
    dt = S.calculateDeltaTime(1.0)
    nb = {t: S.sizes(t).Nb for t in S.types}
    raw = tmp_path / "build" / "out" / name / "raw"
    import oracle
    O = oracle.Oracle(dict(cfg), mesh)      # the modal basis at the quadrature points, to evaluate the file like the View reader does
    phi = {t: O.table(t, 0) for t in S.types}                # [Nq][Nb]: the reader's `coefficient * modal_value_` (RawBinary.cpp:204-205)
    for step in (0, 3):
        if step:
            S.stepSolver(dt, step)
        cap, buf = read_raw_binary(raw / f"{name}_{step}.zst")
        assert cap == compress_bound(len(buf))
        U, G, bnd, av = parse_payload(buf, mesh, nb, ns)
        assert np.all(av == 0.0) and av.size == node_number(mesh)
        for t in S.types:
            assert np.array_equal(U[t], S.get_state(t)), f"{name} step {step} type {t}: modal state"
            uq = np.einsum("ebv,qb->eqv", U[t], phi[t])       # what the reference's view reader reconstructs from the file
            assert cases.rel_l2(uq, S.state_at_quadrature(t)) < 1e-13
            if ns:
                assert np.array_equal(G[t], S.gradient_state(t)), f"{name} step {step} type {t}: gradient coefficients"
        state = {t: S.get_state(t) for t in S.types}
        gb = S.boundary_gradient_state() if ns else None
        at = 0
        for (t, e, lf, u_row, g_row) in bnd:
            assert np.array_equal(u_row, state[t][e])
            if ns:
                assert np.array_equal(g_row, gb[at:at + g_row.size])
                at += g_row.size
        assert len(bnd) == int(mesh.faces["n_bnd"]) > 0
    lines = open(tmp_path / "build" / "out" / name / "error.txt").read().splitlines()
    assert len(lines) == 2 + 3 and lines[0].startswith("|    Time     |     rho     |") and all(len(l) == len(lines[0]) for l in lines)


def _build_restart(tmp_path, kind, poly):
    exe = tmp_path / f"restart_{kind}_{poly}"
    r = subprocess.run([CXX, "-std=c++20", "-O1", "-Wall", "-Wextra", f"-DIC_KIND={kind}", f"-DPOLY={poly}", f"-I{ROOT}/include",
                        os.path.join(ROOT, "tests", "cpp", "restart_driver.cpp"), "-o", str(exe), f"-L{ROOT}/subrosadg_b200", "-lsubrosadg_b200",
                        f"-Wl,-rpath,{ROOT}/subrosadg_b200"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return str(exe)


def test_restart_drivers_compile(built, tmp_path):
    for kind, poly in (("Function", "P3"), ("LastStep", "P3"), ("SpecificFile", "P3")):
        _build_restart(tmp_path, kind, poly)


@pytest.mark.gpu
def test_last_step_and_specific_file_restarts(built, tmp_path):
    """LastStep continues a run from raw/run_<start>.zst; SpecificFile starts order P from the leading columns of an order P-1 file
    (InitialCondition.cpp:41-80).  Both read the files this repo's writeRawBinary wrote."""
    run = lambda exe, *a: subprocess.run([exe] + [str(x) for x in a], capture_output=True, text=True, stdin=subprocess.DEVNULL)
    d = tmp_path / "a"
    r = run(_build_restart(tmp_path, "Function", "P3"), d, 0, 10, 5, tmp_path / "a10.bin")
    assert r.returncode == 0 and "iteration 10" in r.stdout, r.stdout + r.stderr
    assert sorted(os.listdir(d / "raw")) == ["run_0.zst", "run_10.zst", "run_5.zst"]
    err_a = open(d / "error.txt").read().splitlines()
    r = run(_build_restart(tmp_path, "LastStep", "P3"), d, 5, 10, 5, tmp_path / "b10.bin")
    assert r.returncode == 0 and "iteration 10" in r.stdout, r.stdout + r.stderr
    a10, b10 = np.fromfile(tmp_path / "a10.bin"), np.fromfile(tmp_path / "b10.bin")
    assert a10.size == 36 * 16 * 4
    # the device keeps the state at the nodes: a restart goes modal -> nodal -> modal once more (one rounding of the transform)
    assert cases.rel_l2(b10, a10) < 1e-13
    err_b = open(d / "error.txt").read().splitlines()
    assert len(err_a) == len(err_b) == 12 and err_b[:7] == err_a[:7]
    # order P-1 run, then order P from its last file
    c = tmp_path / "c"
    r = run(_build_restart(tmp_path, "Function", "P2"), c, 0, 2, 2, tmp_path / "c2.bin")
    assert r.returncode == 0, r.stdout + r.stderr
    e = tmp_path / "e"
    r = run(_build_restart(tmp_path, "SpecificFile", "P3"), e, 0, 1, 1, tmp_path / "e1.bin", c / "raw" / "run_2.zst")
    assert r.returncode == 0, r.stdout + r.stderr
    mesh = M.periodic_box_fast(2, 6)
    t = 3
    _, buf = read_raw_binary(e / "raw" / "run_0.zst")
    U0, _, _, _ = parse_payload(buf, mesh, {t: 16}, True)
    c2 = np.fromfile(tmp_path / "c2.bin").reshape(36, 9, 4)
    scale = np.abs(c2).max()
    assert np.abs(U0[t][:, :9, :] - c2).max() < 1e-13 * scale and np.abs(U0[t][:, 9:, :]).max() < 1e-13 * scale


def test_shock_driver_compiles(built, tmp_path):
    compile_cpp(os.path.join(ROOT, "tests", "cpp", "shock_driver.cpp"), tmp_path / "shock")


@pytest.mark.gpu
def test_shock_capturing_run_and_its_raw_file(built, tmp_path):
    """ShockCapturingEnum::ArtificialViscosity through System<SC> (setArtificialViscosity, SystemControl.cpp:105-108): same states as the
    ctypes path, and the tail of the raw file is Solver::node_artificial_viscosity_ of the last step (RawBinary.cpp:187-188)."""
    from subrosadg_b200.solver import Solver
    exe = compile_cpp(os.path.join(ROOT, "tests", "cpp", "shock_driver.cpp"), tmp_path / "shock")
    mesh = M.box(1, (24,), 0.0, 1.0, phys_bc={1: M.RIEMANN_FARFIELD, 2: M.RIEMANN_FARFIELD})
    M.write_flat(mesh, tmp_path / "mesh.sdgm")
    steps = 8
    r = subprocess.run([str(exe), str(tmp_path / "mesh.sdgm"), str(tmp_path / "out"), str(steps)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr

    def state(x, *_):
        s = np.clip((x[..., 0] - 0.51) / 0.01, -1.0, 1.0)
        rho, p = 0.5625 - 0.4375 * s, 0.55 - 0.45 * s
        return np.stack([rho, np.zeros_like(rho), 1.4 * p / rho], axis=-1)

    S = Solver(dict(p=2, conv_flux=2, rk=2, av_tolerance=0.5, av_factor=1.0), mesh, device=0)
    S.initializeSolver(state, state)
    dt = S.calculateDeltaTime(0.1)
    assert abs(float(r.stdout.split()[-1]) - dt) <= 1e-5 * dt
    S.stepSolver(dt, steps)
    t = M.LINE
    ref = S.state_at_quadrature(t)
    got = np.fromfile(tmp_path / "out" / "state.bin").reshape(ref.shape)
    assert np.array_equal(got, ref), f"rel-L2 {cases.rel_l2(got, ref):.3e}"
    _, buf = read_raw_binary(tmp_path / "out" / "raw" / f"sod_{steps}.zst")
    U, _, bnd, av = parse_payload(buf, mesh, {t: 3}, False)
    assert np.array_equal(U[t], S.get_state(t)) and len(bnd) == 2
    assert av.max() > 0.0 and np.array_equal(av, S.node_artificial_viscosity())
    _, buf0 = read_raw_binary(tmp_path / "out" / "raw" / "sod_0.zst")
    assert np.all(parse_payload(buf0, mesh, {t: 3}, False)[3] == 0.0)     # written before the first step: InitialCondition.cpp:156-157
