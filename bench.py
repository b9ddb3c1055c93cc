#!/usr/bin/env python
"""bench.py — GDOF·stage/s of the DG residual + RK stage (fp64) on BASELINE.json's config 4
(periodic_3d_ceuler: synthetic structured hex mesh, 128^3 = 2,097,152 elements, p = 3, HLLC, SSPRK3).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--cells C] [--p P]

`value`  : scalar-DOF stage updates per second (Ne*Nb*Nv*stages*steps / device time) with the state resident in HBM.
`e2e`    : the same metric through the reference-facing calls with HOST buffers every step
           (sdg_set_state -> sdg_step -> sdg_get_state; the host<->device copies are inside the timed region).
`roofline`: algorithmic bytes (24 B per scalar DOF·stage = read U, read U_last, write U) over the stage
           kernel's mean duration (CUDA events on the library's stream) against the measured HBM copy bandwidth.
`cpu_baseline` / `--impl reference`: the CPU oracle (restatement of the reference's oneTBB path, OpenMP, all host cores)
           timed on a bounded sample (a smaller cube of the same family, same p / flux / RK).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "GDOF·stage/s per RK stage (fp64)"
UNIT = "GDOF·stage/s"
BYTES_PER_DOF_STAGE = 24.0  # 3*S per element per stage for the inviscid path


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic(model, elements):
    """DRAM bytes per launch of the dominant kernel(s) from the committed `ncu --set full` capture (profiles/traffic.json:
    dram__bytes_read.sum + dram__bytes_write.sum per element of the profiled launch), scaled to this run's element count."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None
    with open(p) as f:
        d = json.load(f)
    per = d.get(model, {}).get("dram_bytes_per_element")
    return None if per is None else float(per) * elements


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in self.rows if len(r) >= 7 for k in range(4) if r[3 + k].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def ic_config4(x):
    """examples/periodic_3d_ceuler.cpp:33-35"""
    rho = 1.0 + 0.2 * np.sin(np.pi * (x[..., 0] + x[..., 1] + x[..., 2]))
    out = np.empty(x.shape[:-1] + (5,))
    out[..., 0] = rho; out[..., 1] = 0.5; out[..., 2] = 0.3; out[..., 3] = 0.2; out[..., 4] = 1.4 / rho
    return out


CFG = dict(p=3, model=0, conv_flux=2, rk=2)
# north_star's Navier-Stokes throughput target: same cube, CompresibleNS, HLLC, BR2, constant mu = 1.4e-3
CFG_NS = dict(p=3, model=1, conv_flux=2, rk=2, visc_flux=2, transport=1, mu=1.4 * 0.2 / 200.0)
BYTES_PER_DOF_STAGE_NS = 144.0  # 3S + 2DS + 16 Nv D Naq per element = 46,080 B per P3 hexahedron


def run_oracle(cells, p, steps, warmup, threads=None, base=None):
    """CPU restatement of the reference path on a cells^3 cube; returns (GDOF·stage/s, seconds, cores)."""
    import oracle
    from subrosadg_b200 import mesh as M
    mesh = M.periodic_box_fast(3, cells)
    cfg = dict(base or CFG); cfg["p"] = p
    cfg["accurate"] = 0   # the timed baseline runs the reference's arithmetic (plain double M^-1 apply), not the checker's extended-precision mode
    if threads is None:   # all host cores this process may use, whatever OMP_NUM_THREADS says (torchrun sets it to 1)
        threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    O = oracle.Oracle(cfg, mesh, threads=threads)
    O.initialize(ic_config4)
    dt = 1e-4
    if warmup:
        O.step(dt, warmup)
    t0 = time.perf_counter()
    O.step(dt, steps)
    sec = time.perf_counter() - t0
    s = O.sizes(6)
    dof = s.n * s.Nb * s.Nv
    return dof * 3 * steps / sec / 1e9, sec, int(threads)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--cells", type=int, default=128)
    ap.add_argument("--p", type=int, default=3)
    ap.add_argument("--cpu-cells", type=int, default=32)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-ns-target", action="store_true", help="skip the Navier-Stokes 96^3 measurement that the default Euler run appends as `ns_target`")
    ap.add_argument("--model", default="euler", choices=["euler", "ns"], help="euler: BASELINE configs[3] (the metric's config); ns: north_star's NS-BR2 target cube")
    a = ap.parse_args()
    base_cfg = CFG if a.model == "euler" else CFG_NS
    bytes_per_dof = BYTES_PER_DOF_STAGE if a.model == "euler" else BYTES_PER_DOF_STAGE_NS
    if a.model == "ns" and a.cells == 128:
        a.cells = 96   # N = 96 so that the NS buffers fit comfortably
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    warmup = max(a.warmup, 3) if a.impl == "ours" else a.warmup
    if a.model == "euler":
        workload = f"periodic_3d_ceuler synthetic structured hex mesh {a.cells}^3, p={a.p}, CompresibleEuler, HLLC, SSPRK3 (BASELINE configs[3])"
    else:
        workload = f"periodic cube of configs[3] with CompresibleNS, HLLC, BR2, constant mu=1.4e-3, {a.cells}^3 hexes, p={a.p}, SSPRK3 (north_star NS target)"

    if a.impl == "reference":
        if rank != 0:
            return
        import oracle   # the reference arm loads the CPU oracle only: the product library is neither built nor dlopened here
        oracle.build()
        # bounded sample of the same workload: a cpu_cells^3 cube of the same family; per-DOF rate is size independent
        val, sec, cores = run_oracle(a.cpu_cells, a.p, max(1, a.steps), a.warmup, base=base_cfg)
        sample = f"{a.cpu_cells}^3 hexes p={a.p}, {max(1, a.steps)} steps x 3 stages, {sec:.1f} s (CPU restatement of the reference algorithm incl. its dense M^-1 and gradient sweeps; the reference itself cannot be built here)"
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
                          "ms_per_step": 1e3 * sec / max(1, a.steps), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                          "data": "synthetic", "config": {"workload": workload, "timed_sample": sample},
                          "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
                          "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import torch
    if world > 1:
        from subrosadg_b200 import parallel
        parallel.bench_main(a, workload, METRIC, UNIT, bytes_per_dof, peaks, ClockSampler, ic_config4, base_cfg, ns_cfg=None if (a.model == "ns" or a.no_ns_target) else CFG_NS,
                            ns_bytes_per_dof=BYTES_PER_DOF_STAGE_NS)
        return

    from subrosadg_b200 import mesh as M
    from subrosadg_b200.solver import Solver
    torch.cuda.init()
    hbm, how = peaks()

    def measure(model, cells, steps, warm):
        """device-timed stage throughput of one model on a cells^3 cube; returns (solver, dict)"""
        cfg = dict(CFG if model == "euler" else CFG_NS); cfg["p"] = a.p
        S = Solver(cfg, M.periodic_box_fast(3, cells), device=0)
        S.initializeSolver(ic_config4)
        t = S.types[0]
        sz = S.sizes(t)
        dof = sz.n * sz.Nb * sz.Nv
        dt = S.calculateDeltaTime(1.0)
        S.step_timed(dt, warm)
        l0 = S.launch_count
        torch.cuda.synchronize()
        err, ms = S.step_timed(dt, steps)
        torch.cuda.synchronize()
        per = BYTES_PER_DOF_STAGE if model == "euler" else BYTES_PER_DOF_STAGE_NS
        stage_ms = ms / (steps * 3)
        achieved = per * dof / (stage_ms * 1e-3) / 1e9
        return S, dict(sz=sz, dof=dof, dt=dt, ms=ms, err=err, launches=S.launch_count - l0, stage_ms=stage_ms, achieved=achieved,
                       value=dof * 3 * steps / (ms * 1e-3) / 1e9, traffic=measured_traffic(model, sz.n), kernel=kernel_names(model, S))

    def kernel_names(model, S):
        k = os.environ.get("SDG_EULER_KERNEL", "trace")
        if model == "euler":
            return {"trace": "nslStageKernel<affine,HLLC,inviscid> (thread per zeta-line, published face traces)", "link": "nslStageKernel<affine,HLLC,inviscid,gather>",
                    "line": "eulerLineKernel<4,8,affine,HLLC>"}.get(k, k)
        return "nslGradKernel<affine> + nslStageKernel<affine,HLLC,viscous> (one stage = both launches)" if not os.environ.get("SDG_NS_NODE_KERNEL") \
            else "nsGradKernel<3,4,4,affine> + nsStageKernel<3,4,4,affine,HLLC>"

    clocks = ClockSampler(); clocks.start()
    S, m = measure(a.model, a.cells, a.steps, warmup)
    ck = clocks.stop()
    sz, dof, dt, ms, err, nst = m["sz"], m["dof"], m["dt"], m["ms"], m["err"], 3
    t = S.types[0]
    out = {"metric": METRIC, "value": m["value"], "unit": UNIT, "n_gpus": 1, "steps": a.steps, "warmup": warmup, "ms_per_step": ms / a.steps,
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": workload, "elements": sz.n, "scalar_dof": dof, "dt": dt, "l2": f"state {dof * 8 / 1e9:.2f} GB per buffer >> 126 MB L2 (inputs larger than L2, no flush needed)",
                      "relative_error": [float(x) for x in err]},
           "gpu_launches": int(m["launches"]), "clocks": ck,
           "roofline": {"bound": "hbm", "achieved": m["achieved"], "peak": hbm, "unit": "GB/s", "frac": m["achieved"] / hbm, "traffic": m["traffic"],
                        "kernel": m["kernel"], "kernel_ms": m["stage_ms"], "algorithmic_bytes_per_launch": bytes_per_dof * dof,
                        "peak_source": how}}

    if not a.no_e2e:
        # end to end through the C ABI with host buffers: modal state in, step, modal state out (pinned host memory)
        U = torch.empty((sz.n, sz.Nb, sz.Nv), dtype=torch.float64).pin_memory()
        Un = U.numpy()
        Un[...] = S.get_state(t)
        n_e2e = max(1, min(a.steps, 3))
        # (a) phase after phase: sdg_set_state -> sdg_step -> sdg_get_state
        S.set_state(t, Un); S.stepSolver(dt, 1); S.get_state(t, out=Un)  # warm
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            S.set_state(t, Un)                 # host (pinned) modal coefficients -> device
            e = S.stepSolver(dt, 1)            # one time step = 3 stages; relative_error_ comes back to the host
            S_out = S.get_state(t, out=Un)     # device -> the caller's host buffer
        torch.cuda.synchronize()
        sec_phases = time.perf_counter() - t0
        # (b) the same step as ONE call on host buffers, streamed: upload groups -> stages of the chunks whose inputs have arrived ->
        #     downloads, PCIe busy in both directions (sdg_step_host; bit-identical to (a), tests/test_step_host.py)
        S.step_host(t, Un, dt, out=Un)         # warm: builds the dependency levels, allocates the staging arrays
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            S_out, e = S.step_host(t, Un, dt, out=Un)
        torch.cuda.synchronize()
        sec_e = time.perf_counter() - t0
        groups, early = S.step_host_info()
        if groups == 0:                        # a context that does not stream: (b) is the composition (a)
            sec_e = min(sec_e, sec_phases)
        out["e2e"] = {"value": dof * nst * n_e2e / sec_e / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(dof * 8), "d2h_bytes_per_step": int(dof * 8 + 8 * sz.Nv),
                      "steps": n_e2e, "ms_per_step": 1e3 * sec_e / n_e2e,
                      "note": "sdg_step_host: host modal coefficients in, one time step, host modal coefficients out (+ relative_error_) every step; "
                              "upload, stages and download streamed per dependency level",
                      "upload_groups": groups, "groups_downloaded_during_upload": early,
                      "phase_after_phase": {"value": dof * nst * n_e2e / sec_phases / 1e9, "ms_per_step": 1e3 * sec_phases / n_e2e,
                                            "note": "sdg_set_state -> sdg_step -> sdg_get_state"}}
        del S_out, e
    if not a.no_cpu:
        import __graft_entry__ as g
        g.build()
        cpu_steps = 6 if a.model == "euler" else 4   # 10-20 s of CPU work on the GPU box's host cores
        val, sec_c, cores = run_oracle(a.cpu_cells, a.p, cpu_steps, 0, base=base_cfg)
        extra = {}
        if a.model == "euler":   # the reference runs the gradient sweeps G1-G4 for Euler although nothing reads them
            cfg_lean = dict(base_cfg); cfg_lean["dead_gradient"] = 0
            val2, sec2, _ = run_oracle(a.cpu_cells, a.p, 3, 0, base=cfg_lean)
            extra = {"value_without_dead_gradient_sweeps": val2}
        out["cpu_baseline"] = {**extra, "value": val, "unit": UNIT, "cores": cores, "kind": "port",
                               "sample": f"{a.cpu_cells}^3 hexes p={a.p}, {cpu_steps} steps x 3 stages, {sec_c:.1f} s; CPU restatement of the reference algorithm (dense per-element M^-1, gradient sweeps included), OpenMP on all host cores"}
    if a.model == "euler" and not a.no_ns_target:
        # north_star's only numeric kernel target — 3-D hex Navier-Stokes p = 3 residual + RK stage >= 50 % of the HBM roofline — measured in
        # the same run (the Euler solver is released first: both fit one GPU, but not next to each other at the end-to-end buffers' size)
        S.close(); del S
        torch.cuda.empty_cache()
        S2, n = measure("ns", 96, max(3, min(a.steps, 5)), 3)
        out["ns_target"] = {"workload": f"periodic cube of configs[3] with CompresibleNS, HLLC, BR2, constant mu=1.4e-3, 96^3 hexes, p={a.p}, SSPRK3", "value": n["value"],
                            "unit": UNIT, "ms_per_stage": n["stage_ms"], "roofline": {"bound": "hbm", "achieved": n["achieved"], "peak": hbm, "unit": "GB/s",
                                                                                       "frac": n["achieved"] / hbm, "traffic": n["traffic"], "kernel": n["kernel"],
                                                                                       "algorithmic_bytes_per_launch": BYTES_PER_DOF_STAGE_NS * n["dof"]},
                            "gpu_launches": int(n["launches"]), "relative_error": [float(x) for x in n["err"]]}
        S2.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
