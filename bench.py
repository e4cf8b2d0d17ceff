"""Benchmark of the hot path on the BASELINE workload (contract: see the task statement).

Step = one application of the hot path to one state: a random orbital rotation
followed by a diagonal Coulomb evolution (BASELINE.json configs[1]: norb=16,
nelec=(5,5), 19.1 M amplitudes, 305 MB -- larger than the 126 MB L2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N > 1 (torchrun, one rank per GPU): every rank runs its own replica of the
workload on its own GPU ("weak" scaling; the row-sharded configuration is a
separate code path); the time is the max over ranks.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

NORB, NELEC = 16, (5, 5)
WORKLOAD = "random orbital rotation + diag-Coulomb evolution on a random state, norb=16 nelec=(5,5), 19.1M amplitudes (305 MB)"
METRIC = "orbital-rotation + diag-Coulomb applications/sec"
UNIT = "applications/s"


def make_inputs(seed: int = 1602):
    """SURVEY.md section 8d, config C2 (one generator threaded through in this order)."""
    from oracle import rand  # generators only; byte-identical to ffsim_b200.random

    rng = np.random.default_rng(seed)
    dim = int(np.prod([__import__("math").comb(NORB, k) for k in NELEC]))
    vec = rand.random_state_vector(dim, seed=rng)
    u = rand.random_unitary(NORB, seed=rng)
    mat = rand.random_real_symmetric_matrix(NORB, seed=rng)
    return vec, u, mat, 1.0


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.samples, self._stop, self._thread = index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(
                    ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                    capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._thread.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx.append(float(s[1]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(names, s[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------- reference arm / cpu baseline
def cpu_step(vec, u, mat, t):
    """One application on the host cores with the restated reference (oracle/cref.py)."""
    from oracle import cref

    out = cref.apply_orbital_rotation(vec, u, NORB, NELEC)
    return cref.apply_diag_coulomb_evolution(out, mat, t, NORB, NELEC, copy=False)


def time_cpu(steps: int, warmup: int):
    # torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm is supposed to use all host cores
    # (the reference sizes its rayon pool by RAYON_NUM_THREADS, default = all cores)
    if "RAYON_NUM_THREADS" not in os.environ:
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    from oracle import cref

    vec, u, mat, t = make_inputs()
    for _ in range(warmup):
        cpu_step(vec, u, mat, t)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_step(vec, u, mat, t)
    dt = time.perf_counter() - t0
    return steps / dt, dt / steps, cref.n_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    value, sec_per_step, cores = time_cpu(args.steps, min(args.warmup, 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": sec_per_step * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "norb": NORB, "nelec": list(NELEC)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "full workload per step: the restated reference (oracle/cref.py: reference Python "
                                   "drivers over a C/OpenMP restatement of its Rust kernels; ffsim itself cannot be "
                                   "built here: no cargo/pyscf)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def ncu_traffic(kernel: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, from the newest committed
    `ncu --set full` capture of this workload (profiles/*_ncu_metrics.csv); None when there is none."""
    import csv
    import glob

    # newest = last by name (r1_..., r1s4_..., r2_...): file times do not survive the copy to the GPU box
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*c2*_ncu_metrics.csv")))
    if not files:
        return None, None
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    try:
        rows = list(csv.reader(open(files[-1])))
        hdr, units = rows[0], rows[1]
        ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        vals = [float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]] for r in rows[2:] if kernel in r[0]]
        return (sum(vals) / len(vals), os.path.relpath(files[-1], ROOT)) if vals else (None, None)
    except Exception:
        return None, None


# --------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    import ffsim_b200 as ffsim
    from ffsim_b200 import _lib
    from ffsim_b200.gates.orbital_rotation import get_plan

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    vec_h, u, mat, t = make_inputs()
    dim = vec_h.size
    ffsim.init_cache(NORB, NELEC)
    vec_d = torch.from_numpy(vec_h).cuda()
    state_bytes = dim * 16

    def step_device():
        ffsim.apply_orbital_rotation(vec_d, u, NORB, NELEC, copy=False)
        ffsim.apply_diag_coulomb_evolution(vec_d, mat, t, NORB, NELEC, copy=False)

    # pinned host buffer for the end-to-end leg (the public API with a NumPy array)
    pinned = torch.empty(dim, dtype=torch.complex128, pin_memory=True)
    pinned.copy_(torch.from_numpy(vec_h))
    vec_pinned = pinned.numpy()

    def step_e2e_api():
        # the call a user makes: NumPy in, NumPy out, one upload and one download per step
        work, kind = ffsim._device.to_device(vec_pinned, copy=True)
        ffsim.apply_orbital_rotation(work, u, NORB, NELEC, copy=False)
        ffsim.apply_diag_coulomb_evolution(work, mat, t, NORB, NELEC, copy=False)
        return ffsim._device.from_device(work, kind)

    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()

    with ClockSampler(local_rank) as clocks:
        _lib.profile_begin()
        barrier()
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        for _ in range(args.steps):
            step_device()
        stop.record()
        barrier()
        elapsed_ms = start.elapsed_time(stop)
        prof = _lib.profile_end()

        # end-to-end leg: host buffers, copies inside the timed region
        result = None
        for _ in range(3):  # same buffer lifetime as the timed loop (the previous result is alive
            result = step_e2e_api()  # while the next one is produced): warms both pooled host buffers
        barrier()
        t0 = time.perf_counter()
        e2e_steps = max(3, min(args.steps, 10))
        for _ in range(e2e_steps):
            result = step_e2e_api()
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        barrier()

    t_all = torch.tensor([elapsed_ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_all, op=dist.ReduceOp.MAX)
    elapsed_ms, e2e_ms = float(t_all[0]), float(t_all[1])

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "B200_PROFILING.md fallback 6650"
        traffic, traffic_src = ncu_traffic("fused_pass_kernel")
        fused = prof["fused_pass_kernel"]
        launches_timed = max(fused["timed"], 1)
        achieved = fused["bytes"] / launches_timed / (fused["ms"] / launches_timed * 1e-3) / 1e9 if fused["ms"] > 0 else 0.0
        diag = prof["diag_kernel"]
        diag_gbs = diag["bytes"] / (diag["ms"] * 1e-3) / 1e9 if diag["ms"] > 0 else 0.0
        plan = get_plan(NORB, NELEC, u, u)
        n_launch = sum(v["launches"] for v in prof.values())
        # second ceiling of the fused kernel: the FP64 pipe.  One rotation on one amplitude pair is
        # 4 DMUL + 8 DFMA; a rotation of one spin touches C(norb-2, nocc-1) * dim_other pairs.
        import math

        from ffsim_b200.linalg import givens_decomposition
        n_rot = len(givens_decomposition(u)[0])
        pairs = math.comb(NORB - 2, NELEC[0] - 1) * math.comb(NORB, NELEC[1])
        dfma_per_launch = 12.0 * n_rot * pairs
        fp64_peak = float(peaks.get("fp64_tflops", 36.8))  # scripts/micro/fp64_peak.cu on this pool: profiles/r1_fp64_peak.jsonl
        fp64_tflops = 2.0 * dfma_per_launch / (fused["ms"] / launches_timed * 1e-3) / 1e12 if fused["ms"] > 0 else 0.0
        hbm_floor_ms = fused["bytes"] / launches_timed / (peak * 1e9) * 1e3
        fp64_floor_ms = 2.0 * dfma_per_launch / (fp64_peak * 1e12) * 1e3
        # CPU baseline beside it: at N=1 only (contract), a bounded sample of the same workload
        cpu_value, cpu_sec, cores = time_cpu(3, 1) if world == 1 else (None, None, None)
        line = {
            "metric": METRIC,
            "value": world * args.steps / (elapsed_ms * 1e-3),
            "unit": UNIT,
            "n_gpus": world,
            "steps": args.steps,
            "warmup": max(args.warmup, 3),
            "ms_per_step": elapsed_ms / args.steps,
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "norb": NORB, "nelec": list(NELEC), "state_bytes": state_bytes,
                       "l2": "state (305 MB) is larger than the 126 MB L2; no flush needed",
                       "parallelism": "single GPU" if world == 1 else f"{world} independent replicas (one per GPU)",
                       "plan": plan.describe()},
            "clocks": clocks.summary(),
            "e2e": {"value": world * e2e_steps / (e2e_ms * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": state_bytes, "d2h_bytes_per_step": state_bytes,
                    "ms_per_step": e2e_ms / e2e_steps,
                    "path": "ffsim_b200 public ops on a pinned NumPy array: upload, apply_orbital_rotation, "
                            "apply_diag_coulomb_evolution, download"},
            "gpu_launches": n_launch,
            "roofline": {"bound": "hbm", "kernel": "fused_pass_kernel", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": peak_src,
                         "launches": fused["launches"], "avg_launch_ms": fused["ms"] / launches_timed,
                         "algorithmic_bytes_per_launch": fused["bytes"] / launches_timed,
                         "note": "32 B per amplitude per launch; each launch fuses all n(n-1)/2 Givens rotations "
                                 "+ phases of one spin sector into one HBM round trip, which makes the FP64 pipe "
                                 "(not HBM) the binding ceiling: see fp64 below and DESIGN.md 4.1",
                         "fp64": {"rotations_per_launch": n_rot, "dfma_pipe_ops_per_launch": dfma_per_launch,
                                  "achieved_tflops": fp64_tflops, "peak_tflops": fp64_peak,
                                  "peak_source": "measured DFMA throughput, scripts/micro/fp64_peak.cu "
                                                 "(profiles/r1_fp64_peak.jsonl)",
                                  "frac": fp64_tflops / fp64_peak, "floor_ms": fp64_floor_ms},
                         "hbm_floor_ms": hbm_floor_ms,
                         "frac_of_binding_roofline": max(hbm_floor_ms, fp64_floor_ms) / (fused["ms"] / launches_timed)
                         if fused["ms"] > 0 else 0.0,
                         "diag_kernel": {"achieved": diag_gbs, "frac": diag_gbs / peak,
                                         "avg_launch_ms": diag["ms"] / max(diag["timed"], 1)},
                         "step_algorithmic_TBps": 96.0 * dim / (elapsed_ms / args.steps * 1e-3) / 1e12},
        }
        if world == 1:
            line["cpu_baseline"] = {"value": cpu_value, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": "3 full applications of the same workload with the restated reference "
                                              "(oracle/cref.py over oracle/c/ref_kernels.c, OpenMP, all host cores)"}
        assert np.isfinite(result).all()
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", type=str, default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        if args.steps > 5:
            args.steps = 5  # bounded: ~2 s of host work per step
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
