"""Benchmark of the hot path (contract: see the task statement).

Step = one application of the hot path to one state: a random orbital rotation (both spins) followed
by a diagonal Coulomb evolution.  The state is the largest single-GPU configuration of BASELINE.json
(configs[2]'s state: norb=18, nelec=(7,7), 1.013e9 amplitudes, 16.2 GB) at every N, so that

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config c3|c2]

measures STRONG scaling of the sharded path: at N > 1 (torchrun, one rank per GPU) the ONE state is
distributed over the ranks (ffsim_b200/distributed.py: alpha rows or beta columns per rank, one
all-to-all over NVLink per two-spin rotation), at N = 1 it is a plain CUDA tensor.  The N = 1 line also
carries a second, complete record for BASELINE configs[1] (norb=16, nelec=(5,5), 305 MB) under "c2".
Times are CUDA-event times on the launching stream, max over ranks.
"""

from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

CONFIGS = {
    "c3": {"norb": 18, "nelec": (7, 7),
           "workload": "random orbital rotation + diag-Coulomb evolution on a random state, norb=18 nelec=(7,7), "
                       "1.013e9 amplitudes (16.2 GB): the state of BASELINE configs[2], the largest that fits one GPU"},
    "c2": {"norb": 16, "nelec": (5, 5),
           "workload": "random orbital rotation + diag-Coulomb evolution on a random state, norb=16 nelec=(5,5), "
                       "19.1M amplitudes (305 MB)"},
}
METRIC = "orbital-rotation + diag-Coulomb applications/sec"
UNIT = "applications/s"
SEED = 1803


def operators(norb: int):
    """The step's operators (one generator threaded through in this order)."""
    from ffsim_b200 import random as frandom  # bit-identical to the reference's generators (tests/golden)

    rng = np.random.default_rng(SEED)
    u = frandom.random_unitary(norb, seed=rng)
    mat = frandom.random_real_symmetric_matrix(norb, seed=rng)
    return u, mat, 1.0


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


# --------------------------------------------------------------------------- CPU arm (the restated reference)
class CpuSample:
    """A bounded sample of one step on the host cores, with the restated reference (oracle/cref.py: the
    reference's Python drivers over a C/OpenMP restatement of its Rust kernels; ffsim itself cannot be
    built here: no cargo, no pyscf).  The reference's step is, in this order: n(n-1)/2 Givens calls + n
    phase calls on the alpha index, a transposed copy, the same on the beta index, a transposed copy back,
    one diagonal Coulomb sweep -- every one of them a loop over independent columns (rows for the
    diagonal sweep).  The sample runs exactly these calls on ``cols`` of the dim columns (rows), i.e. the
    fraction cols/dim of the step's work; the full step is the sample time divided by that fraction."""

    def __init__(self, cfg, cols: int):
        if "RAYON_NUM_THREADS" not in os.environ:  # torchrun exports OMP_NUM_THREADS=1; the reference's
            os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)  # rayon pool takes all cores
        from oracle import cistring, cref, gates, givens

        self.cref, self.gates = cref, gates
        self.norb, self.nelec = cfg["norb"], cfg["nelec"]
        self.dim_a, self.dim_b = math.comb(self.norb, self.nelec[0]), math.comb(self.norb, self.nelec[1])
        self.cols = min(cols, self.dim_a, self.dim_b)
        u, mat, t = operators(self.norb)
        self.decomp = givens.givens_decomposition(u)
        self.mats = [np.ascontiguousarray(m, dtype=complex) for m in gates.get_mat_exp(mat, t, self.norb, False)]
        self.occ_a = np.ascontiguousarray(cistring.gen_occslst(range(self.norb), self.nelec[0]), dtype=np.uint64)
        self.occ_b = np.ascontiguousarray(cistring.gen_occslst(range(self.norb), self.nelec[1]), dtype=np.uint64)
        rng = np.random.default_rng(7)
        self.slab = rng.standard_normal((self.dim_a, self.cols)) + 1j * rng.standard_normal((self.dim_a, self.cols))
        self.rows = rng.standard_normal((self.cols, self.dim_b)) + 1j * rng.standard_normal((self.cols, self.dim_b))
        self.fraction = self.cols / self.dim_b
        self.cores = cref.n_threads()

    def step(self) -> None:
        cref = self.cref
        a = self.slab.copy()                                             # copy=True of the public call
        cref._rotate_one_spin(a, self.decomp, self.norb, self.nelec[0])  # alpha side of `cols` columns
        bt = cref._transpose(self.rows)                                  # transposed copy: [dim_b x cols]
        cref._rotate_one_spin(bt, self.decomp, self.norb, self.nelec[1]) # beta side of `cols` alpha rows
        back = cref._transpose(bt)                                       # and back: [cols x dim_b]
        aa, ab, bb = self.mats
        cref.lib().ref_apply_diag_coulomb_evolution_in_place_num_rep(
            cref._p(back), self.cols, self.dim_b, cref._p(aa), cref._p(ab), cref._p(bb), self.norb,
            cref._p(self.occ_a), self.nelec[0], cref._p(self.occ_b), self.nelec[1])

    def describe(self) -> str:
        return (f"{self.cols} of {self.dim_b} columns (fraction {self.fraction:.4f}) of every call of the step: "
                f"alpha rotation on [dim_a x {self.cols}], transposed copies, beta rotation on [dim_b x {self.cols}], "
                f"diagonal Coulomb sweep of {self.cols} rows; oracle/cref.py over oracle/c/ref_kernels.c "
                f"(OpenMP, {self.cores} threads); full-step time = sample time / fraction")


def time_cpu(cfg, steps: int, warmup: int, cols: int):
    sample = CpuSample(cfg, cols)
    for _ in range(warmup):
        sample.step()
    t0 = time.perf_counter()
    for _ in range(steps):
        sample.step()
    dt = (time.perf_counter() - t0) / steps
    full = dt / sample.fraction
    return 1.0 / full, full, sample


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    # every step is a bounded sample (see CpuSample): ~1.7 s of host work at c3, so that the driver's
    # --steps 20 --warmup 5 finishes within a minute; the full C2 step is small enough to run whole
    cols = 1024 if args.config == "c3" else math.comb(cfg["norb"], cfg["nelec"][1])
    value, sec_per_step, sample = time_cpu(cfg, args.steps, args.warmup, cols)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec_per_step * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": cfg["workload"], "norb": cfg["norb"], "nelec": list(cfg["nelec"])},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": sample.cores, "kind": "port",
                         "sample": sample.describe()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def ncu_traffic(kernel: str, tag: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, from the newest committed
    `ncu --set full` capture of this workload (profiles/*<tag>*_ncu_metrics.csv); None when there is none."""
    import csv
    import glob

    # newest = last by name (r1_..., r1s4_..., r2_...): file times do not survive the copy to the GPU box
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", f"*{tag}*_ncu_metrics.csv")))
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
def roofline_block(prof, peaks, fp64_peak, tag, state_bytes, ms_per_step):
    """Roofline of the dominant kernel from the library's own CUDA-event records of the timed region."""
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "B200_PROFILING.md fallback 6650"
    fused = prof["fused_pass_kernel"]
    n = max(fused["timed"], 1)
    avg_ms = fused["ms"] / n
    hbm_gbs = fused["bytes"] / n / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
    tflops = 2.0 * fused["dfma_ops"] / n / (avg_ms * 1e-3) / 1e12 if avg_ms > 0 else 0.0
    hbm_floor = fused["bytes"] / n / (peak * 1e9) * 1e3
    fp64_floor = 2.0 * fused["dfma_ops"] / n / (fp64_peak * 1e12) * 1e3
    traffic, traffic_src = ncu_traffic("fused_pass_kernel", tag)
    diag = prof["diag_kernel"]
    diag_gbs = diag["bytes"] / (diag["ms"] * 1e-3) / 1e9 if diag["ms"] > 0 else 0.0
    tr = prof["transpose_kernel"]
    tr_gbs = tr["bytes"] / (tr["ms"] * 1e-3) / 1e9 if tr["ms"] > 0 else 0.0
    fp64_bound = fp64_floor >= hbm_floor
    block = {
        "bound": "fp64" if fp64_bound else "hbm", "kernel": "fused_pass_kernel",
        "achieved": tflops if fp64_bound else hbm_gbs, "peak": fp64_peak if fp64_bound else peak,
        "unit": "TFLOP/s" if fp64_bound else "GB/s",
        "frac": (tflops / fp64_peak) if fp64_bound else (hbm_gbs / peak),
        "traffic": traffic, "traffic_source": traffic_src,
        "peak_source": ("dense DFMA throughput measured in this run by ffb_measure_fp64_peak (2 flops per DFMA)"
                        if fp64_bound else peak_src),
        "launches": fused["launches"], "avg_launch_ms": avg_ms,
        "algorithmic_bytes_per_launch": fused["bytes"] / n,
        "dfma_pipe_ops_per_launch": fused["dfma_ops"] / n,
        "share_of_step": fused["ms"] / max(prof["_steps"], 1) / ms_per_step if ms_per_step > 0 else None,
        "note": "a launch is one sweep over the state (32 B per amplitude) that applies every Givens rotation of "
                "its pass: 4 DMUL + 8 DFMA per rotation and amplitude pair.  The FP64 floor of a sweep is above "
                "its HBM floor, so the FP64 pipe is the binding roofline; the HBM fraction is given beside it",
        "fp64": {"achieved_tflops": tflops, "peak_tflops": fp64_peak, "frac": tflops / fp64_peak, "floor_ms": fp64_floor},
        "hbm": {"achieved_gbs": hbm_gbs, "peak_gbs": peak, "frac": hbm_gbs / peak, "floor_ms": hbm_floor,
                "peak_source": peak_src},
        "diag_kernel": {"achieved_gbs": diag_gbs, "frac": diag_gbs / peak,
                        "avg_launch_ms": diag["ms"] / max(diag["timed"], 1)},
        "transpose_kernel": {"achieved_gbs": tr_gbs, "frac": tr_gbs / peak, "launches": tr["launches"],
                             "avg_launch_ms": tr["ms"] / max(tr["timed"], 1)},
        "step_algorithmic_TBps": 6.0 * state_bytes / (ms_per_step * 1e-3) / 1e12 if ms_per_step > 0 else None,
    }
    return block


def run_config(cfg, tag, args, world, rank, local_rank, clocks_holder, with_cpu):
    """One complete record (device-timed value, e2e, roofline, optional cpu_baseline) for one configuration."""
    import torch
    import torch.distributed as dist

    import ffsim_b200 as ffsim
    from ffsim_b200 import _lib, distributed
    from ffsim_b200.distributed import ShardedVector
    from ffsim_b200.gates.orbital_rotation import get_plan

    norb, nelec = cfg["norb"], cfg["nelec"]
    dim_a, dim_b = math.comb(norb, nelec[0]), math.comb(norb, nelec[1])
    dim = dim_a * dim_b
    state_bytes = dim * 16
    dev = torch.device("cuda", local_rank)
    u, mat, t = operators(norb)
    ffsim.init_cache(norb, nelec)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # synthetic state: every rank draws its own rows on the device, then the whole is normalised
    a_off = distributed.partition(dim_a, world)
    n_rows = a_off[rank + 1] - a_off[rank]
    gen = torch.Generator(device=dev)
    gen.manual_seed(SEED + rank)
    local = torch.empty(n_rows * dim_b, dtype=torch.complex128, device=dev)
    torch.view_as_real(local).normal_(generator=gen)
    if world > 1:
        vec = ShardedVector(local, norb, nelec)
        local.mul_(1.0 / vec.norm())
    else:
        vec = local
        vec.mul_(1.0 / float(torch.linalg.vector_norm(vec)))
    shard_bytes = local.numel() * 16

    def step_device():
        ffsim.apply_orbital_rotation(vec, u, norb, nelec, copy=False)
        ffsim.apply_diag_coulomb_evolution(vec, mat, t, norb, nelec, copy=False)

    warm = max(args.warmup, 3)
    for _ in range(warm):
        step_device()
    barrier()
    distributed.STATS.update(exchanges=0, bytes_sent=0, events=[], time=True)
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
    prof["_steps"] = args.steps
    exch_ms = sum(a.elapsed_time(b) for a, b in distributed.STATS.get("events", []))
    n_exch, bytes_sent = distributed.STATS["exchanges"], distributed.STATS["bytes_sent"]
    distributed.STATS.update(time=False, events=[])
    clocks_holder.append(clocks.summary())
    norm_after = vec.norm() if world > 1 else float(torch.linalg.vector_norm(vec))

    # ---- end-to-end leg: the user's buffers are HOST memory (pinned); every step uploads its input,
    # runs the two public operations and downloads the result
    if world > 1:
        vec.set_layout(distributed.ROWS)
        src = vec.local
    else:
        src = vec
    host = torch.empty(src.numel(), dtype=torch.complex128, pin_memory=True)
    host.copy_(src)
    del src, vec, local
    torch.cuda.empty_cache()
    host_np = host.numpy()

    def step_e2e_plain():
        work = ffsim.to_device(host_np)
        if world > 1:
            sv = ShardedVector(work, norb, nelec)
            ffsim.apply_orbital_rotation(sv, u, norb, nelec, copy=False)
            ffsim.apply_diag_coulomb_evolution(sv, mat, t, norb, nelec, copy=False)
            sv.set_layout(distributed.ROWS)  # the caller's buffer holds rows: same distribution out as in
            return ffsim.to_host(sv.local)
        ffsim.apply_orbital_rotation(work, u, norb, nelec, copy=False)
        ffsim.apply_diag_coulomb_evolution(work, mat, t, norb, nelec, copy=False)
        return ffsim.to_host(work)

    def step_e2e_streamed():
        # one public call: column strips are uploaded while the alpha side rotates the strips that have arrived,
        # row blocks are downloaded while the beta side and the diagonal kernel finish the next ones
        return ffsim.evolve_host(host_np, [("orbital_rotation", u), ("diag_coulomb", mat, t)], norb, nelec)

    ops = [("orbital_rotation", u), ("diag_coulomb", mat, t)]
    e2e_steps = max(3, min(args.steps, 10 if state_bytes < (1 << 30) else 4))
    pipe_steps = max(4, min(args.steps, 10 if state_bytes < (1 << 30) else 8))

    def time_e2e(step):
        result = None
        for _ in range(2):
            del result
            result = step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            del result  # one result buffer alive at a time (it goes back to the pinned pool)
            result = step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        barrier()
        ok = bool(np.isfinite(result[:: max(1, result.size // 4096)]).all())
        return dt, ok, result

    def run_pipelined(n):
        # a stream of independent applications through the public asynchronous call: every step uploads its
        # input and downloads its result; the upload of step k+1 overlaps the kernels and the download of step k
        # (at N > 1 every rank does this with its block of alpha rows: evolve_host_rows_async, collective)
        start_one = ((lambda: ffsim.evolve_host_async(host_np, ops, norb, nelec)) if world == 1 else
                     (lambda: ffsim.evolve_host_rows_async(host_np, ops, norb, nelec)))
        pending, last = [], None
        for _ in range(n):
            if len(pending) >= 2:  # two device buffers: the third application waits for the first to come back
                last = None        # (its result buffer returns to the pinned pool)
                last = pending.pop(0).result()
            pending.append(start_one())
        while pending:
            last = None
            last = pending.pop(0).result()
        return last

    e2e_plain_s, finite, result = time_e2e(step_e2e_plain)
    single_s = None
    stride = max(1, result.size // 65536)
    check = result[::stride].copy()
    del result
    same = 0.0
    if world == 1:
        single_s, finite2, result = time_e2e(step_e2e_streamed)
        # both paths apply the same operators to the same host buffer
        same = float(np.linalg.norm(result[::stride] - check) / max(np.linalg.norm(check), 1e-300))
        finite = finite and finite2
        del result
    run_pipelined(4)  # fills the device ring / the symmetric-memory pool and the pinned result pool
    barrier()
    t0 = time.perf_counter()
    result = run_pipelined(pipe_steps)
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) * e2e_steps / pipe_steps  # scaled to e2e_steps like the other legs
    barrier()
    if check.size:
        same = max(same, float(np.linalg.norm(result[::stride] - check) / max(np.linalg.norm(check), 1e-300)))
    finite = finite and same < 1e-12
    ffsim.release_device_buffers()
    del result, host_np, host

    t_all = torch.tensor([elapsed_ms, e2e_s * 1e3, exch_ms, e2e_plain_s * 1e3], dtype=torch.float64, device=dev)
    single_ms = single_s * 1e3 / e2e_steps if single_s is not None else None
    if world > 1:
        dist.all_reduce(t_all, op=dist.ReduceOp.MAX)
    elapsed_ms, e2e_ms, exch_ms, e2e_plain_ms = (float(x) for x in t_all)
    if rank != 0:
        return None

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    fp64_peak = _lib.measure_fp64_peak()
    ms_per_step = elapsed_ms / args.steps
    plan = get_plan(norb, nelec, u, u)
    n_launch = sum(v["launches"] for k, v in prof.items() if isinstance(v, dict))
    record = {
        "metric": METRIC,
        "value": args.steps / (elapsed_ms * 1e-3),
        "unit": UNIT,
        "n_gpus": world,
        "steps": args.steps,
        "warmup": warm,
        "ms_per_step": ms_per_step,
        "higher_is_better": True,
        "scaling": "strong",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": cfg["workload"], "norb": norb, "nelec": list(nelec), "state_bytes": state_bytes,
                   "l2": f"the state ({state_bytes / 1e6:.0f} MB) is larger than the 126 MB L2; no flush needed",
                   "parallelism": "single GPU" if world == 1 else
                   f"one state sharded over {world} ranks (alpha rows / beta columns per rank; one all-to-all per "
                   f"two-spin rotation, the state stays in the distribution the last operation left it in)",
                   "plan": plan.describe(), "norm_after": norm_after, "finite": finite},
        "e2e": {"value": e2e_steps / (e2e_ms * 1e-3), "unit": UNIT,
                "h2d_bytes_per_step": shard_bytes * world if world > 1 else state_bytes,
                "d2h_bytes_per_step": shard_bytes * world if world > 1 else state_bytes,
                "ms_per_step": e2e_ms / e2e_steps, "steps": pipe_steps,
                "path": ("public API on pinned host memory: a stream of ffsim_b200.evolve_host_async(vec, [orbital_rotation, "
                         "diag_coulomb]) calls, at most two in flight, each result awaited -- every step uploads its "
                         "input (column strips, the alpha side rotates the strips that have arrived) and downloads its "
                         "result (row blocks, as the beta side and the diagonal kernel finish them); the upload of one "
                         "step overlaps the kernels and the download of the step before" if world == 1 else
                         "public API on pinned host memory: a stream of ffsim_b200.evolve_host_rows_async(rows, "
                         "[orbital_rotation, diag_coulomb]) calls on every rank (its block of alpha rows), at most two "
                         "in flight, each result awaited -- every step uploads the rank's shard, runs the two public "
                         "operations on a ShardedVector (one exchange over NVLink, the result brought back to the row "
                         "distribution) and downloads it; the upload of one step overlaps the kernels and the "
                         "download of the step before"),
                "pipelined_steps_timed": pipe_steps,
                "one_call_at_a_time_ms_per_step": single_ms,
                "one_call_at_a_time_path": ("ffsim_b200.evolve_host, each call awaited before the next starts "
                                            "(copies overlap the kernels of the same application only)"
                                            if world == 1 else None),
                "unstreamed_ms_per_step": e2e_plain_ms / e2e_steps,
                "unstreamed_path": "ffsim_b200.to_device, apply_orbital_rotation, apply_diag_coulomb_evolution, "
                                   "ffsim_b200.to_host (copies and kernels one after the other)",
                "streamed_vs_unstreamed_rel_diff": same},
        "gpu_launches": n_launch,
        "roofline": roofline_block(prof, peaks, fp64_peak, tag, state_bytes, ms_per_step),
    }
    if world > 1:
        nvlink_peak = 770.0  # GB/s per direction, B200_PROFILING.md (peer copy measured on this pool)
        sent_gbs = bytes_sent / (exch_ms * 1e-3) / 1e9 if exch_ms > 0 else 0.0
        fused, diag = prof["fused_pass_kernel"], prof["diag_kernel"]
        record["sharded"] = {
            "exchanges_per_step": n_exch / args.steps,
            "exchange": {"p2p": "peer-memory stores (ffb_copy_blocks over NVLink, symmetric memory)",
                         "nccl": "NCCL all_to_all_single + ffb_copy_blocks pack/unpack"}.get(
                             distributed.STATS.get("mode"), "none"),
            "per_rank_ms_per_step": {"rotation_kernels": fused["ms"] / args.steps, "diag_kernel": diag["ms"] / args.steps,
                                     "transposes": prof["transpose_kernel"]["ms"] / args.steps,
                                     "exchange": exch_ms / args.steps},
            "nvlink_bytes_sent_per_rank_per_step": bytes_sent / args.steps,
            "nvlink_achieved_gbs_per_rank": sent_gbs, "nvlink_peak_gbs": nvlink_peak,
            "nvlink_frac": sent_gbs / nvlink_peak,
            "limiting_term": max((("rotation kernels (FP64 pipe)", fused["ms"]), ("exchange (NVLink)", exch_ms),
                                  ("diagonal kernel (HBM)", diag["ms"]),
                                  ("transposes (HBM)", prof["transpose_kernel"]["ms"])), key=lambda kv: kv[1])[0],
        }
    if with_cpu:
        cols = 2048 if tag == "c3" else dim_b
        cpu_value, cpu_sec, sample = time_cpu(cfg, 3, 1, cols)
        record["cpu_baseline"] = {"value": cpu_value, "unit": UNIT, "cores": sample.cores, "kind": "port",
                                  "ms_per_step": cpu_sec * 1e3, "sample": "3 steps of " + sample.describe()}
    return record


def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    clocks = []
    line = run_config(CONFIGS[args.config], args.config, args, world, rank, local_rank, clocks, with_cpu=world == 1)
    if rank == 0:
        line["clocks"] = clocks[0]
    if world == 1 and args.config == "c3" and not args.no_c2:
        # second record, same contract, for BASELINE configs[1] (the round-1 headline)
        clocks2 = []
        sub = argparse.Namespace(**vars(args))
        sub.steps = max(args.steps, 20)
        rec = run_config(CONFIGS["c2"], "c2", sub, world, rank, local_rank, clocks2, with_cpu=True)
        rec["clocks"] = clocks2[0]
        line["c2"] = rec
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", type=str, default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=str, default="c3", choices=sorted(CONFIGS))
    ap.add_argument("--no-c2", action="store_true", help="skip the second (configs[1]) record at N=1")
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: whatever libraries print on file descriptor 1 meanwhile (NCCL
    # announces its version there) goes to stderr instead
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
    sys.stdout.flush()


if __name__ == "__main__":
    main()
