"""Turn one gpurun round (gpurun_out/<tag>_*) into the tracked evidence under profiles/.

usage: python scripts/summarize_profile.py <tag> [<out-name>]
Reads   gpurun_out/<tag>_launches.csv   (ncu --metrics gpu__time_duration.sum launch list)
        gpurun_out/<tag>_prof.ncu-rep   (ncu --set full capture of the top kernels)
        gpurun_out/<tag>_bench.json / _bench_ref.json
Writes  profiles/<out>_launches.csv, profiles/<out>_ncu_metrics.csv, profiles/<out>_summary.md
"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
out = sys.argv[2] if len(sys.argv) > 2 else tag
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
]


def to_ns(v, u):
    v = float(v.replace(",", ""))
    return v * {"ns": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "s": 1e9, "second": 1e9, "nsecond": 1}.get(u, 1)


md = [f"# ncu evidence `{out}`", ""]
bench = os.path.join(G, f"{tag}_bench.json")
if os.path.exists(bench):
    for path in (bench, os.path.join(G, f"{tag}_bench_ref.json")):
        if os.path.exists(path):
            txt = open(path).read().strip().splitlines()
            if txt:
                j = json.loads(txt[-1])
                keep = {k: j.get(k) for k in ("impl", "metric", "value", "unit", "ms_per_step", "e2e", "roofline", "cpu_baseline", "clocks", "gpu_launches") if k in j}
                md += [f"## bench line ({os.path.basename(path)}; NOT taken under a profiler)", "", "```json", json.dumps(keep, indent=1), "```", ""]

launches = os.path.join(G, f"{tag}_launches.csv")
if os.path.exists(launches):
    shutil.copy(launches, os.path.join(P, f"{out}_launches.csv"))
    rows = [r for r in csv.reader(open(launches)) if len(r) > 10]
    hdr = rows[0]
    iN, iV, iU = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        a = agg.setdefault(r[iN].split("(")[0][-60:], [0, 0.0])
        a[0] += 1
        a[1] += to_ns(r[iV], r[iU])
    tot = sum(v[1] for v in agg.values())
    md += ["## launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`, bench.py --steps 2 --warmup 3)",
           "", "Per-launch times under ncu are cold-cache and serialised; the SHARE is what must agree with the live bench.",
           "", "| kernel | launches | total ms | share | avg us |", "|---|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        md.append(f"| `{k}` | {v[0]} | {v[1] / 1e6:.3f} | {100 * v[1] / tot:.1f}% | {v[1] / v[0] / 1e3:.1f} |")
    md.append("")

rep = os.path.join(G, f"{tag}_prof.ncu-rep")
raw_csv = os.path.join(G, f"{tag}_raw.csv")  # extracted on the GPU box when the report is too large to bring back
if os.path.exists(rep) or os.path.exists(raw_csv):
    if os.path.exists(raw_csv) and os.path.getsize(raw_csv) > 0:
        raw = open(raw_csv).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    cols = [hdr.index("Kernel Name")] + [hdr.index(k) for k in KEEP if k in hdr]
    with open(os.path.join(P, f"{out}_ncu_metrics.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([hdr[c] for c in cols])
        w.writerow([units[c] for c in cols])
        for r in rows[2:]:
            w.writerow([r[c] for c in cols])
    md += ["## `ncu --set full --clock-control none --import-source on` (scripts/profile_target.py)", "",
           "| kernel | time | dram read | dram write | dram % | fp64 pipe % | issue active % | regs | warps active % |", "|---|---|---|---|---|---|---|---|---|"]
    ix = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        g = lambda k: r[ix[k]] if k in ix else "?"  # noqa: E731
        u = lambda k: units[ix[k]] if k in ix else ""  # noqa: E731
        md.append(f"| `{g('Kernel Name').split('(')[0][-40:]}` | {g('gpu__time_duration.sum')} {u('gpu__time_duration.sum')} | "
                  f"{g('dram__bytes_read.sum')} {u('dram__bytes_read.sum')} | {g('dram__bytes_write.sum')} {u('dram__bytes_write.sum')} | "
                  f"{float(g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')):.1f} | "
                  f"{float(g('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active')):.1f} | "
                  f"{float(g('smsp__issue_active.avg.pct_of_peak_sustained_active')):.1f} | {g('launch__registers_per_thread')} | "
                  f"{float(g('sm__warps_active.avg.pct_of_peak_sustained_active')):.1f} |")
    md.append("")
open(os.path.join(P, f"{out}_summary.md"), "w").write("\n".join(md) + "\n")
print("\n".join(md))
