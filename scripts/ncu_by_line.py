"""Attribute an ncu SASS-page CSV to CUDA source lines using nvdisasm line info.

usage: ncu_by_line.py <sass.csv> <nvdisasm -g -c output> <kernel mangled substring>
"""
import collections
import csv
import re
import sys

csv_path, sass_path, kern = sys.argv[1:4]
rows = list(csv.reader(open(csv_path)))
start = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hdr = rows[start[0]]
end = start[1] - 1 if len(start) > 1 else len(rows)
data = [r for r in rows[start[0] + 1:end] if len(r) >= len(hdr)]
ci = {h: i for i, h in enumerate(hdr)}

# instruction -> line list from nvdisasm
lines = []
cur = None
active = False
inl = None
for ln in open(sass_path):
    if ln.startswith(".text."):
        active = kern in ln
        continue
    if not active:
        continue
    m = re.match(r'\s*//## File ".*?", line (\d+)(.*)', ln)
    if m:
        cur = int(m.group(1))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s", ln):
        lines.append(cur)
print("csv instr", len(data), "nvdisasm instr", len(lines), file=sys.stderr)
n = min(len(data), len(lines))
by_line_smp = collections.Counter()
by_line_inst = collections.Counter()
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
by_line_stall = collections.defaultdict(collections.Counter)
ts = ti = 0
for k in range(n):
    r = data[k]
    s = int(r[ci["# Samples"]] or 0)
    i = int(r[ci["Instructions Executed"]] or 0)
    by_line_smp[lines[k]] += s
    by_line_inst[lines[k]] += i
    ts += s
    ti += i
    for h in stall_cols:
        v = int(r[ci[h]] or 0)
        if v:
            by_line_stall[lines[k]][h[6:]] += v
src = open("/root/repo/ffsim_b200/csrc/givens_kernels.cu").read().split("\n")
print(f"total samples {ts} instructions {ti}")
for line, s in by_line_smp.most_common(int(sys.argv[4]) if len(sys.argv) > 4 else 28):
    top = ", ".join(f"{k}:{100*v/max(s,1):.0f}%" for k, v in by_line_stall[line].most_common(3))
    text = src[line - 1].strip()[:70] if line and line <= len(src) else ""
    print(f"{line:5d} smp {100*s/ts:5.1f}% inst {100*by_line_inst[line]/ti:5.1f}%  [{top}]  {text}")
