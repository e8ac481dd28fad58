#!/usr/bin/env python
"""Region / stall breakdown of one kernel from `ncu -i X.ncu-rep --page source --csv --print-source sass`.
    python profiles/sass_regions.py src.csv WARPS_TIMES_STEPS [dump_from dump_to]
Prints the stall-reason mix over all samples, the instruction mix by opcode, and the instruction count /
sample share of the code between marker instructions (LDTM, UTCHMMA, BAR, STG, SYNCS, MUFU.EX2/LG2)."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) >= len(hdr)]
N = float(sys.argv[2])
S = sum(int(r[ix["# Samples"]]) for r in data)
names = [h for h in hdr if h.startswith("stall_") and "(" not in h]
tot = {n: sum(int(r[ix[n]]) for r in data) for n in names}
print("stall mix %:", {k[6:]: round(100 * v / S, 1) for k, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v})
ops = collections.Counter(); osamp = collections.Counter()
for r in data:
    src = r[ix["Source"]].strip().split()
    op = (src[1] if src[0].startswith("@") else src[0]).split(".")[0]
    ops[op] += int(r[ix["Instructions Executed"]]); osamp[op] += int(r[ix["# Samples"]])
T = sum(ops.values())
print(f"instructions per warp-step: {T / N:.1f}")
print("opcode mix:", ", ".join(f"{o} {n / N:.0f} ({100 * osamp[o] / S:.1f}%s)" for o, n in ops.most_common(22)))
marks = ("LDTM", "UTCHMMA", "BAR.SYNC", "STG", "SYNCS.PHASECHK", "UTCBAR", "MUFU.EX2", "MUFU.LG2", "UBLKCP", "MEMBAR")
acc = accs = 0
for i, r in enumerate(data):
    src = r[ix["Source"]].strip(); n = int(r[ix["Instructions Executed"]]); s = int(r[ix["# Samples"]])
    acc += n; accs += s
    if any(m in src for m in marks) and n / N > 0.01:
        print(f"{i:5d} +{acc / N:7.1f} instr +{100 * accs / S:5.1f}% samp | {n / N:5.2f}x {100 * s / S:5.2f}% {src[:80]}")
        acc = accs = 0
print(f"tail +{acc / N:.1f} instr +{100 * accs / S:.1f}% samp")
if len(sys.argv) > 4:
    for i in range(int(sys.argv[3]), int(sys.argv[4])):
        r = data[i]
        st = {n[6:]: int(r[ix[n]]) for n in names if int(r[ix[n]]) > 0}
        print(i, r[ix["Source"]].strip()[:72].ljust(72), r[ix["# Samples"]], st)
