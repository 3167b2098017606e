#!/usr/bin/env python
"""Summarise `ncu --page source --csv` of one kernel: executed warp instructions by opcode and the
hottest SASS lines by stall samples.  usage: ncu_source_summary.py <rep> <kernel-regex> [top] [which]
(`which`: index of the captured launch among those the regex matches, default 0)"""
import csv
import subprocess
import sys
from collections import Counter

rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# several kernels may follow each other: take the block asked for
which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
hdr_i = [i for i, r in enumerate(rows) if r and r[0] == "Address"][which]
hdr = rows[hdr_i]
body = []
for r in rows[hdr_i + 1:]:
    if not r or r[0] in ("Kernel Name", "Address"):
        break
    body.append(r)
ci = {n: hdr.index(n) for n in ("Source", "# Samples", "Instructions Executed", "Thread Instructions Executed")}
stall_cols = [i for i, n in enumerate(hdr) if n.startswith("stall_") and "Not Issued" not in n]
tot_inst = sum(int(r[ci["Instructions Executed"]]) for r in body)
tot_samp = sum(int(r[ci["# Samples"]]) for r in body)
print(f"{len(body)} SASS lines, {tot_inst} warp instructions executed, {tot_samp} samples")
ops = Counter()
samp = Counter()
for r in body:
    op = r[ci["Source"]].split()[0]
    if op.startswith("@"):
        op = r[ci["Source"]].split()[1]
    op = op.split(".")[0]
    ops[op] += int(r[ci["Instructions Executed"]])
    samp[op] += int(r[ci["# Samples"]])
print("by opcode (warp instr, share, samples share):")
for op, n in ops.most_common(22):
    print(f"  {op:10s} {n:12d} {100*n/tot_inst:5.1f}%  samples {100*samp[op]/max(tot_samp,1):5.1f}%")
stalls = Counter()
for r in body:
    for i in stall_cols:
        stalls[hdr[i]] += int(r[i] or 0)
print("stall reasons:", ", ".join(f"{k}={v}" for k, v in stalls.most_common(8)))
print("hottest lines by samples:")
for r in sorted(body, key=lambda r: -int(r[ci["# Samples"]]))[:top]:
    print(f"  {int(r[ci['# Samples']]):7d} {int(r[ci['Instructions Executed']]):10d}  {r[ci['Source']].strip()[:90]}")
