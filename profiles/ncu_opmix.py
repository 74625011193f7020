#!/usr/bin/env python
"""Opcode mix + hottest SASS lines of one kernel from an .ncu-rep: ncu_opmix.py <rep> <kernel-regex> [min_frac]"""
import csv, subprocess, sys, collections, io
rep, kre = sys.argv[1], sys.argv[2]
minfrac = float(sys.argv[3]) if len(sys.argv) > 3 else 0.3
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
data = []
for r in rows[2:]:
    if len(r) < 10:
        continue
    if r[0] in ("Address", "Kernel Name"):
        if data:
            break
        continue
    data.append(r)
isrc, iex, ist, iat = (hdr.index(k) for k in ("Source", "Instructions Executed", "Warp Stall Sampling (All Samples)", "Avg. Threads Executed"))
tot = sum(int(r[iex]) for r in data)
print("total warp instr", tot, "static", len(data))
ops = collections.Counter()
for r in data:
    t = r[isrc].split()
    op = t[1] if t and t[0].startswith("@") else (t[0] if t else "?")
    ops[op.split(".")[0]] += int(r[iex])
for k, v in ops.most_common(24):
    print("%-10s %6.2f%%" % (k, 100 * v / tot))
mx = max(int(r[iex]) for r in data)
# bucket static instructions by execution count
b = collections.Counter()
for r in data:
    b[int(r[iex])] += 1
print("exec-count buckets (count -> #static instrs, share of dynamic):")
for k, v in sorted(b.items(), key=lambda kv: -kv[0] * kv[1])[:12]:
    print("  %12d x %4d  %5.1f%%" % (k, v, 100.0 * k * v / tot))
if "--list" in sys.argv:
    for i, r in enumerate(data):
        if int(r[iex]) > minfrac * mx:
            print("%5d %12s %5s %7s  %s" % (i, r[iex], r[iat][:5], r[ist], r[isrc][:100]))
