#!/usr/bin/env python3
"""Per-instruction view of one kernel of an .ncu-rep (needs -lineinfo + --import-source): opcode mix weighted by
executions, warp-level instruction count, and the instructions that collect the most stall samples.
  python scripts/ncu_hotspots.py REPORT KERNEL_REGEX WARPS [TOP]"""
import collections, csv, io, subprocess, sys
rep, rx, warps = sys.argv[1], sys.argv[2], float(sys.argv[3])
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", "regex:k_"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
blocks, cur, hdr, name = [], None, None, None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = []; blocks.append((r[1], cur)); continue
    if r and r[0] == "Address":
        hdr = r; continue
    if cur is not None and r:
        cur.append(r)
blocks = [x for x in blocks if __import__("re").search(rx, x[0])]
name, b = blocks[0]
ie, at, sm = hdr.index("Instructions Executed"), hdr.index("Avg. Threads Executed"), hdr.index("# Samples")
tot = sum(int(r[ie]) for r in b); ts = sum(int(r[sm]) for r in b)
print("%s\n%d SASS instructions, %d warp-instructions executed = %.1f per warp (%g warps), %d stall samples" % (name, len(b), tot, tot / warps, warps, ts))
ops, smp = collections.Counter(), collections.Counter()
for r in b:
    t = r[1].strip().split()
    o = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    ops[o] += int(r[ie]); smp[o] += int(r[sm])
print("\nopcode        per warp   share  stall samples")
for o, n in ops.most_common(22):
    print("%-10s %10.1f  %5.1f%%  %5.1f%%" % (o, n / warps, 100.0 * n / tot, 100.0 * smp[o] / ts))
print("\nmost stalled instructions (index, executions per warp, avg active threads, share of stall samples)")
for k, r in sorted(enumerate(b), key=lambda kr: -int(kr[1][sm]))[:top]:
    print("%5d  %6.2f  %5s  %5.1f%%  %s" % (k, int(r[ie]) / warps, r[at], 100.0 * int(r[sm]) / ts, r[1].strip()))
