#!/usr/bin/env python3
"""Static SASS statistics of one kernel of libleb200.so: python scripts/sass_stats.py SUBSTRING [-l] (opcode histogram, spills; -l lists the code)"""
import collections, re, subprocess, sys
so = "lammps_le_b200/libleb200.so"
want = sys.argv[1]
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
cur, funcs = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1); funcs[cur] = []; continue
    m = re.match(r"\s*/\*([0-9a-f]{4})\*/\s+(.*?);", line)
    if m and cur:
        funcs[cur].append(m.group(2).strip())
for name, code in funcs.items():
    if want not in name:
        continue
    ops = collections.Counter()
    for ins in code:
        t = ins.split()
        op = t[1] if t[0].startswith("@") else t[0]
        ops[op.split(".")[0]] += 1
    print("%s: %d instructions" % (name, len(code)))
    print("  " + "  ".join("%s %d" % kv for kv in ops.most_common(30)))
    if "-l" in sys.argv:
        for k, ins in enumerate(code):
            print("%4d  %s" % (k, ins))
