#!/usr/bin/env python3
"""Per-source-line warp instructions (per warp) and stall-sample share of ONE captured launch:
python scripts/ncu_lines2.py REPORT LAUNCH_INDEX WARPS [MIN_PER_WARP]"""
import csv, io, subprocess, sys
rep, idx, warps = sys.argv[1], sys.argv[2], float(sys.argv[3])
mn = float(sys.argv[4]) if len(sys.argv) > 4 else 3.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", idx, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = None; out = []; cur = None
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': print(r[1]); continue
    if r[0] == 'Line No' or r[0] == '#': hdr = r; continue
    if hdr and 'Instructions Executed' in hdr:
        try:
            ie = hdr.index('Instructions Executed'); sm = hdr.index('# Samples')
            out.append((int(r[ie] or 0), int(r[sm] or 0), cur, int(r[0]), r[1].strip()[:105]))
        except Exception: pass
tot = 0; ts = sum(o[1] for o in out) or 1
for o in sorted(out, key=lambda o: (o[2], o[3])):
    if o[0] / warps >= mn: print("%6.1f %4.1f%% %s:%d %s" % (o[0] / warps, 100 * o[1] / ts, o[2][:9], o[3], o[4]))
    tot += o[0] / warps
print("total per warp %.1f" % tot)
