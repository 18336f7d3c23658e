#!/usr/bin/env python3
"""Per-source-line executed instructions and stall samples from an .ncu-rep: python scripts/ncu_lines.py REPORT [topN]"""
import csv, subprocess, sys, io
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(io.StringIO(raw)))
cur_file = None; hdr = None; out = []
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur_file = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': func = r[1]; continue
    if r[0] == 'Line No': hdr = r; continue
    if r[0] != '' and hdr:   # a source line summary row
        try:
            ie = hdr.index('Instructions Executed'); sm = hdr.index('# Samples')
            out.append((int(r[ie] or 0), int(r[sm] or 0), cur_file, r[0], r[1].strip()[:110], func))
        except (ValueError, IndexError): pass
ti = sum(o[0] for o in out) or 1; ts = sum(o[1] for o in out) or 1
print("total inst %d samples %d" % (ti, ts))
for o in sorted(out, key=lambda o: -o[1])[:top]:
    print("%5.1f%% inst %5.1f%% smp  %s:%s  %s" % (100 * o[0] / ti, 100 * o[1] / ts, o[2], o[3], o[4]))
