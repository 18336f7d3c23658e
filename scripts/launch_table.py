#!/usr/bin/env python3
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: python scripts/launch_table.py FILE"""
import csv, collections, statistics, sys
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith('==')]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    k = row['Kernel Name']; v = float(row['Metric Value'].replace(',', '')); u = row['Metric Unit']
    if u == 'ns': v /= 1000
    elif u == 'ms': v *= 1000
    a = agg.setdefault(k, [0, 0.0, []]); a[0] += 1; a[1] += v; a[2].append(v)
tot = sum(a[1] for a in agg.values())
print("total %.1f us over %d launches" % (tot, sum(a[0] for a in agg.values())))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-44s n=%5d sum=%10.1fus %5.1f%% med=%8.2f max=%8.2f" % (k[:44], a[0], a[1], 100 * a[1] / tot, statistics.median(a[2]), max(a[2])))
