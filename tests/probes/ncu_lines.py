"""Aggregate an `ncu --page source --csv --print-source cuda` dump by source line: instructions executed and
stall samples per line, per kernel. usage: ncu_lines.py dump.csv [top]"""
import csv, sys
from collections import defaultdict
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(sys.argv[1])))
fn, path, hdr = None, None, None
acc = defaultdict(lambda: defaultdict(lambda: [0, 0, ""]))
for r in rows:
    if not r:
        continue
    if r[0] == "Function Name":
        fn = r[1].split("(")[0][-40:]; continue
    if r[0] == "File Path":
        path = r[1].split("/")[-1]; hdr = None; continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr is None or len(r) < len(hdr):
        continue
    try:
        line = int(r[0])
    except ValueError:
        continue
    d = dict(zip(hdr, r))
    try:
        ins = int(d.get("Instructions Executed", "0") or 0); smp = int(d.get("# Samples", "0") or 0)
    except ValueError:
        continue
    a = acc[fn][(path, line)]
    a[0] += ins; a[1] += smp; a[2] = r[1][:110]
for fn, lines in acc.items():
    ti = sum(v[0] for v in lines.values()); ts = sum(v[1] for v in lines.values())
    print("==== %s: %d warp-instructions, %d samples" % (fn, ti, ts))
    for (p, l), v in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%5.1f%% inst %5.1f%% smp  %s:%d  %s" % (100.0 * v[0] / max(ti, 1), 100.0 * v[1] / max(ts, 1), p, l, v[2].strip()))
