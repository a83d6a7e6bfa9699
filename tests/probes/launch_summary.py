"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py: per-kernel count / total time / share
over the LAST training step (a step starts at its csr_batch_gather launch; the launch after the last step belongs to
bench.py's own structure query). usage: launch_summary.py launches.csv"""
import csv, sys
from collections import defaultdict
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1e3 if unit in ("ns", "nsecond") else v * (1e3 if unit in ("ms", "msecond") else 1.0)
    rows.append((r["Kernel Name"], us))
starts = [i for i, (k, _) in enumerate(rows) if "csr_batch_gather_kernel" in k]
win = rows[starts[-2]:starts[-1]]
acc = defaultdict(lambda: [0, 0.0])
for k, us in win:
    k = k.split("(")[0].replace("<unnamed>::", "").replace("void ", "")[:72]
    acc[k][0] += 1; acc[k][1] += us
tot = sum(v[1] for v in acc.values())
own = sum(v[1] for k, v in acc.items() if not (k.startswith("at::") or k.startswith("native::") or "cutlass" in k or "cublas" in k))
n_own = sum(v[0] for k, v in acc.items() if not (k.startswith("at::") or k.startswith("native::") or "cutlass" in k or "cublas" in k))
print("# launches in the last step: %d (%d libgnm, %d torch), total %.1f us (cold-cache, serialised: compare SHARES); "
      "libgnm share %.1f%%" % (len(win), n_own, len(win) - n_own, tot, 100 * own / tot))
print("%-72s %6s %12s %7s" % ("kernel", "count", "total_us", "share"))
for k, v in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    print("%-72s %6d %12.1f %6.1f%%" % (k, v[0], v[1], 100 * v[1] / tot))
