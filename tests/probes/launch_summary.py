"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count / total time / share over
the LAST training step (window = launches between the last two dgi_score_fwd_kernel launches, ending at the last
launch). usage: launch_summary.py launches.csv"""
import csv, sys
from collections import defaultdict
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1e3 if unit in ("ns", "nsecond") else v * (1e3 if unit in ("ms", "msecond") else 1.0)
    rows.append((r["Kernel Name"], us))
idx = [i for i, (k, _) in enumerate(rows) if "dgi_score_fwd_kernel" in k]
period = idx[-1] - idx[-2]
# a step starts at its csr_batch_gather launch
starts = [i for i, (k, _) in enumerate(rows) if "csr_batch_gather_kernel" in k]
lo = max(s for s in starts if s <= idx[-1])
win = rows[lo:lo + period]
acc = defaultdict(lambda: [0, 0.0])
for k, us in win:
    k = k.split("(")[0].replace("<unnamed>::", "")[:72]
    acc[k][0] += 1; acc[k][1] += us
tot = sum(v[1] for v in acc.values())
print("# launches in the window: %d, total %.1f us (cold-cache, serialised: compare SHARES)" % (len(win), tot))
print("%-72s %6s %12s %7s" % ("kernel", "count", "total_us", "share"))
for k, v in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    print("%-72s %6d %12.1f %6.1f%%" % (k, v[0], v[1], 100 * v[1] / tot))
