"""Per-source-line digest of an `ncu --set full --import-source on` report: stall samples, executed warp instructions,
dominant stall reasons, shared-memory wavefront excess.
usage: ncu_src_lines.py report.ncu-rep [top] [extra `ncu -i` filters, e.g. --kernel-name regex:aggregate_tc --launch-count 1]"""
import csv, io, subprocess, sys
from collections import defaultdict
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "cuda,sass"] + sys.argv[3:],
                     capture_output=True, text=True).stdout
path, hdr = None, None
lines = []
for r in csv.reader(io.StringIO(out)):
    if not r:
        continue
    if r[0] == "File Path":
        path = r[1].split("/")[-1]; continue
    if r[0] == "Line No":
        hdr = r; continue
    if r[0] in ("Function Name", "Kernel Name") or hdr is None or not r[0].strip():
        continue
    d = {}
    for k, v in zip(hdr, r):
        d.setdefault(k, v)
    try:
        smp = int(d["# Samples"]); ins = int(d["Instructions Executed"])
    except (ValueError, KeyError):
        continue
    stalls = {k[6:]: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v) > 0}
    exc = int(d.get("L1 Wavefronts Shared Excessive", "0") or 0); wf = int(d.get("L1 Wavefronts Shared", "0") or 0)
    lines.append((smp, ins, path, int(r[0]), r[1].strip()[:90], stalls, exc, wf))
ts = sum(l[0] for l in lines); ti = sum(l[1] for l in lines)
print("# %d stall samples, %d warp-instructions executed, shared wavefronts %d (excess %d)" %
      (ts, ti, sum(l[7] for l in lines), sum(l[6] for l in lines)))
for smp, ins, p, ln, src, stalls, exc, wf in sorted(lines, key=lambda l: -l[0])[:top]:
    main = ", ".join("%s %d" % kv for kv in sorted(stalls.items(), key=lambda kv: -kv[1])[:3])
    print("%5.1f%% smp %5.1f%% inst  %s:%d  %-90s | %s%s" % (100.0 * smp / max(ts, 1), 100.0 * ins / max(ti, 1), p, ln, src, main,
                                                           ("  smem wf %d (+%d)" % (wf, exc)) if wf else ""))
