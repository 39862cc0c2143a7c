"""ncu launch list (--metrics gpu__time_duration.sum --csv --log-file X) -> per-kernel totals and shares.
Usage: python tools/launch_list.py gpurun_out/<tag>_launches.csv "<header note>" > profiles/<tag>_launch_list.txt"""
import csv
import re
import sys
from collections import OrderedDict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [ln for ln in f if not ln.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ms = v / 1e6 if unit in ("ns", "nsecond") else v / 1e3 if unit in ("us", "usecond") else v
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("bgnn::", "")
        rows.append((name, ms))
tot = sum(ms for _, ms in rows)
agg = OrderedDict()
for name, ms in rows:
    a = agg.setdefault(name, [0.0, 0])
    a[0] += ms
    a[1] += 1
print("# ncu --metrics gpu__time_duration.sum --clock-control none -k regex:<this library's kernels>  python bench.py --steps 2 --warmup 3 --no-cpu-baseline")
print("# (the same command exited 0 without ncu just before: tools/evidence.sh).  Times are cold-cache, serialised: compare SHARES, not absolutes.")
if len(sys.argv) > 2:
    print("# " + sys.argv[2])
print("# %d launches, %.1f ms" % (len(rows), tot))
for name, (ms, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print("%10.3f ms %5.1f%% n=%5d avg=%9.4f ms  %s" % (ms, 100 * ms / tot, n, ms / n, name))
