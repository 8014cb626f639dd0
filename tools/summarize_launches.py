"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and
share of the device time (cold-cache, serialised launches: compare SHARES, not absolutes)."""
import csv
import collections
import sys

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
    rows.append((r["Kernel Name"], v))
# optional: keep only the launches from the LAST occurrence of a marker kernel on (e.g. the start of the last step)
if len(sys.argv) > 2:
    marks = [i for i, (n, _) in enumerate(rows) if sys.argv[2] in n]
    if marks:
        rows = rows[marks[-1]:]
        print("(from the last '%s' launch on: one step)" % sys.argv[2])
agg = collections.OrderedDict()
for name, us in rows:
    short = name.split("(")[0][:90]
    c, t = agg.get(short, (0, 0.0))
    agg[short] = (c + 1, t + us)
total = sum(t for _, t in agg.values())
print("launches %d   total %.1f us" % (len(rows), total))
OURS = ("tail::", "tc::", "tailtc::", "upconv::", "spctc::", "train::")                  # ncu drops the outer namespace of nested ones
ours = sum(t for k, (c, t) in agg.items() if "riders" in k or any(ns in k for ns in OURS))
print("our kernels (namespace riders): %.1f us = %.1f %% of device time" % (ours, 100 * ours / max(total, 1e-9)))
print("%-92s %7s %12s %7s" % ("kernel", "count", "total_us", "share"))
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-92s %7d %12.1f %6.1f%%" % (k, c, t, 100 * t / max(total, 1e-9)))
