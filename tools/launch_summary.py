"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel
count, total and mean device time, share of the total."""
import csv
import sys
from collections import defaultdict


def main(path, skip=0):
    rows = []
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(unit, 1e-3)
        name = r["Kernel Name"]
        grid = r.get("Grid Size", "")
        rows.append((name, grid, v * scale))
    rows = rows[skip:]
    agg = defaultdict(lambda: [0, 0.0])
    for n, g, us in rows:
        key = n.split("(")[0][:60] + " " + g
        agg[key][0] += 1
        agg[key][1] += us
    tot = sum(v[1] for v in agg.values())
    print("launches %d  total %.1f us" % (len(rows), tot))
    for k, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print("%6.1f%%  n=%5d  mean %9.2f us  %s" % (100 * us / tot, c, us / c, k))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
