"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list (one row per kernel launch):
    python tools/launch_summary.py gpurun_out/launches_<model>.csv "<command that was profiled>" > profiles/r01_launches_<model>.txt
Per-launch times under ncu are cold-cache and serialised: compare SHARES of the captured time, not absolutes.
"""
import csv
import re
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    cmd = sys.argv[2] if len(sys.argv) > 2 else ""
    lines = [l for l in open(path, errors="replace").read().splitlines() if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    agg = OrderedDict()
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(anonymous namespace\)|tfimm::|void ", "", r["Kernel Name"])
        name = re.sub(r"\(.*$", "", name)[:72]
        us = float(r["Metric Value"].replace(",", ""))
        if r.get("Metric Unit", "us").startswith("ns"):
            us /= 1e3
        elif r.get("Metric Unit", "us").startswith("ms"):
            us *= 1e3
        key = (name, r.get("Grid Size", ""), r.get("Block Size", ""))
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += us
    total = sum(v[1] for v in agg.values()) or 1.0
    print(f"# {cmd}")
    print("# cold-cache, serialised launches: compare SHARES, not absolutes.  "
          "columns: kernel | grid | block | launches | avg us | share of captured time")
    for (name, grid, block), (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{name:72s} {grid:16s} {block:14s} {n:4d} {us / n:9.1f} {100 * us / total:6.1f}%")


if __name__ == "__main__":
    main()
