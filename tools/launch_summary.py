#!/usr/bin/env python
"""Summarises an `ncu --metrics gpu__time_duration.sum --csv --log-file X.csv` launch list: per-kernel totals and the share of
each of OUR kernels inside a convolver step.   usage: python tools/launch_summary.py launches.csv "<command that was profiled>" """
import collections
import csv
import sys


def main():
    path, cmd = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            v = float(r["Metric Value"].replace(",", ""))
            unit = r["Metric Unit"]
            scale = {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1.0)
            rows.append((r["Kernel Name"], v * scale))
    tot = collections.Counter()
    cnt = collections.Counter()
    for k, v in rows:
        tot[k] += v
        cnt[k] += 1
    total = sum(tot.values())
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none: {cmd}")
    print(f"# per-kernel totals over the {len(rows)} captured launches (cold-cache, serialised: compare shares). unit=ns")
    for k, v in tot.most_common():
        print(f"{v:14.1f}  {100 * v / total:5.1f}%  x{cnt[k]:<4d} {k[:150]}")
    ours = {k: v for k, v in tot.items() if "neo_b200::" in k and "partition_r2c_io" not in k and "frame_filter_io" not in k
            and "channel_energy" not in k and "scale_by_min" not in k}
    t2 = sum(ours.values())
    print("\n# share inside one convolver step (our per-step kernels only; filter preparation excluded):")
    for k, v in sorted(ours.items(), key=lambda kv: -kv[1]):
        print(f"{100 * v / t2:5.1f}%  x{cnt[k]:<4d} {k[:150]}")


if __name__ == "__main__":
    main()
