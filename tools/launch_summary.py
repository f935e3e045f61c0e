#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel mean / share of the step."""
import collections
import csv
import sys


def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = row["Kernel Name"].split("(")[0][:70]
        v = float(row["Metric Value"])
        unit = row.get("Metric Unit", "ns")
        if unit in ("us", "usecond"):
            v *= 1000.0
        agg.setdefault(k, []).append(v)
    ours = {k: v for k, v in agg.items() if "clr::" in k or k.startswith("clr")}
    steps = min(len(v) for v in ours.values()) if ours else 1
    tot = sum(sum(v) / steps for v in ours.values())
    print("%-72s %4s %10s %7s" % ("kernel", "n", "mean_us", "share"))
    for k, v in agg.items():
        per_step = sum(v) / steps / 1000.0
        mark = per_step * 1000.0 / tot if k in ours else float("nan")
        print("%-72s %4d %10.2f %6.1f%%" % (k, len(v), sum(v) / len(v) / 1000.0, 100 * mark))
    print("sum of library kernels per step: %.1f us over %d steps" % (tot / 1000.0, steps))


if __name__ == "__main__":
    main(sys.argv[1])
