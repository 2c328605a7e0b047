"""Summarise an ncu launch list (``--metrics gpu__time_duration.sum --csv``) per kernel: launches, total time, share."""
import csv
import re
import sys
from collections import OrderedDict


def main(path, command):
    rows = [l for l in open(path) if l.startswith('"')]
    rd = csv.DictReader(rows)
    tot = OrderedDict()
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*$", "", r["Kernel Name"]).strip()
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v
        n, t = tot.get(name, (0, 0.0))
        tot[name] = (n + 1, t + us)
    total = sum(t for _, t in tot.values())
    print(f"launch list of `{command}` under ncu (gpu__time_duration.sum, --clock-control none); per-launch times are cold-cache and serialised")
    print("kernel | launches | total us | share")
    for k, (n, t) in tot.items():
        print(f"{k} | {n} | {t:.1f} | {100 * t / total:.1f} %")
    print(f"total | {sum(n for n, _ in tot.values())} | {total:.1f} | 100 %")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "python bench.py --steps 2 --warmup 3 --no-cpu --no-secondary --e2e-steps 1")
