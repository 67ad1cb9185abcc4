"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
    python profiles/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launch_summary.txt
Durations are cold-cache, serialised, at unconstrained clocks: compare SHARES, not absolutes (B200_PROFILING.md)."""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in csv.DictReader(lines):
        name = row.get("Kernel Name")
        if not name:
            continue
        t = float(row["Metric Value"].replace(",", ""))
        unit = row.get("Metric Unit", "ns")
        t = {"ns": t / 1e3, "us": t, "ms": t * 1e3, "s": t * 1e6}.get(unit, t / 1e3)
        short = re.sub(r"\(.*", "", name)
        short = re.sub(r"^void ", "", short)[:110]
        agg[short][0] += 1
        agg[short][1] += t
        tot += t
    n = sum(a[0] for a in agg.values())
    print(f"# {path}: {n} launches, {tot / 1e3:.2f} ms total (serialised)")
    mine = sum(t for k, (c, t) in agg.items() if k.startswith("rtti::"))
    print(f"# rtti_b200 kernels: {mine / 1e3:.2f} ms = {100 * mine / tot:.1f} % of the listed time")
    for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:45]:
        print(f"{t:10.1f} us {100 * t / tot:5.1f}%  n={c:4d}  {k}")


if __name__ == "__main__":
    main(sys.argv[1])
