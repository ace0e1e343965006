"""Dev tool: per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections
import csv
import re
import sys


def main(path, per_step_marker="step_finalize", per_step_count=1):
    rows = []
    lines = [l for l in open(path) if not l.startswith("==")]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        v = v / 1000.0 if unit == "ns" else (v * 1000.0 if unit == "ms" else v)
        rows.append((r["Kernel Name"], v, r["Grid Size"], r["Block Size"]))
    agg = collections.OrderedDict()
    for n, v, g, b in rows:
        short = re.sub(r"<.*", "", n.split("(")[0])[:44]
        a = agg.setdefault(short, [0, 0.0, g, b])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    nsteps = max(1.0, sum(1 for n, *_ in rows if per_step_marker in n) / per_step_count)
    print("%s: %d launches, ~%.1f steps, %.1f us of kernels per step" % (path, len(rows), nsteps, tot / nsteps))
    for k, (c, t, g, b) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:18]:
        print("  %-44s grid=%-16s n/step=%5.1f avg=%8.1fus per-step=%7.1fus share=%5.1f%%" % (k, g, c / nsteps, t / c, t / nsteps, 100 * t / tot))


if __name__ == "__main__":
    main(*sys.argv[1:2])
