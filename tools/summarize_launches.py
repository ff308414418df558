"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per
kernel name, launches, total and mean duration, share of the total."""
import csv
import re
import sys
from collections import OrderedDict


def main(path, skip=0):
    rows = []
    with open(path) as fh:
        lines = [l for l in fh if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"]
        name = re.sub(r"\(.*", "", name)
        name = name.replace("void ", "").replace("rs::<unnamed>::", "rs::")
        val = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ns = val * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        rows.append((int(r["ID"]), name, ns))
    rows = [r for r in rows if r[0] >= skip]
    agg = OrderedDict()
    for _, name, ns in rows:
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ns
    total = sum(a[1] for a in agg.values())
    print("%-70s %8s %12s %12s %7s" % ("kernel", "launches", "total_ms", "mean_us", "share"))
    for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-70s %8d %12.3f %12.1f %6.1f%%" % (name[:70], n, ns / 1e6, ns / n / 1e3, 100 * ns / total))
    print("%-70s %8d %12.3f" % ("TOTAL", len(rows), total / 1e6))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
