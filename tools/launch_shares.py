"""Per-kernel totals and shares of an `ncu --metrics gpu__time_duration.sum --csv` launch list (profiles/*_launches.csv)."""
import collections
import csv
import sys


def main(path, header):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hi]
    ni, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg, tot = collections.OrderedDict(), 0.0
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1.0)
        a = agg.setdefault(r[ni][:64], [0.0, 0])
        a[0] += v
        a[1] += 1
        tot += v
    print(header)
    print("        us     n  share  kernel")
    for k, (v, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print("%10.1f %5d %5.1f%%  %s" % (v, n, 100 * v / tot, k))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
