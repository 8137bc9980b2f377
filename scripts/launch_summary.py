"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total ms and share."""
import collections, csv, re, sys

def load(fn):
    rows = list(csv.reader(l for l in open(fn) if l.startswith('"')))
    h = rows[0]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        v *= {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}[r[ui]]
        name = re.sub(r"\(.*", "", r[ki])
        name = re.sub(r"^void ", "", name)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1; a[1] += v
    return agg

if __name__ == "__main__":
    agg = load(sys.argv[1])
    tot = sum(v[1] for v in agg.values())
    print("total %.2f ms in %d launches" % (tot, sum(v[0] for v in agg.values())))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
        print("%-78s %5d %9.3f ms %5.1f%%" % (k[:78], v[0], v[1], 100 * v[1] / tot))
