"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total
and share of the device time (cold-cache, serialised: compare SHARES, not absolutes)."""
import collections, csv, re, sys
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
h = rows[0]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
tot = collections.OrderedDict()
for r in rows[1 + skip:]:
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("mphx::", "")
    v = float(r[vi].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6}.get(r[ui], 1e-6)
    c = tot.setdefault(name, [0, 0.0])
    c[0] += 1
    c[1] += v
all_ms = sum(v for _, v in tot.values())
print(f"# {sys.argv[1]}: {len(rows) - 1 - skip} launches, {all_ms:.3f} ms device time (launches after the first {skip})")
print(f"{'kernel':44s} {'launches':>8s} {'total ms':>10s} {'avg ms':>9s} {'share':>7s}")
for k, (n, v) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:44s} {n:8d} {v:10.3f} {v / n:9.4f} {100 * v / all_ms:6.1f}%")
