"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals over the last
`frac` of the captured launches (the steady-state steps). Usage: summarize_launches.py file.csv [frac]"""
import collections
import csv
import sys

path = sys.argv[1]
frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.4
rows = list(csv.reader(open(path)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
gi, bi = hdr.index("Grid Size"), hdr.index("Block Size")
sel = data[int(len(data) * (1 - frac)):]
agg = collections.OrderedDict()
for r in sel:
    name = r[ki]
    name = name[:name.index("(")] if "(" in name else name
    v = float(r[vi].replace(",", ""))
    v = v / 1000 if r[ui] == "ns" else (v * 1000 if r[ui] == "ms" else v)
    a = agg.setdefault(name[:90], [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print("launches %d (last %.0f%% of %d captured)  total %.1f us" % (len(sel), frac * 100, len(data), tot))
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
    print("%9.1f us %5.1f%%  n=%4d  avg %8.2f us  %s" % (t, 100 * t / tot, c, t / c, k))
