"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches and median duration per kernel."""
import csv, collections, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
d = collections.OrderedDict()
for r in rows[1:]:
    d.setdefault(r[ki][:78], []).append(float(r[vi].replace(",", "")))
tot = sum(sum(v) for v in d.values())
for k, v in d.items():
    print("%-80s n=%4d  median %8.1f us  share %5.1f%%" % (k, len(v), sorted(v)[len(v) // 2] / 1e3, 100 * sum(v) / tot))
