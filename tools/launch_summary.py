"""Per-kernel totals of the LAST pass in an ncu launch list (--metrics gpu__time_duration.sum --csv) of tools/ncu_pass.py."""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
start = next(i for i, r in enumerate(rows) if r[0] == "ID")
hdr = rows[start]
kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
data = rows[start + 1:]
begins = [i for i, r in enumerate(data) if r[kn].startswith("k_pass_begin")]
last = data[begins[-1]:]
agg = collections.OrderedDict()
tot = 0.0
for r in last:
    v = float(r[mv].replace(",", "")) / 1000.0
    tot += v
    n = r[kn].split("(")[0]
    a = agg.setdefault(n, [0.0, 0])
    a[0] += v; a[1] += 1
print(f"{len(last)} launches in the last pass, serialised kernel time {tot:.1f} us")
for k, (v, c) in sorted(agg.items(), key=lambda x: -x[1][0]):
    print(f"{v:9.1f} us  {c:3d} x  {k}   ({100 * v / tot:.1f} %)")
