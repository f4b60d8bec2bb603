"""ncu launch list (--metrics gpu__time_duration.sum --csv) -> per-kernel shares.  Usage: launch_list.py X.csv > profiles/rN_launch_list.txt"""
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
iN, iV = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = {}
for r in rows[1:]:
    n = r[iN]
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1; a[1] += float(r[iV].replace(",", ""))
tot = sum(v[1] for v in agg.values())
print("# ncu --metrics gpu__time_duration.sum --clock-control none -c %d, python bench.py --steps 10 --warmup 3 (first %d launches: the timed loop,"
      % (len(rows) - 1, len(rows) - 1))
print("# the K = 16M / 64M strong-scaling ticks (the ~5 ms means), then the first extras)")
print("# per-launch times are cold-cache and serialised: compare SHARES")
for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{n[:90]:90s} launches={c:3d} total_ns={t:10.0f} mean_ns={t / c:9.0f} share={100 * t / tot:5.1f}%")
