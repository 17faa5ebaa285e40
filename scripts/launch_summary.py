"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0]); tot = 0.0
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
    v = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
    name = re.sub(r"\(.*", "", row["Kernel Name"])[:72]
    agg[name][0] += 1; agg[name][1] += v; tot += v
print(f"total {tot:.0f} us over {sum(n for n, _ in agg.values())} launches")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 16]:
    print(f"{t / tot * 100:6.2f}% {t:11.1f} us {n:5d}  {k}")
