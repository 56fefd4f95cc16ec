"""Summarise an ncu --csv launch list (gpu__time_duration.sum) into per-kernel totals / shares."""
import csv, sys, collections, re
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
r = csv.DictReader(lines)
tot = collections.defaultdict(lambda: [0, 0.0])
for row in r:
    if "gpu__time_duration" not in row.get("Metric Name", ""):
        continue
    name = re.sub(r"\(.*", "", row["Kernel Name"])[:70]
    v = float(row["Metric Value"].replace(",", ""))
    unit = row.get("Metric Unit", "ns")
    ms = v / 1e6 if unit in ("ns", "nsecond") else v / 1e3 if unit in ("us", "usecond") else v
    tot[name][0] += 1
    tot[name][1] += ms
allms = sum(v[1] for v in tot.values())
print(f"total kernel time {allms:.2f} ms over {sum(v[0] for v in tot.values())} launches")
print(f"{'kernel':70s} {'launches':>8s} {'ms':>9s} {'share':>7s}")
for name, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{name:70s} {n:8d} {ms:9.3f} {100*ms/allms:6.1f}%")
