"""Summarise an ncu --csv launch list (gpu__time_duration.sum [+ dram__bytes_read.sum, dram__bytes_write.sum]) into
per-kernel totals / shares / DRAM traffic per launch."""
import csv, sys, collections, re
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
tot = collections.defaultdict(lambda: dict(n=0, ms=0.0, rd=0.0, wr=0.0))
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
for row in csv.DictReader(lines):
    name = re.sub(r"\(.*", "", row["Kernel Name"])[:70]
    metric, unit = row.get("Metric Name", ""), row.get("Metric Unit", "")
    v = float(row["Metric Value"].replace(",", ""))
    t = tot[name]
    if "gpu__time_duration" in metric:
        t["n"] += 1
        t["ms"] += v / 1e6 if unit in ("ns", "nsecond") else v / 1e3 if unit in ("us", "usecond") else v
    elif "dram__bytes_read" in metric:
        t["rd"] += v * SCALE.get(unit, 1.0)
    elif "dram__bytes_write" in metric:
        t["wr"] += v * SCALE.get(unit, 1.0)
allms = sum(v["ms"] for v in tot.values())
print(f"total kernel time {allms:.2f} ms over {sum(v['n'] for v in tot.values())} launches "
      f"(ncu per-launch times are serialised and cold-cache: read the SHARES)")
print(f"{'kernel':70s} {'launches':>8s} {'ms':>9s} {'share':>7s} {'dram MB/launch':>15s} {'GB/s':>8s}")
for name, t in sorted(tot.items(), key=lambda kv: -kv[1]["ms"])[:45]:
    mb = (t["rd"] + t["wr"]) / max(t["n"], 1) / 1e6
    gbs = (t["rd"] + t["wr"]) / max(t["ms"], 1e-9) / 1e6
    print(f"{name:70s} {t['n']:8d} {t['ms']:9.3f} {100 * t['ms'] / allms:6.1f}% {mb:15.2f} {gbs:8.0f}")

# optional second argument: write the per-step / per-class DRAM traffic that bench.py reports as roofline.traffic
if len(sys.argv) > 2:
    import json
    conv = [t for name, t in tot.items() if name.startswith("conv_fprop_tc") or name.startswith("void conv_thin_tc")
            or name.startswith("conv_thin_tc")]
    cn = sum(t["n"] for t in conv)
    out = {"source": "ncu launch list of one plain G+D step (tools/run_ncu.sh, --clock-control none): " + sys.argv[1].split("/")[-1],
           "launches_per_step": sum(v["n"] for v in tot.values()),
           "dram_bytes_per_step": sum(v["rd"] + v["wr"] for v in tot.values()),
           "conv_launches": cn,
           "conv_dram_bytes_per_launch": (sum(t["rd"] + t["wr"] for t in conv) / cn) if cn else None,
           "kernel_ms_ncu": allms}
    json.dump(out, open(sys.argv[2], "w"), indent=1)
