"""ncu raw-page CSV (ncu -i X.ncu-rep --page raw --csv) -> markdown table of the metrics quoted in DESIGN.md / profiles."""
import csv, sys
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_l1tex2xbar_write_bytes.sum",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.max"]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
names = [r[idx["Kernel Name"]].split("(")[0].replace("void ", "")[:34] for r in data]
print("| metric | unit | " + " | ".join(f"{i}: {n}" for i, n in enumerate(names)) + " |")
print("|---|---|" + "---|" * len(data))
for k in KEYS:
    if k in idx:
        vals = []
        for r in data:
            v = r[idx[k]]
            try:
                v = f"{float(v.replace(',', '')):.4g}"
            except ValueError:
                pass
            vals.append(v)
        print(f"| {k} | {units[idx[k]]} | " + " | ".join(vals) + " |")
