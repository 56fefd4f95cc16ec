"""ncu source page CSV -> the instructions with the most warp-stall samples of a kernel and their dominant stall reasons.
usage: python tools/ncu_top_stalls.py source.csv <kernel name substring> [N]"""
import csv, sys
path = sys.argv[1]
rows = list(csv.reader(open(path)))
# split into kernels
kernels = []
for r in rows:
    if r and r[0] == "Kernel Name":
        kernels.append({"name": r[1], "hdr": None, "data": []})
    elif r and r[0] == "Address":
        kernels[-1]["hdr"] = r
    elif kernels and kernels[-1]["hdr"] is not None and len(r) > 10:
        kernels[-1]["data"].append(r)
want = sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 30
for k in kernels:
    if want not in k["name"]:
        continue
    hdr, data = k["hdr"], k["data"]
    idx = {h: i for i, h in enumerate(hdr)}
    S = lambda r, h: int(r[idx[h]] or 0)
    tot = sum(S(r, "# Samples") for r in data)
    print("==", k["name"][:60], "total samples", tot)
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = {h: sum(S(r, h) for r in data) for h in stall_cols}
    print("by reason:", [(k2[6:], round(100 * v / tot, 1)) for v, k2 in sorted(((v, k2) for k2, v in agg.items() if v), reverse=True)[:9]])
    top = sorted(range(len(data)), key=lambda i: -S(data[i], "# Samples"))[:N]
    for i in sorted(top):
        r = data[i]
        n = S(r, "# Samples")
        reasons = sorted(((S(r, h), h[6:]) for h in stall_cols), reverse=True)[:2]
        print(f"{i:5d} {100*n/tot:5.1f}%  {r[idx['Source']].strip()[:64]:64s} {reasons}")
