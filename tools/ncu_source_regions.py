"""ncu source page (ncu -i X.ncu-rep --page source --csv) -> cumulative share of the warp-stall samples up to every
barrier / TMEM / TMA / memory instruction of a kernel: where the time of a warp-specialised kernel sits.
usage: python tools/ncu_source_regions.py source.csv <kernel name substring>"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
kernels = []
for r in rows:
    if r and r[0] == "Kernel Name":
        kernels.append({"name": r[1], "hdr": None, "data": []})
    elif r and r[0] == "Address":
        kernels[-1]["hdr"] = r
    elif kernels and kernels[-1]["hdr"] is not None and len(r) > 10:
        kernels[-1]["data"].append(r)
seen = set()
for k in kernels:
    if sys.argv[2] not in k["name"] or k["name"] in seen:
        continue
    seen.add(k["name"])
    hdr, data = k["hdr"], k["data"]
    idx = {h: i for i, h in enumerate(hdr)}
    S = lambda r, h: int(r[idx[h]] or 0)
    tot = sum(S(r, "# Samples") for r in data)
    marks = ("LDTM", "BAR.SYNC", "UTMALDG", "UTCHMMA", "EXIT", "MEMBAR", "UTCBAR", "STG", "SYNCS.PHASECHK", "SYNCS.ARRIVE", "LDGSTS", "LDGDEPBAR", "DEPBAR")
    acc = 0
    last = 0
    print("==", k["name"][:50], tot)
    for i, r in enumerate(data):
        src = r[idx["Source"]].strip()
        acc += S(r, "# Samples")
        if any(m in src for m in marks):
            print(f"{i:5d} cum {100*acc/tot:5.1f}% (+{100*(acc-last)/tot:4.1f})  {src[:70]}  exec={r[idx['Instructions Executed']]}")
            last = acc
