"""SASS instruction census of the tensor-core objects (cuobjdump -sass of the in-tree .o files): per kernel, the counts of the
mnemonics that prove the Blackwell-native path (UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, UTMALDG = TMA load, LDTM =
tcgen05.ld, LDGSTS = cp.async, SYNCS = mbarrier).  usage: python tools/sass_census.py > profiles/rNN_sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CS = os.path.join(ROOT, "gigagan_pytorch_b200", "csrc")
KEYS = ["UTCHMMA", "UTCBAR", "UTMALDG", "UTMASTG", "LDTM", "LDGSTS", "SYNCS", "ELECT", "R2UR", "HMMA", "STG", "LDG", "STS", "LDS",
        "FFMA", "MUFU", "BAR"]
print("SASS instruction census of the tensor-core objects (cuobjdump -sass gigagan_pytorch_b200/csrc/<obj>.o, sm_100a, this build).\n"
      "UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, UTMALDG = cp.async.bulk.tensor (TMA load), LDTM = tcgen05.ld, LDGSTS = cp.async,\n"
      "SYNCS = mbarrier ops.  No UTMASTG (epilogues store with STG.E.256), no HMMA (mma.sync) in any tensor-core kernel.\n")
for obj in ("conv_tc.o", "conv_thin_tc.o", "attn_tc2.o", "attn_tc.o", "bmm_tc.o"):
    out = subprocess.run(["cuobjdump", "-sass", os.path.join(CS, obj)], capture_output=True, text=True).stdout
    fn, cnt, n = None, None, 0
    def flush():
        if fn:
            print(f"{obj}  {fn}\n    {n} instructions; " + ", ".join(f"{k} {cnt[k]}" for k in KEYS if cnt[k]))
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            flush()
            fn, cnt, n = m.group(1), collections.Counter(), 0
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)", line)
        if m and fn:
            n += 1
            op = m.group(1).split(".")[0]
            if op in cnt or op in KEYS:
                cnt[op] += 1
    flush()
