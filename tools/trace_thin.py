"""Timeline of CTA 0 of the thin-layer convolution kernel (debug): per pipeline role, when each ring row was waited for,
issued, landed and consumed.  Prints per-row timestamps (us, relative to the first event) and per-stage averages."""
import os, sys, ctypes, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gigagan_pytorch_b200 import ops, _lib
dev = torch.device("cuda:0")
n, h, ci, co = 16, 256, int(sys.argv[1]) if len(sys.argv) > 1 else 16, int(sys.argv[2]) if len(sys.argv) > 2 else 32
x = torch.randn(n, h, h, ci, device=dev).to(torch.bfloat16)
w = (torch.randn(co, 3, 3, ci, device=dev) * 0.1).to(torch.bfloat16)
b = torch.randn(co, device=dev)
g = ops.ConvGeom(3, 3, 1, 1, False, act=1)
for _ in range(3):
    y = ops._conv_fprop_raw(x, w, b, None, g, co)
buf = torch.zeros(32 * 64 * 8, dtype=torch.int64, device=dev)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.int8, device=dev)
flush.zero_()
L = _lib.lib()
L.gg_debug_thin_trace(ctypes.c_void_p(buf.data_ptr()))
y = ops._conv_fprop_raw(x, w, b, None, g, co)
torch.cuda.synchronize()
L.gg_debug_thin_trace(ctypes.c_void_p(0))
t = buf.cpu().tolist()
ev = []
for role in range(32):
    for row in range(64):
        for stage in range(8):
            v = t[(role * 64 + row) * 8 + stage]
            if v:
                ev.append((role, row, stage, v))
t0 = min(e[3] for e in ev)
GHZ = 1.9
by = collections.defaultdict(dict)
for role, row, stage, clk in ev:
    by[(role, row)][stage] = (clk - t0) / GHZ / 1e3
print(f"{len(ev)} events; roles: 1 = MMA (0 start,1 rows full,2 acc free,3 issued), 2-5 = epilogue (0 wait,1 acc full,2 stored), "
      f"10+ = loader warp (0 wait slot,1 slot free,2 issued,3 landed,4 fenced)")
rows_m = sorted(r for (role, r) in by if role == 1)
base_row = 0
print("row   loader: free  issued landed fenced |  MMA: start full  accfree issued | epi(w2): wait  full  stored")
lo = {}
for (role, c), st in by.items():
    if role >= 10:
        lo[c] = st
for r in rows_m:
    m = by[(1, r)]
    e = by.get((2, r), {})
    l = lo.get(r - base_row + 2, {})          # ring entry c = output row + 2 is the last row it needs
    f = lambda d, k: f"{d[k]:7.2f}" if k in d else "      -"
    print(f"{r - base_row:3d}   {f(l,1)} {f(l,2)} {f(l,3)} {f(l,4)} |   {f(m,0)} {f(m,1)} {f(m,2)} {f(m,3)} |   {f(e,0)} {f(e,1)} {f(e,2)}")
def avg(role_pred, a, b_):
    d = [st[b_] - st[a] for (role, r), st in by.items() if role_pred(role) and a in st and b_ in st]
    return sum(d) / max(len(d), 1), len(d)
print("loader: wait-for-slot %.2f us, issue %.2f, memory %.2f, fence %.2f (n=%d)" % (avg(lambda r: r >= 10, 0, 1)[0], avg(lambda r: r >= 10, 1, 2)[0],
      avg(lambda r: r >= 10, 2, 3)[0], avg(lambda r: r >= 10, 3, 4)[0], avg(lambda r: r >= 10, 3, 4)[1]))
print("MMA: wait rows %.2f us, wait acc %.2f, issue %.2f" % (avg(lambda r: r == 1, 0, 1)[0], avg(lambda r: r == 1, 1, 2)[0], avg(lambda r: r == 1, 2, 3)[0]))
print("epilogue: wait acc %.2f us, drain+store %.2f" % (avg(lambda r: 2 <= r <= 5, 0, 1)[0], avg(lambda r: 2 <= r <= 5, 1, 2)[0]))
print("last event at %.2f us" % max((e[3] - t0) / GHZ / 1e3 for e in ev))
