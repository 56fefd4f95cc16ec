"""Per-layer timing of the convolution kernels over the discriminator/generator shapes of the 256x256 config (B=16).
CUDA events, L2 flushed between iterations.  Prints one JSON line per (layer, op)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gigagan_pytorch_b200 import ops

dev = torch.device("cuda:0")
B = 16
# (name, N, H, Cin, Cout, k, stride)
L = []
dims = [(256, 16, 32, 1), (128, 32, 64, 1), (64, 64, 128, 2), (32, 128, 256, 4), (16, 256, 512, 8), (8, 512, 512, 16), (4, 512, 512, 16)]
for r, ci, co, s in dims:
    n = B * s
    L.append((f"D{r}_block1", n, r, ci, co, 3, 1))
    L.append((f"D{r}_block2", n, r, co, co, 3, 1))
    if r > 4:
        L.append((f"D{r}_down2x2s2", n, r, co, co, 2, 2))
        L.append((f"D{r}_res1x1s2", n, r, ci, co, 1, 2))
for r, c, s in [(32, 256, 2), (16, 512, 4), (8, 512, 8), (4, 512, 16)]:
    L.append((f"P{r}_3x3", B * s, r, c, c, 3, 1))
L.append(("A32_q1x1", B * 4, 32, 256, 512, 1, 1))
L.append(("A32_ff1", B * 4, 32, 256, 1024, 1, 1))
L.append(("A32_ff2", B * 4, 32, 1024, 256, 1, 1))
L.append(("A16_ff1", B * 8, 16, 512, 2048, 1, 1))
L.append(("Dfromrgb64_7x7", B, 64, 16, 64, 7, 1))
L.append(("D256_first", B, 256, 16, 32, 3, 1))
for r, ci, co in [(16, 512, 256), (32, 256, 128), (64, 128, 64), (128, 64, 32), (256, 32, 16)]:
    L.append((f"G{r}_conv1_persample", B, r, ci, co, 3, 1))
flush = torch.empty(256 * 1024 * 1024, dtype=torch.int8, device=dev)


def timeit(fn, iters=5):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


only = sys.argv[1] if len(sys.argv) > 1 else None
tot = dict(fprop=0.0, dgrad=0.0, wgrad=0.0)
for name, n, h, ci, co, k, s in L:
    if only and only not in name:
        continue
    ps = "persample" in name
    x = torch.randn(n, h, h, ci, device=dev).to(torch.bfloat16)
    wshape = ((n,) if ps else ()) + (co, k, k, ci)
    w = (torch.randn(*wshape, device=dev) * (ci * k * k) ** -0.5).to(torch.bfloat16)
    b = torch.randn(co, device=dev)
    pad = k // 2 if s == 1 else 0
    g = ops.ConvGeom(k, k, s, pad, ps, act=0)
    y = ops._conv_fprop_raw(x, w, b, None, g, co)
    gy = torch.randn_like(y)
    flops = 2.0 * n * y.shape[1] * y.shape[2] * co * ci * k * k
    byts = 2.0 * (x.numel() + y.numel() + w.numel())
    res = dict(name=name, gflop=round(flops / 1e9, 1), mbytes=round(byts / 1e6, 1))
    for op, fn in (("fprop", lambda: ops._conv_fprop_raw(x, w, b, None, g, co)),
                   ("dgrad", lambda: ops._conv_dgrad_raw(gy, w, g, tuple(x.shape))),
                   ("wgrad", lambda: ops._conv_wgrad_raw(x, gy, g))):
        ms = timeit(fn)
        tot[op] += ms
        res[op] = dict(ms=round(ms, 4), tflops=round(flops / ms / 1e9, 1), gbs=round(byts / ms / 1e6, 1))
    print(json.dumps(res), flush=True)
print(json.dumps(dict(total_ms=tot)))
