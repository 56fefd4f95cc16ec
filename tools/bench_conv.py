"""Micro-benchmark of the convolution kernels on the discriminator's shapes (B=16): CUDA-event timing, L2 flushed
between iterations.  usage: python tools/bench_conv.py [ffma]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gigagan_pytorch_b200 import ops, _lib

dev = torch.device("cuda:0")
shapes = [  # (name, N, H, Cin, Cout, k)
    ("d_res16_512_3x3_b256", 256, 16, 512, 512, 3), ("d_res32_256_3x3_b64", 64, 32, 256, 256, 3),
    ("d_res8_512_3x3_b256", 256, 8, 512, 512, 3), ("d_res64_128_3x3_b32", 32, 64, 128, 128, 3),
    ("d_res128_64_3x3_b16", 16, 128, 64, 64, 3), ("d_res256_32_3x3_b16", 16, 256, 32, 32, 3),
    ("d_res4_512_3x3_b256", 256, 4, 512, 512, 3), ("ff_res32_256_1024_1x1_b64", 64, 32, 256, 1024, 1),
]
flush = torch.empty(256 * 1024 * 1024, dtype=torch.int8, device=dev)
L = _lib.lib()
if len(sys.argv) > 1 and sys.argv[1] == "ffma":
    L.gg_set_flags(1)
out = []
for name, n, h, ci, co, k in shapes:
    x = torch.randn(n, h, h, ci, device=dev).to(torch.bfloat16)
    w = (torch.randn(co, k, k, ci, device=dev) * (ci * k * k) ** -0.5).to(torch.bfloat16)
    b = torch.randn(co, device=dev)
    g = ops.ConvGeom(k, k, 1, k // 2, False, act=1)
    for _ in range(3):
        y = ops._conv_fprop_raw(x, w, b, None, g, co)
    ts = []
    for _ in range(8):
        flush.zero_()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        y = ops._conv_fprop_raw(x, w, b, None, g, co)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    flops = 2.0 * n * h * h * co * ci * k * k
    byts = 2.0 * (x.numel() + y.numel() + w.numel())
    out.append(dict(name=name, ms=round(ms, 4), tflops=round(flops / ms / 1e9, 1), gbs=round(byts / ms / 1e6, 1)))
    print(json.dumps(out[-1]), flush=True)
