"""Fused attention (tcgen05) timing at the discriminator's / generator's shapes: forward and backward, CUDA events, for the
kernel variants (gg_set_flags 8: first generation; 16 / 32: second generation with 8 / 16 softmax warps everywhere;
64: two-pass L2 forward; 128: forward with one CTA per SM;
0: the default - second generation, forward two CTAs per SM x 8 warps, backward 16 warps).  Also prints the largest deviation of every output/gradient from the first generation."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gigagan_pytorch_b200 import ops, _lib
dev = torch.device("cuda:0")
L = _lib.lib()
flush = torch.empty(256 * 1024 * 1024, dtype=torch.int8, device=dev)
VARIANTS = (("gen1_8w", 8), ("gen2_8w", 16), ("gen2_16w", 32), ("gen2_2pass", 64), ("gen2_fwd_1cta", 128), ("gen2_default", 0))
ONLY = sys.argv[1] if len(sys.argv) > 1 else None        # optional: one shape name, one flag value (ncu captures)
if len(sys.argv) > 2:
    VARIANTS = tuple(v for v in VARIANTS if v[1] == int(sys.argv[2]))
for name, B, n, l2 in (("D_res32_l2", 64, 1024, True), ("D_res16_l2", 128, 256, True), ("G_res32_dot", 16, 1024, False),
                       ("G_res16_dot", 16, 256, False)):
    if ONLY and name != ONLY:
        continue
    heads, d = 8, 64
    torch.manual_seed(0)
    qkv = (torch.randn(B, n, 3 * heads * d, device=dev) * 0.5).to(torch.bfloat16).requires_grad_()
    nk = torch.randn(2, heads, d, device=dev).requires_grad_()
    q, k, v = qkv[..., :512], qkv[..., 512:1024], qkv[..., 1024:]
    go = torch.randn(B, n, heads * d, device=dev).to(torch.bfloat16)

    def fwd():
        return ops.fused_attention(q, q if l2 else k, v, nk, heads, d ** -0.5, l2=l2)

    ref = None
    for vname, flag in VARIANTS:
        old = L.gg_set_flags(flag)
        try:
            for _ in range(2):
                o = fwd(); g = torch.autograd.grad(o, (qkv, nk), go)
            outs = [o.detach().float(), g[0].float(), g[1].float()]
            tf, tb = [], []
            for _ in range(5):
                flush.zero_()
                e0, e1, e2 = (torch.cuda.Event(True) for _ in range(3))
                e0.record(); o = fwd(); e1.record(); torch.autograd.grad(o, (qkv, nk), go); e2.record()
                torch.cuda.synchronize()
                tf.append(e0.elapsed_time(e1)); tb.append(e1.elapsed_time(e2))
        finally:
            L.gg_set_flags(old)
        tf, tb = sorted(tf)[2], sorted(tb)[2]
        flops = 4.0 * B * heads * n * (n + 1) * d
        dev_rel = None
        if ref is None:
            ref = outs
        else:
            dev_rel = [round(((a - b).abs().max() / b.abs().max()).item(), 5) for a, b in zip(outs, ref)]
        print(json.dumps(dict(name=name, variant=vname, fwd_ms=round(tf, 4), bwd_ms=round(tb, 4),
                              fwd_tflops=round(flops / tf / 1e9, 1), bwd_tflops=round(2.5 * flops / tb / 1e9, 1),
                              exps_per_s_fwd=round(B * heads * n * n / tf / 1e9 * 1e3, 1),
                              max_rel_dev_vs_gen1=dev_rel, finite=bool(all(torch.isfinite(t).all() for t in outs)))),
              flush=True)
