"""Batched-GEMM timings at the shapes of the gradient-penalty attention node (discriminator res-32 layer: 128 images x 8
heads, 1024 queries, 1088 padded keys): every product reads or writes one (tokens x keys) bf16 tensor (2.28 GB), so the
figure of merit is GB/s against the HBM peak.  CUDA events, operands larger than L2."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gigagan_pytorch_b200 import ops
dev = torch.device("cuda:0")
b, h, n, m, d, D = (int(sys.argv[1]) if len(sys.argv) > 1 else 128), 8, 1024, 1088, 64, 80
bf = torch.bfloat16


def phys(rows, cols):          # (b, h, rows, cols) view of (b, rows, h, cols) storage, like the attention operands
    return (torch.randn(b, rows, h, cols, device=dev) * 0.1).to(bf).permute(0, 2, 1, 3)


P = (torch.rand(b, h, n, m, device=dev) * 0.01).to(bf)
qa, ka, v, go = phys(n, D), phys(m, D), phys(m, d), phys(n, d)
qq, kk2 = phys(n, 2 * D), phys(m, 2 * D)
g2, v2 = phys(n, 2 * d), phys(m, 2 * d)
T = lambda x: x.transpose(-1, -2)
cases = [
    ("S = qa ka^T (K=80, writes P-size)", lambda: ops._bmm_raw(qa, T(ka), None, 0.25, False), 1),
    ("dP = go v^T (K=64, writes)", lambda: ops._bmm_raw(go, T(v), None, 1.0, False), 1),
    ("G = [uq|qa][ka|uk]^T (K=160, writes)", lambda: ops._bmm_raw(qq, T(kk2), None, 0.25, False), 1),
    ("dP_tot = [go|go1][v|uv]^T (K=128, writes)", lambda: ops._bmm_raw(g2, T(v2), None, 1.0, False), 1),
    ("o = P v (reads, N=64)", lambda: ops._bmm_raw(P, v, None, 1.0, True), 1),
    ("dv = P^T go (reads, MN-major A)", lambda: ops._bmm_raw(T(P), go, None, 1.0, v), 1),
    ("dqa = dS ka (reads, N=80)", lambda: ops._bmm_raw(P, ka, None, 0.25, qa), 1),
    ("dka = dS^T qa (reads, MN-major A, N=80)", lambda: ops._bmm_raw(T(P), qa, None, 0.25, ka), 1),
]
psize = P.numel() * 2
for name, fn, units in cases:
    for _ in range(2):
        fn()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[2]
    print(json.dumps(dict(case=name, ms=round(ms, 4), gbs=round(units * psize / ms / 1e6, 1))), flush=True)
