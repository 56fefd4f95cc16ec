"""Per-phase profile of the training step: CUDA-event times of the graph-replayed D step / G step / penalty D step,
then (eager) a torch.profiler kernel table for the D step and the G step separately, and the convolution launches of a
plain step grouped by shape.  For optimisation guidance; nothing here is a bench value.
usage: python tools/profile_phases.py [size=256] [batch=16]"""
import collections
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gigagan_pytorch_b200 as g
from gigagan_pytorch_b200 import ops
from bench import G_CFG, D_CFG, real_batch
from gigagan_pytorch_b200.trainer import cycle

size = int(sys.argv[1]) if len(sys.argv) > 1 else 256
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
dev = torch.device("cuda:0")
torch.manual_seed(0)
gan = g.GigaGAN(generator=dict(G_CFG, image_size=size), discriminator=dict(D_CFG, image_size=size), amp=True,
                mixed_precision_type="bf16", log_steps_every=10 ** 9, save_and_sample_every=0).to(dev)
if os.environ.get("GG_NO_MERGE"):
    gan.merge_real_fake = False


class Pool:
    batch_size = B

    def __iter__(self):
        return iter([real_batch(s, 1, 0, B, size).to(dev) for s in range(2)])


it = cycle(Pool())
gan.use_cuda_graphs = True
for rep in range(3):                      # eager warm-up, capture, replay of every variant
    for gp in (False, True):
        gan.train_discriminator_step(dl_iter=it, apply_gradient_penalty=gp)
        gan.train_generator_step(batch_size=B, dl_iter=it)
torch.cuda.synchronize()
ev = lambda: torch.cuda.Event(enable_timing=True)
acc = collections.defaultdict(list)
for rep in range(6):
    for name, gp in (("plain", False), ("gp", True)):
        e0, e1, e2 = ev(), ev(), ev()
        e0.record()
        gan.train_discriminator_step(dl_iter=it, apply_gradient_penalty=gp)
        e1.record()
        gan.train_generator_step(batch_size=B, dl_iter=it)
        e2.record()
        torch.cuda.synchronize()
        acc["D-" + name].append(e0.elapsed_time(e1))
        acc["G"].append(e1.elapsed_time(e2))
for k, v in acc.items():
    v = sorted(v)
    print(f"[graph replay] {k:8s} median {v[len(v) // 2]:.2f} ms  (min {v[0]:.2f})", flush=True)
med = lambda k: sorted(acc[k])[len(acc[k]) // 2]
cyc = (3 * med("D-plain") + med("D-gp")) / 4 + med("G")
print(f"[graph replay] cycle mean {cyc:.2f} ms/step -> {B / cyc * 1e3:.1f} img/s", flush=True)

if os.environ.get("GG_TIMING_ONLY"):
    sys.exit(0)
gan.use_cuda_graphs = False
from torch.profiler import profile, ProfilerActivity


def table(fn, title):
    fn()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        fn()
        torch.cuda.synchronize()
    from torch.autograd import DeviceType
    evs = [k for k in prof.key_averages() if k.self_device_time_total > 0]
    rows = [(k.self_device_time_total / 1e3, k.count, k.key[:100]) for k in evs if k.device_type == DeviceType.CUDA]
    ops_ = [(k.self_device_time_total / 1e3, k.count, k.key[:100]) for k in evs if k.device_type != DeviceType.CUDA]
    tot, n = sum(r[0] for r in rows), sum(r[1] for r in rows)
    print(f"===== {title}: {tot:.2f} ms kernel time in {n} launches (ms, count, kernel)")
    for r in sorted(rows, reverse=True)[:70]:
        print(f"{r[0]:9.3f} {r[1]:6d}  {r[2]}")
    small = [r for r in rows if r[0] / r[1] < 0.02]
    print(f"      kernels averaging < 20 us: {sum(r[1] for r in small)} launches, {sum(r[0] for r in small):.2f} ms")
    at = [r for r in rows if "at::" in r[2] or r[2].startswith("Mem")]
    print(f"      ATen / memset / memcpy nodes: {sum(r[1] for r in at)} launches, {sum(r[0] for r in at):.2f} ms")
    print(f"  --- by launching operator (device ms attributed, calls)")
    for r in sorted(ops_, key=lambda r: -r[1])[:40]:
        print(f"{r[0]:9.3f} {r[1]:6d}  {r[2]}")


table(lambda: gan.train_discriminator_step(dl_iter=it, apply_gradient_penalty=False), "D step (plain)")
table(lambda: gan.train_generator_step(batch_size=B, dl_iter=it), "G step")
if not os.environ.get("GG_SKIP_GP"):
    table(lambda: gan.train_discriminator_step(dl_iter=it, apply_gradient_penalty=True), "D step (gradient penalty)")

prof = ops.ConvProfiler()
with prof:
    gan.train_discriminator_step(dl_iter=it, apply_gradient_penalty=False)
    gan.train_generator_step(batch_size=B, dl_iter=it)
torch.cuda.synchronize()
by = collections.defaultdict(lambda: [0, 0.0, 0.0])
for rec in prof.records:
    kind, flops, e0, e1, dt = rec[:5]
    key = rec[5] if len(rec) > 5 else kind
    b = by[key]
    b[0] += 1
    b[1] += e0.elapsed_time(e1)
    b[2] += flops
print("===== convolution fprop/dgrad launches of a plain step by shape (count, ms, TFLOP/s, shape)")
for k, (n, ms, fl) in sorted(by.items(), key=lambda kv: -kv[1][1]):
    print(f"{n:4d} {ms:8.3f} {fl / ms / 1e9 if ms else 0:8.0f}  {k}")
print(f"total {sum(v[1] for v in by.values()):.2f} ms, {sum(v[2] for v in by.values()) / 1e12:.2f} TFLOP")
