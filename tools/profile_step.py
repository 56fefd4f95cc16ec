"""Kernel-time breakdown of one training step (torch.profiler / CUPTI), for optimisation guidance only."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gigagan_pytorch_b200 as g
from bench import G_CFG, D_CFG, real_batch
from gigagan_pytorch_b200.trainer import cycle

size = int(sys.argv[1]) if len(sys.argv) > 1 else 256
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
dev = torch.device("cuda:0")
torch.manual_seed(0)
gan = g.GigaGAN(generator=dict(G_CFG, image_size=size), discriminator=dict(D_CFG, image_size=size), amp=True,
                mixed_precision_type="bf16", log_steps_every=10 ** 9).to(dev)


class Pool:
    batch_size = B

    def __iter__(self):
        return iter([real_batch(s, 1, 0, B, size).to(dev) for s in range(2)])


it = cycle(Pool())
for gp in (False, True):
    gan.train_discriminator_step(dl_iter=it, apply_gradient_penalty=gp)
    gan.train_generator_step(batch_size=B, dl_iter=it)
torch.cuda.synchronize()
import time
for name, gp in (("plain", False), ("gp", True)):
    t0 = time.perf_counter()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    gan.train_discriminator_step(dl_iter=it, apply_gradient_penalty=gp)
    e1.record()
    gan.train_generator_step(batch_size=B, dl_iter=it)
    e2.record()
    torch.cuda.synchronize()
    print(f"[{name}] D-step {e0.elapsed_time(e1):.1f} ms, G-step {e1.elapsed_time(e2):.1f} ms, wall {1e3*(time.perf_counter()-t0):.1f} ms", flush=True)
from torch.profiler import profile, ProfilerActivity
for name, gp in (("plain", False), ("gp", True)):
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        gan.train_discriminator_step(dl_iter=it, apply_gradient_penalty=gp)
        gan.train_generator_step(batch_size=B, dl_iter=it)
        torch.cuda.synchronize()
    print(f"===== {name} step: top kernels by CUDA time")
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=32, max_name_column_width=70))
    ka = [k for k in prof.key_averages() if k.device_type.name == "CUDA" or getattr(k, "self_device_time_total", 0) > 0]
    rows = sorted(((k.count, k.self_device_time_total / 1e3, k.key[:90]) for k in prof.key_averages()
                   if k.self_device_time_total > 0), reverse=True)[:45]
    print(f"===== {name} step: kernels/ops by launch count (count, self CUDA ms, name)")
    for r in rows:
        print(f"{r[0]:6d} {r[1]:9.3f}  {r[2]}")
# where do torch's own copy / cast / accumulate kernels come from?  (input shapes of the plain step)
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True) as prof:
    gan.train_discriminator_step(dl_iter=it, apply_gradient_penalty=False)
    gan.train_generator_step(batch_size=B, dl_iter=it)
    torch.cuda.synchronize()
print("===== plain step: aten copy/cast/add/fill by input shape (count, CUDA ms, op, shapes)")
rows = [(k.count, k.device_time_total / 1e3, k.key, str(k.input_shapes)[:110]) for k in prof.key_averages(group_by_input_shape=True)
        if k.key in ("aten::copy_", "aten::_to_copy", "aten::add_", "aten::add", "aten::fill_", "aten::cat", "aten::contiguous",
                     "aten::clone", "aten::zeros", "aten::zero_", "aten::mul", "aten::pad", "aten::constant_pad_nd")]
for r in sorted(rows, key=lambda r: -r[1])[:40]:
    print(f"{r[0]:6d} {r[1]:9.3f}  {r[2]:22s} {r[3]}")
