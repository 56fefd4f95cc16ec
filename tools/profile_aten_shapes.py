"""Which ATen operators (autograd accumulations, copies, fills, cats) a training step still launches, by input shape and by
the Python frame that issued them: torch.profiler with record_shapes + with_stack over one eager step of each kind.
usage: python tools/profile_aten_shapes.py [gp|plain|g]"""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gigagan_pytorch_b200 as g
from bench import G_CFG, D_CFG, real_batch
from gigagan_pytorch_b200.trainer import cycle
from torch.profiler import profile, ProfilerActivity

which = sys.argv[1] if len(sys.argv) > 1 else "gp"
size, B = 256, 16
dev = torch.device("cuda:0")
torch.manual_seed(0)
gan = g.GigaGAN(generator=dict(G_CFG, image_size=size), discriminator=dict(D_CFG, image_size=size), amp=True,
                mixed_precision_type="bf16", log_steps_every=10 ** 9, save_and_sample_every=0).to(dev)
gan.use_cuda_graphs = False


class Pool:
    batch_size = B

    def __iter__(self):
        return iter([real_batch(s, 1, 0, B, size).to(dev) for s in range(2)])


it = cycle(Pool())
fn = {"gp": lambda: gan.train_discriminator_step(dl_iter=it, apply_gradient_penalty=True),
      "plain": lambda: gan.train_discriminator_step(dl_iter=it, apply_gradient_penalty=False),
      "g": lambda: gan.train_generator_step(batch_size=B, dl_iter=it)}[which]
for _ in range(2):
    fn()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True, with_stack=True) as prof:
    fn()
    torch.cuda.synchronize()
WANT = ("aten::copy_", "aten::add_", "aten::add", "aten::fill_", "aten::cat", "aten::zero_", "aten::contiguous", "aten::clone",
        "aten::_to_copy", "aten::mul", "aten::sum")
rows = collections.defaultdict(lambda: [0.0, 0])
for e in prof.events():
    if e.name in WANT and e.device_time_total > 0:
        frame = "?"
        for s in (e.stack or []):
            if "gigagan_pytorch_b200" in s or "trainer.py" in s:
                frame = s.split("/")[-1][:60]
                break
        key = (e.name, str(e.input_shapes)[:90], frame)
        rows[key][0] += e.device_time_total / 1e3
        rows[key][1] += 1
tot = sum(v[0] for v in rows.values())
print(f"== {which}: {tot:.2f} ms in {sum(v[1] for v in rows.values())} ATen calls (ms, calls, op, shapes, frame)")
for k, v in sorted(rows.items(), key=lambda kv: -kv[1][0])[:45]:
    print(f"{v[0]:8.3f} {v[1]:4d}  {k[0]:14s} {k[1]:90s} {k[2]}")
