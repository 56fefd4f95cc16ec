"""One warm-up step + one NVTX-marked plain training step (eager, no CUDA graph): the target of the ncu launch list."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gigagan_pytorch_b200 as g
from bench import G_CFG, D_CFG, real_batch
from gigagan_pytorch_b200.trainer import cycle
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
gan = g.GigaGAN(generator=dict(G_CFG), discriminator=dict(D_CFG), amp=True, mixed_precision_type="bf16", log_steps_every=10**9).to(dev)
class Pool:
    batch_size = B
    def __iter__(self):
        return iter([real_batch(0, 1, 0, B, 256).to(dev)])
it = cycle(Pool())
gan.train_discriminator_step(dl_iter=it, apply_gradient_penalty=False)
gan.train_generator_step(batch_size=B, dl_iter=it)
torch.cuda.synchronize()
torch.cuda.profiler.start()          # ncu --profile-from-start off: captures every thread (autograd worker too)
gan.train_discriminator_step(dl_iter=it, apply_gradient_penalty=False)
gan.train_generator_step(batch_size=B, dl_iter=it)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
