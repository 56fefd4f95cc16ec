"""Loss-trajectory sanity check: this repo's trainer (GPU, fp32 and bf16) vs the oracle trainer (CPU) from identical
initial weights and real batches (latent/noise draws differ: device vs host RNG), 64x64 config."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gigagan_pytorch_b200 as g
from gigagan_pytorch_b200.trainer import cycle
from oracle import gigagan_oracle as O

GC = dict(dim_capacity=4, style_network=dict(dim=64, depth=4), image_size=64, dim_max=512, num_skip_layers_excite=4, unconditional=True)
DC = dict(dim_capacity=4, dim_max=512, image_size=64, num_skip_layers_excite=4, unconditional=True)
B, STEPS = 4, int(sys.argv[1]) if len(sys.argv) > 1 else 12
reals = [torch.rand(B, 3, 64, 64, generator=torch.Generator().manual_seed(100 + s)) for s in range(STEPS)]

torch.manual_seed(0)
G0, D0 = g.Generator(**GC), g.Discriminator(**DC)
sdg, sdd = {k: v.clone() for k, v in G0.state_dict().items()}, {k: v.clone() for k, v in D0.state_dict().items()}

torch.manual_seed(1)
tr = O.OracleTrainer(sdg, O.generator_plan(64, 4, 512, num_skip_layers_excite=4), sdd, O.discriminator_plan(64, 4, 512, num_skip_layers_excite=4))
print("oracle (CPU fp32): step, D loss, G loss")
for s in range(STEPS):
    d, gl = tr.step(reals[s], (s + 1) % 4 == 0)
    print(f"  {s+1:3d} {d:12.4f} {gl:12.4f}", flush=True)

for amp in (False, True):
    g.set_compute_dtype(torch.float32)
    torch.manual_seed(0)
    gan = g.GigaGAN(generator=dict(GC), discriminator=dict(DC), amp=amp, mixed_precision_type="bf16", log_steps_every=10 ** 9).cuda()
    gan.G.load_state_dict(sdg); gan.D.load_state_dict(sdd)
    class Pool:
        batch_size = B
        def __iter__(self):
            return iter(reals)
    it = cycle(Pool())
    torch.manual_seed(1)
    print(f"ours (GPU {'bf16' if amp else 'fp32'}): step, D total-ish (div+0.1ms+gp+aux), G loss")
    for s in range(STEPS):
        d = gan.train_discriminator_step(dl_iter=it, apply_gradient_penalty=(s + 1) % 4 == 0)
        gl = gan.train_generator_step(batch_size=B, dl_iter=it)
        dt = float(d.divergence) + 0.1 * float(d.multiscale_divergence) + float(d.gradient_penalty) + float(d.aux_reconstruction)
        gt = float(gl.divergence) + 0.1 * float(gl.multiscale_divergence)
        print(f"  {s+1:3d} {dt:12.4f} {gt:12.4f}", flush=True)
