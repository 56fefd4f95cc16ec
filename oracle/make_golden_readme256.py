"""Writes tests/golden/ka8_readme256.pt from the UNMODIFIED reference at the BENCHMARKED configuration (README
unconditional 256: G dim_capacity 8 / D dim_capacity 16 / dim_max 512 / 4 skip-layer-excite), batch 2, CPU:

    python oracle/make_golden_readme256.py          # build container only (needs /root/reference); a few minutes

Seeded construction is bit-identical between the reference and the drop-in classes (tests/test_oracle_vs_reference.py),
so the 92 M parameters are NOT stored: the fixture holds the seeds, per-tensor checksums of the state_dicts, the
reference's outputs, and for every parameter the gradient's L2 norm plus 512 evenly spaced entries of it.  The same
quantities are computed a second time under CPU bf16 autocast (what `amp=True, mixed_precision_type='bf16'` runs) and
their deviation from the fp32 run is stored per tensor: the GPU bf16 tests bound |ours_bf16 - ref_fp32| by a stated
multiple of the reference's own |ref_bf16 - ref_fp32|.  Test infrastructure only.
"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "ref_shims"))
sys.path.insert(0, "/root/reference")
import gigagan_pytorch as ref  # noqa: E402
from gigagan_pytorch.gigagan_pytorch import discriminator_hinge_loss, gradient_penalty  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "ka8_readme256.pt")
GCFG = dict(dim_capacity=8, style_network=dict(dim=64, depth=4), image_size=256, dim_max=512,
            num_skip_layers_excite=4, unconditional=True)
DCFG = dict(dim_capacity=16, dim_max=512, image_size=256, num_skip_layers_excite=4, unconditional=True)
NS = 512


def rn(k, *s):
    return torch.randn(*s, generator=torch.Generator().manual_seed(k))


def sample_idx(n):
    return torch.linspace(0, n - 1, min(NS, n)).long()


def grad_summary(named):
    out = {}
    for k, p in named:
        if p.grad is None:
            continue
        g = p.grad.detach().float().flatten()
        out[k] = dict(norm=g.norm().item(), sample=g[sample_idx(g.numel())].clone())
    return out


def relmax(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-30)).item()


def build():
    torch.manual_seed(0)
    G = ref.Generator(**GCFG)
    torch.manual_seed(1)
    D = ref.Discriminator(**DCFG)
    with torch.no_grad():                      # make the per-layer noise path matter (weights are zero at init)
        gen = torch.Generator().manual_seed(11)
        for n, p in G.named_parameters():
            if n.endswith(".1.1.weight") or n.endswith(".1.4.weight"):
                p.copy_(torch.randn(p.shape, generator=gen) * 0.1)
    return G, D


def run_g(G, autocast):
    G.zero_grad()
    z = rn(1, 2, 64)
    torch.manual_seed(2)                        # the reference draws its layer noises from the CPU generator
    with torch.autocast("cpu", dtype=torch.bfloat16, enabled=autocast):
        rgb, rgbs = G(noise=z, return_all_rgbs=True)
        loss = (rgb.float() ** 2).mean()
    loss.backward()
    return dict(rgb=rgb.detach().float(), rgbs=[t.detach().float() for t in rgbs], loss=loss.detach(),
                grads=grad_summary(G.named_parameters()))


def run_d(D, autocast):
    """the discriminator step's objective as the reference trainer builds it (ref gigagan_pytorch.py:2318-2417):
    hinge + 0.1 * multiscale hinge + gradient penalty on real and fake, auxiliary reconstruction off (random patches)"""
    D.zero_grad()
    D.train()
    img = torch.rand(2, 3, 256, 256, generator=torch.Generator().manual_seed(3))
    fake = torch.rand(2, 3, 256, 256, generator=torch.Generator().manual_seed(4)) * 2 - 1
    r = img.clone().requires_grad_()
    f = fake.clone().requires_grad_()
    frgbs = [t.detach().requires_grad_() for t in D.real_images_to_rgbs(f)]
    with torch.autocast("cpu", dtype=torch.bfloat16, enabled=autocast):
        fl, fm, _ = D(f, frgbs, calc_aux_loss=False)
        rl, rm, _ = D(r, D.real_images_to_rgbs(r), calc_aux_loss=False)
        div = discriminator_hinge_loss(rl, fl)
        msl = sum(discriminator_hinge_loss(b, a) for a, b in zip(fm, rm))
        w = [1.0] + [0.1] * len(rm)
        gp = gradient_penalty(r, [rl, *rm], w) + gradient_penalty(f, [fl, *fm], w)
        total = div + gp + 0.1 * msl
    total.backward()
    return dict(real_logits=rl.detach().float(), fake_logits=fl.detach().float(),
                real_ms=[t.detach().float() for t in rm], fake_ms=[t.detach().float() for t in fm],
                loss=dict(total=total.detach().float(), divergence=div.detach().float(), multiscale=msl.detach().float(),
                          gradient_penalty=gp.detach().float()),
                grads=grad_summary(D.named_parameters()))


def deviation(lo, hi):
    """per-tensor deviation of the bf16-autocast run (lo) from the fp32 run (hi)"""
    dev = {}
    for k, v in hi.items():
        if k == "grads":
            dev["grads"] = {n: dict(norm_rel=abs(lo["grads"][n]["norm"] - g["norm"]) / max(g["norm"], 1e-30),
                                    sample_rel=relmax(lo["grads"][n]["sample"], g["sample"]))
                            for n, g in v.items()}
        elif isinstance(v, dict):
            dev[k] = {n: relmax(lo[k][n], t) for n, t in v.items()}
        elif isinstance(v, list):
            dev[k] = [relmax(a, b) for a, b in zip(lo[k], v)]
        else:
            dev[k] = relmax(lo[k], v)
    return dev


def main():
    t0 = time.time()
    G, D = build()
    checks = {"G": {k: v.double().abs().sum().item() for k, v in G.state_dict().items()},
              "D": {k: v.double().abs().sum().item() for k, v in D.state_dict().items()}}
    g32 = run_g(G, False)
    print("G fp32", time.time() - t0, flush=True)
    g16 = run_g(G, True)
    print("G bf16", time.time() - t0, flush=True)
    d32 = run_d(D, False)
    print("D fp32", time.time() - t0, flush=True)
    d16 = run_d(D, True)
    print("D bf16", time.time() - t0, flush=True)
    # keep the big image tensors compact: the full-resolution rgb in fp16 would lose the 2e-4 check -> store fp32 but only
    # the final rgb and the rgbs up to 64x64 (the 128/256 intermediate rgbs are covered through the final rgb)
    g32["rgbs"] = [t for t in g32["rgbs"] if t.shape[-1] <= 64]
    g16["rgbs"] = [t for t in g16["rgbs"] if t.shape[-1] <= 64]
    fx = dict(gcfg=GCFG, dcfg=DCFG, seeds=dict(G=0, D=1, noise_weights=11, z=1, layer_noise=2, img=3, fake=4),
              checksums=checks, g=g32, d=d32, g_bf16_dev=deviation(g16, g32), d_bf16_dev=deviation(d16, d32),
              torch_version=torch.__version__)
    torch.save(fx, OUT)
    print(OUT, os.path.getsize(OUT), "bytes", time.time() - t0, "s")
    print("bf16 deviations: rgb", fx["g_bf16_dev"]["rgb"], "real_logits", fx["d_bf16_dev"]["real_logits"],
          "gp", fx["d_bf16_dev"]["loss"]["gradient_penalty"])
    gd = fx["d_bf16_dev"]["grads"]
    worst = sorted(((v["sample_rel"], k) for k, v in gd.items()), reverse=True)[:5]
    print("worst D grad sample deviations (bf16 autocast vs fp32):", worst)


if __name__ == "__main__":
    main()
