"""Diagnostics for two parity questions (not a test):
 (a) fp32 README-256 D step: are the few-1e-3 deviations of two bias gradients a property of the problem (LeakyReLU sign
     flips at |y|~0 under a different summation order) or of our kernels?  Runs the oracle (plain torch) on the GPU in
     fp32 and compares fixture / oracle-on-GPU / ours pairwise.
 (b) trainer: parameter updates of eager run A, eager run B (same seeds) and a CUDA-graph run.
usage: python tools/diag_parity.py [a] [b]"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
import gigagan_pytorch_b200 as g
from gigagan_pytorch_b200 import ops
dev = torch.device("cuda:0")
which = sys.argv[1:] or ["a", "b"]


def relmax(a, b):
    return ((a.float().cpu() - b.float().cpu()).abs().max() / b.float().abs().max().cpu().clamp_min(1e-30)).item()


if "a" in which:
    from oracle import gigagan_oracle as O
    from gigagan_pytorch_b200.trainer import discriminator_hinge_loss, gradient_penalty
    fx = torch.load(os.path.join(ROOT, "tests", "golden", "ka8_readme256.pt"), weights_only=False)
    g.set_compute_dtype(torch.float32)
    torch.manual_seed(1)
    D = g.Discriminator(**fx["dcfg"]).to(dev)
    D.train()
    img = torch.rand(2, 3, 256, 256, generator=torch.Generator().manual_seed(3)).to(dev)
    fake = (torch.rand(2, 3, 256, 256, generator=torch.Generator().manual_seed(4)) * 2 - 1).to(dev)
    # oracle on the GPU
    sd = {k: v.detach().clone().requires_grad_(v.dtype.is_floating_point) for k, v in D.state_dict().items()}
    plan = O.discriminator_plan(256, 16, 512, num_skip_layers_excite=4)
    r, f = img.clone().requires_grad_(), fake.clone().requires_grad_()
    frgbs = [t.detach() for t in O.real_images_to_rgbs(f, plan)]
    fl, fm, _ = O.discriminator_forward(sd, plan, f, frgbs, True, False)
    rl, rm, _ = O.discriminator_forward(sd, plan, r, O.real_images_to_rgbs(r, plan), True, False)
    w = [1.0] + [0.1] * len(rm)
    tot = (O.discriminator_hinge_loss(rl, fl) + 0.1 * sum(O.discriminator_hinge_loss(b, a) for a, b in zip(fm, rm))
           + O.gradient_penalty(r, [rl, *rm], w) + O.gradient_penalty(f, [fl, *fm], w))
    keys = [k for k, v in sd.items() if v.requires_grad]
    og = dict(zip(keys, torch.autograd.grad(tot, [sd[k] for k in keys], allow_unused=True)))
    # ours
    r, f = img.clone().requires_grad_(), fake.clone().requires_grad_()
    fn, rn_ = ops.to_nhwc(f, 3, torch.float32), ops.to_nhwc(r, 3, torch.float32)
    frgbs = [t.detach() for t in D.real_images_to_rgbs_nhwc(fn)]
    fl, fm, _ = D.forward_nhwc(fn, frgbs, True, False, fused_attention=False)
    rl, rm, _ = D.forward_nhwc(rn_, D.real_images_to_rgbs_nhwc(rn_), True, False, fused_attention=False)
    total = (discriminator_hinge_loss(rl, fl) + 0.1 * sum(discriminator_hinge_loss(b, a) for a, b in zip(fm, rm))
             + gradient_penalty(r, [rl, *rm], w) + gradient_penalty(f, [fl, *fm], w))
    total.backward()
    named = dict(D.named_parameters())
    rows = []
    for k, ref in fx["d"]["grads"].items():
        if og.get(k) is None:
            continue
        idx = torch.linspace(0, named[k].numel() - 1, min(512, named[k].numel())).long().to(dev)
        ours, orc = named[k].grad.flatten()[idx], og[k].flatten()[idx]
        rows.append((relmax(ours, ref["sample"]), relmax(orc, ref["sample"]), relmax(ours, orc), k))
    rows.sort(reverse=True)
    print("[a] fp32 D step README-256, gradient samples: ours-vs-fixture | oracle(GPU torch)-vs-fixture | ours-vs-oracle(GPU)")
    for rr in rows[:10]:
        print(f"   {rr[0]:.2e} {rr[1]:.2e} {rr[2]:.2e}  {rr[3]}")
    for k in ("layers.5.5.layers.0.2.bias", "layers.5.2.0.bias"):
        gr, o = named[k].grad, og[k]
        d = (gr - o).abs()
        print(f"   {k}: |ours-oracle| top5 {[f'{v:.3e}' for v in d.flatten().topk(5).values.tolist()]} max|grad| {o.abs().max().item():.3e}")

if "b" in which:
    def run(upsampler, amp, graphs):
        g.set_compute_dtype(torch.float32)
        torch.manual_seed(0)
        if upsampler:
            gen = dict(dim=8, image_size=64, input_image_size=16, style_network=dict(dim=16, depth=2), dim_mults=(1, 2, 4),
                       full_attn=(False, False, True), cross_attn=(False, False, True), attn_depths=(1, 1, 1),
                       self_attn_dim_head=8, self_attn_heads=2, cross_attn_dim_head=8, unconditional=True)
            disc = dict(dim_capacity=2, dim_max=16, image_size=64, num_skip_layers_excite=2, unconditional=True,
                        attn_resolutions=(8,), attn_dim_head=8, attn_heads=2, multiscale_input_resolutions=(32, 16))
        else:
            gen = dict(dim_capacity=2, style_network=dict(dim=16, depth=2), image_size=64, dim_max=16, dim_latent=16,
                       num_skip_layers_excite=2, unconditional=True, self_attn_resolutions=(16,), self_attn_dim_head=8,
                       self_attn_heads=2)
            disc = dict(dim_capacity=2, dim_max=16, image_size=64, num_skip_layers_excite=2, unconditional=True,
                        attn_resolutions=(8,), attn_dim_head=8, attn_heads=2, multiscale_input_resolutions=(32, 16, 8))
        gan = g.GigaGAN(generator=gen, discriminator=disc, train_upsampler=upsampler, amp=amp, mixed_precision_type="bf16",
                        log_steps_every=10 ** 9, create_ema_generator_at_init=False, save_and_sample_every=0).to(dev)
        gan.use_cuda_graphs = graphs
        reals = [torch.rand(4, 3, 64, 64, generator=torch.Generator().manual_seed(10 + s)).to(dev) for s in range(8)]

        class Pool:
            batch_size = 4

            def __iter__(self):
                return iter(reals)
        from gigagan_pytorch_b200.trainer import cycle
        it = cycle(Pool())
        torch.manual_seed(5)
        names = [("G." + n) for n, _ in gan.G.named_parameters()] + [("D." + n) for n, _ in gan.D.named_parameters()]
        flat = lambda: [p.detach().flatten().float().clone() for p in list(gan.G.parameters()) + list(gan.D.parameters())]
        p0 = flat()
        losses = []
        for step in range(1, 6):
            d = gan.train_discriminator_step(dl_iter=it, apply_gradient_penalty=step % 4 == 0)
            gl = gan.train_generator_step(batch_size=4, dl_iter=it)
            losses.append((float(d.divergence), float(d.gradient_penalty), float(gl.divergence)))
        torch.cuda.synchronize()
        return names, [b - a for a, b in zip(p0, flat())], losses

    for upsampler, amp in ((False, True), (True, False)):
        names, a, la = run(upsampler, amp, False)
        _, b, lb = run(upsampler, amp, False)
        _, c, lc = run(upsampler, amp, True)
        mx = max(t.abs().max().item() for t in a)
        print(f"[b] upsampler={upsampler} amp={amp}: max |delta| {mx:.3e}")
        for tag, other, lo in (("eager B", b, lb), ("graph", c, lc)):
            per = sorted(((x - y).abs().max().item() / mx, n) for n, x, y in zip(names, a, other))[::-1]
            print(f"   eager A vs {tag}: worst {[(f'{v:.3f}', n) for v, n in per[:4]]}")
            print(f"      losses A {la}\n      losses {tag} {lo}")
