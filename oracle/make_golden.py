"""Writes tests/golden/*.pt from the UNMODIFIED reference (run in the build container only):
    python oracle/make_golden.py
Each fixture holds a tiny model's state_dict, the seeded inputs and the reference's outputs /
gradients, so that the GPU box (which has no /root/reference) can check oracle and CUDA path."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "ref_shims"))
sys.path.insert(0, "/root/reference")
import gigagan_pytorch as ref  # noqa: E402
from gigagan_pytorch.gigagan_pytorch import (SelfAttentionBlock, discriminator_hinge_loss,  # noqa: E402
                                             gradient_penalty)

OUT = os.path.join(ROOT, "tests", "golden")


def rn(k, *s):
    return torch.randn(*s, generator=torch.Generator().manual_seed(k))


def sd_of(m):
    return {k: v.detach().clone() for k, v in m.state_dict().items()}


def main():
    os.makedirs(OUT, exist_ok=True)
    # ---- KA1: AdaptiveConv2DMod fwd+bwd (SURVEY 8c KA1 recipe)
    torch.manual_seed(0)
    m = ref.AdaptiveConv2DMod(8, 6, 3, num_conv_kernels=2)
    x = rn(1, 2, 8, 5, 5).requires_grad_()
    mod = rn(2, 2, 8).requires_grad_()
    km = rn(3, 2, 2).requires_grad_()
    y = m(x, mod=mod, kernel_mod=km)
    (y ** 2).sum().backward()
    torch.save(dict(weights=m.weights.detach().clone(), x=x.detach(), mod=mod.detach(), kernel_mod=km.detach(),
                    y=y.detach(), dweights=m.weights.grad.clone(), dx=x.grad.clone(), dmod=mod.grad.clone(),
                    dkernel_mod=km.grad.clone()), os.path.join(OUT, "ka1_adaptive_conv.pt"))
    # ---- KA2: SelfAttentionBlock (L2 and dot)
    for dot in (False, True):
        torch.manual_seed(0)
        blk = SelfAttentionBlock(16, dim_head=8, heads=2, dot_product=dot)
        x = rn(1, 2, 16, 4, 4).requires_grad_()
        y = blk(x)
        (y ** 2).sum().backward()
        torch.save(dict(sd=sd_of(blk), x=x.detach(), y=y.detach(), dx=x.grad.clone(),
                        grads={k: p.grad.clone() for k, p in blk.named_parameters()}),
                   os.path.join(OUT, f"ka2_attn_block_{'dot' if dot else 'l2'}.pt"))
    # ---- KA3: StyleNetwork
    torch.manual_seed(0)
    sn = ref.StyleNetwork(dim=64, depth=4)
    z = rn(1, 2, 64)
    torch.save(dict(sd=sd_of(sn), z=z, y=sn(z).detach()), os.path.join(OUT, "ka3_style_network.pt"))
    # ---- KA4: tiny Generator (image 32, capacity 2) fwd + grads of sum(rgb^2)
    gcfg = dict(dim_capacity=2, style_network=dict(dim=16, depth=2), image_size=32, dim_max=16, dim_latent=16,
                num_skip_layers_excite=2, unconditional=True, self_attn_resolutions=(16,),
                self_attn_dim_head=8, self_attn_heads=2)
    torch.manual_seed(0)
    G = ref.Generator(**gcfg)
    with torch.no_grad():
        for n, p in G.named_parameters():
            if n.endswith(".1.1.weight") or n.endswith(".1.4.weight"):
                p.copy_(torch.randn_like(p) * 0.1)
    z = rn(1, 2, 16)
    torch.manual_seed(2)
    rgb, rgbs = G(noise=z, return_all_rgbs=True)
    (rgb ** 2).mean().backward()
    torch.save(dict(cfg=gcfg, sd=sd_of(G), z=z, noise_seed=2, rgb=rgb.detach(), rgbs=[t.detach() for t in rgbs],
                    grads={k: p.grad.clone() for k, p in G.named_parameters() if p.grad is not None}),
               os.path.join(OUT, "ka4_generator.pt"))
    # ---- KA5: tiny Discriminator (image 32) fwd, D-step objective with GP, param grads
    dcfg = dict(dim_capacity=2, dim_max=16, image_size=32, num_skip_layers_excite=2, unconditional=True,
                attn_resolutions=(8,), attn_dim_head=8, attn_heads=2, multiscale_input_resolutions=(16, 8),
                aux_recon_resolutions=(8,))
    torch.manual_seed(0)
    D = ref.Discriminator(**dcfg)
    img = torch.rand(2, 3, 32, 32, generator=torch.Generator().manual_seed(3))
    fake = torch.rand(2, 3, 32, 32, generator=torch.Generator().manual_seed(4))
    with torch.no_grad():
        D.eval()
        logits, ms, _ = D(img, D.real_images_to_rgbs(img), calc_aux_loss=False)
        D.train()
    r = img.clone().requires_grad_()
    f = fake.clone().requires_grad_()
    frgbs = [t.detach().requires_grad_() for t in D.real_images_to_rgbs(f)]
    fl, fm, _ = D(f, frgbs, calc_aux_loss=False)
    rl, rm, _ = D(r, D.real_images_to_rgbs(r), calc_aux_loss=False)
    div = discriminator_hinge_loss(rl, fl)
    msl = sum(discriminator_hinge_loss(b, a) for a, b in zip(fm, rm))
    w = [1.0] + [0.1] * len(rm)
    gp = gradient_penalty(r, [rl, *rm], w) + gradient_penalty(f, [fl, *fm], w)
    total = div + gp + 0.1 * msl
    total.backward()
    torch.save(dict(cfg=dcfg, sd=sd_of(D), img=img, fake=fake, logits=logits, ms=ms,
                    loss=dict(total=total.detach(), divergence=div.detach(), multiscale=msl.detach(),
                              gradient_penalty=gp.detach()),
                    grads={k: p.grad.clone() for k, p in D.named_parameters() if p.grad is not None}),
               os.path.join(OUT, "ka5_discriminator.pt"))
    # ---- KA6: tiny UnetUpsampler (16 -> 32) fwd + grads of mean(rgb^2)
    ucfg = dict(dim=4, image_size=32, input_image_size=16, style_network=dict(dim=16, depth=2), dim_mults=(1, 2, 4),
                full_attn=(False, False, True), cross_attn=(False, False, True), attn_depths=(1, 1, 1),
                self_attn_dim_head=8, self_attn_heads=2, cross_attn_dim_head=8, unconditional=True)
    torch.manual_seed(0)
    U = ref.UnetUpsampler(**ucfg)
    low = torch.rand(2, 3, 16, 16, generator=torch.Generator().manual_seed(3))
    z = rn(1, 2, 16)
    rgb, rgbs = U(low, noise=z, return_all_rgbs=True)
    (rgb ** 2).mean().backward()
    torch.save(dict(cfg=ucfg, sd=sd_of(U), low=low, z=z, rgb=rgb.detach(), rgbs=[t.detach() for t in rgbs],
                    grads={k: p.grad.clone() for k, p in U.named_parameters() if p.grad is not None}),
               os.path.join(OUT, "ka6_unet_upsampler.pt"))
    # ---- KA7: text-conditioned generator + discriminator (TextEncoder on given CLIP-style encodings; the
    #      CLIP tower itself is a third-party dependency outside the path, so text_encodings are the boundary)
    from gigagan_pytorch.gigagan_pytorch import TextEncoder as RefTextEncoder
    from gigagan_pytorch.open_clip import OpenClipAdapter

    class _NoClip(OpenClipAdapter):                       # only .dim_latent is consulted when encodings are given
        def __init__(self):
            torch.nn.Module.__init__(self)

        @property
        def dim_latent(self):
            return 32

    tecfg = dict(dim=24, depth=1, dim_head=8, heads=2)
    gcfg7 = dict(dim_capacity=2, style_network=dict(dim=16, depth=2, dim_text_latent=24), image_size=32, dim_max=16,
                 dim_latent=16, num_skip_layers_excite=2, self_attn_resolutions=(16,), self_attn_dim_head=8,
                 self_attn_heads=2, cross_attn_resolutions=(16, 8), cross_attn_dim_head=8, cross_attn_heads=2,
                 unconditional=False)
    dcfg7 = dict(dim_capacity=2, dim_max=16, image_size=32, num_skip_layers_excite=2, attn_resolutions=(8,),
                 attn_dim_head=8, attn_heads=2, multiscale_input_resolutions=(16, 8), unconditional=False)
    torch.manual_seed(0)
    G7 = ref.Generator(text_encoder=RefTextEncoder(clip=_NoClip(), **tecfg), **gcfg7)
    torch.manual_seed(1)
    D7 = ref.Discriminator(text_encoder=RefTextEncoder(clip=_NoClip(), **tecfg), **dcfg7)
    enc = rn(5, 2, 6, 32)
    enc[1, 4:] = 0.                                        # padded tokens -> mask False (ref :851)
    z = rn(1, 2, 16)
    torch.manual_seed(2)
    rgb, rgbs = G7(noise=z, text_encodings=enc, return_all_rgbs=True)
    (rgb ** 2).mean().backward()
    img = torch.rand(2, 3, 32, 32, generator=torch.Generator().manual_seed(4))
    real_rgbs = D7.real_images_to_rgbs(img)
    logits, ms, _ = D7(img, real_rgbs, text_encodings=enc, return_multiscale_outputs=True, calc_aux_loss=False)
    (logits.sum() + sum((m_ ** 2).sum() for m_ in ms)).backward()
    torch.save(dict(te_cfg=dict(clip_dim_latent=32, **tecfg), gcfg=gcfg7, dcfg=dcfg7, gsd=sd_of(G7), dsd=sd_of(D7),
                    enc=enc, z=z, noise_seed=2, rgb=rgb.detach(), rgbs=[t.detach() for t in rgbs], img=img,
                    logits=logits.detach(), ms=[t.detach() for t in ms],
                    ggrads={k: p.grad.clone() for k, p in G7.named_parameters() if p.grad is not None},
                    dgrads={k: p.grad.clone() for k, p in D7.named_parameters() if p.grad is not None}),
               os.path.join(OUT, "ka7_text_conditional.pt"))
    make_aux_decoder()
    make_text_step()
    for f_ in sorted(os.listdir(OUT)):
        print(f_, os.path.getsize(os.path.join(OUT, f_)))


def make_aux_decoder():
    """KA5b: the auxiliary reconstruction decoder (SimpleDecoder, ref :1290-1317 via :1812-1827), calc_aux_loss=True.
    eval mode (no dropout) so that the only randomness is the CPU randn patch permutation (:1310), drawn right after
    manual_seed(patch_seed)."""
    dcfg = dict(dim_capacity=2, dim_max=16, image_size=32, num_skip_layers_excite=2, unconditional=True,
                attn_resolutions=(8,), attn_dim_head=8, attn_heads=2, multiscale_input_resolutions=(16, 8),
                aux_recon_resolutions=(8,))
    torch.manual_seed(0)
    D = ref.Discriminator(**dcfg)
    D.eval()
    img = torch.rand(2, 3, 32, 32, generator=torch.Generator().manual_seed(3))
    torch.manual_seed(7)
    logits, ms, aux = D(img, D.real_images_to_rgbs(img), calc_aux_loss=True)
    assert len(aux) == 1
    aux[0].backward()
    grads = {k: p.grad.clone() for k, p in D.named_parameters() if p.grad is not None}
    # the same quantities under CPU bf16 autocast (what amp=True runs): their deviation from the fp32 run is the yardstick
    # of the GPU bf16 test (|ours_bf16 - ref_fp32| <= K * |ref_bf16 - ref_fp32| per tensor)
    D.zero_grad()
    torch.manual_seed(7)
    with torch.autocast("cpu", dtype=torch.bfloat16):
        _, _, aux16 = D(img, D.real_images_to_rgbs(img), calc_aux_loss=True)
    aux16[0].float().backward()
    rel = lambda a, b: ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-30)).item()
    dev = dict(aux=rel(aux16[0].detach(), aux[0].detach()),
               grads={k: rel(p.grad, grads[k]) for k, p in D.named_parameters() if p.grad is not None})
    torch.save(dict(cfg=dcfg, sd=sd_of(D), img=img, patch_seed=7, aux=[a.detach() for a in aux], grads=grads,
                    bf16_dev=dev), os.path.join(OUT, "ka5b_aux_decoder.pt"))


def make_text_step():
    """KA9: one text-conditional discriminator step objective (with gradient penalty) and one generator step objective
    built from the reference's Generator / Discriminator exactly as its trainer does (ref :2263-2417, :2518-2551), on
    pre-encoded text_encodings (the OpenCLIP tower is outside the path).  Auxiliary losses off."""
    from gigagan_pytorch.gigagan_pytorch import TextEncoder as RefTextEncoder, generator_hinge_loss
    from gigagan_pytorch.open_clip import OpenClipAdapter

    class _NoClip(OpenClipAdapter):
        def __init__(self):
            torch.nn.Module.__init__(self)

        @property
        def dim_latent(self):
            return 32

    tecfg = dict(dim=24, depth=1, dim_head=8, heads=2)
    gcfg = dict(dim_capacity=2, style_network=dict(dim=16, depth=2, dim_text_latent=24), image_size=32, dim_max=16,
                dim_latent=16, num_skip_layers_excite=2, self_attn_resolutions=(16,), self_attn_dim_head=8,
                self_attn_heads=2, cross_attn_resolutions=(16, 8), cross_attn_dim_head=8, cross_attn_heads=2,
                unconditional=False)
    dcfg = dict(dim_capacity=2, dim_max=16, image_size=32, num_skip_layers_excite=2, attn_resolutions=(8,),
                attn_dim_head=8, attn_heads=2, multiscale_input_resolutions=(16, 8), unconditional=False)
    torch.manual_seed(0)
    G = ref.Generator(text_encoder=RefTextEncoder(clip=_NoClip(), **tecfg), **gcfg)
    torch.manual_seed(1)
    D = ref.Discriminator(text_encoder=RefTextEncoder(clip=_NoClip(), **tecfg), **dcfg)
    enc = rn(5, 2, 6, 32)
    enc[1, 4:] = 0.
    z = rn(1, 2, 16)
    real = torch.rand(2, 3, 32, 32, generator=torch.Generator().manual_seed(4)).requires_grad_()
    G.train(); D.train()
    with torch.no_grad():
        torch.manual_seed(2)
        fake, rgbs = G(noise=z, text_encodings=enc, return_all_rgbs=True)
    fake = fake.detach().requires_grad_()
    rgbs = [t.detach().requires_grad_() for t in rgbs]
    fl, fm, _ = D(fake, rgbs, text_encodings=enc, calc_aux_loss=False)
    rl, rm, _ = D(real, D.real_images_to_rgbs(real), text_encodings=enc, calc_aux_loss=False)
    div = discriminator_hinge_loss(rl, fl)
    msl = sum(discriminator_hinge_loss(b, a) for a, b in zip(fm, rm))
    w = [1.0] + [0.1] * len(rm)
    gp = gradient_penalty(real, [rl, *rm], w) + gradient_penalty(fake, [fl, *fm], w)
    total = div + gp + 0.1 * msl
    total.backward()
    dgrads = {k: p.grad.clone() for k, p in D.named_parameters() if p.grad is not None}
    dloss = dict(total=total.detach(), divergence=div.detach(), multiscale=msl.detach(), gradient_penalty=gp.detach())
    D.zero_grad()
    torch.manual_seed(2)
    fake, rgbs = G(noise=z, text_encodings=enc, return_all_rgbs=True)
    logits, ms, _ = D(fake, rgbs, text_encodings=enc, calc_aux_loss=False)
    gdiv = generator_hinge_loss(logits)
    gms = sum(generator_hinge_loss(m_) for m_ in ms)
    gtotal = gdiv + 0.1 * gms
    gtotal.backward()
    ggrads = {k: p.grad.clone() for k, p in G.named_parameters() if p.grad is not None}
    torch.save(dict(te_cfg=dict(clip_dim_latent=32, **tecfg), gcfg=gcfg, dcfg=dcfg, gsd=sd_of(G), dsd=sd_of(D), enc=enc,
                    z=z, real=real.detach(), dloss=dloss, dgrads=dgrads,
                    gloss=dict(total=gtotal.detach(), divergence=gdiv.detach(), multiscale=gms.detach()), ggrads=ggrads),
               os.path.join(OUT, "ka9_text_step.pt"))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "aux":
        make_aux_decoder()
    elif len(sys.argv) > 1 and sys.argv[1] == "text_step":
        make_text_step()
    else:
        main()
