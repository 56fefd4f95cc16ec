"""Pins the oracle: every oracle function vs the unmodified reference module on the same
state_dict / seeds (fp32 CPU).  Only runs where /root/reference exists (build container)."""
import pytest
import torch

from conftest import import_reference
from oracle import gigagan_oracle as O

pytestmark = pytest.mark.reference
TOL = dict(rtol=1e-4, atol=1e-5)


def rn(k, *s):
    return torch.randn(*s, generator=torch.Generator().manual_seed(k))


def test_adaptive_conv():
    ref = import_reference()
    torch.manual_seed(0)
    m = ref.AdaptiveConv2DMod(8, 6, 3, num_conv_kernels=2)
    x, mod, km = rn(1, 4, 8, 5, 5), rn(2, 2, 8), rn(3, 2, 2)      # mod batch 2 repeated over scale (b=4)
    torch.testing.assert_close(O.adaptive_conv2d_mod(m.weights, x, mod, km), m(x, mod=mod, kernel_mod=km), **TOL)
    m1 = ref.AdaptiveConv2DMod(8, 3, 1, num_conv_kernels=1, demod=False)
    torch.testing.assert_close(O.adaptive_conv2d_mod(m1.weights, x, mod, rn(4, 2, 0), demod=False),
                               m1(x, mod=mod, kernel_mod=rn(4, 2, 0)), **TOL)


@pytest.mark.parametrize("dot", [False, True])
def test_self_attention_block(dot):
    ref = import_reference()
    from gigagan_pytorch.gigagan_pytorch import SelfAttentionBlock
    torch.manual_seed(0)
    m = SelfAttentionBlock(16, dim_head=8, heads=2, dot_product=dot)
    x = rn(1, 2, 16, 4, 4)
    torch.testing.assert_close(O.self_attention_block(dict(m.state_dict()), x, dot, heads=2, dim_head=8), m(x), **TOL)


def test_style_network():
    ref = import_reference()
    torch.manual_seed(0)
    m = ref.StyleNetwork(dim=64, depth=4)
    z = rn(1, 2, 64)
    torch.testing.assert_close(O.style_network(dict(m.state_dict()), z, 4), m(z), **TOL)


GCFG = dict(dim_capacity=4, style_network=dict(dim=64, depth=4), image_size=64, dim_max=512,
            num_skip_layers_excite=4, unconditional=True)
DCFG = dict(dim_capacity=4, dim_max=512, image_size=64, num_skip_layers_excite=4, unconditional=True)


def test_generator():
    ref = import_reference()
    torch.manual_seed(0)
    G = ref.Generator(**GCFG)
    with torch.no_grad():   # make the noise path matter
        for n, p in G.named_parameters():
            if n.endswith(".1.1.weight") or n.endswith(".1.4.weight"):
                p.copy_(torch.randn_like(p) * 0.1)
    plan = O.generator_plan(64, 4, 512, num_skip_layers_excite=4)
    z = rn(1, 2, 64)
    torch.manual_seed(2)
    rgb_ref, rgbs_ref = G(noise=z, return_all_rgbs=True)
    torch.manual_seed(2)
    rgb, rgbs = O.generator_forward(dict(G.state_dict()), plan, z, return_all_rgbs=True)
    torch.testing.assert_close(rgb, rgb_ref, **TOL)
    for a, b in zip(rgbs, rgbs_ref):
        torch.testing.assert_close(a, b, **TOL)


def test_discriminator_and_losses():
    ref = import_reference()
    from gigagan_pytorch.gigagan_pytorch import gradient_penalty, discriminator_hinge_loss
    torch.manual_seed(0)
    D = ref.Discriminator(**DCFG)
    plan = O.discriminator_plan(64, 4, 512, num_skip_layers_excite=4)
    sd = {k: v.detach().clone().requires_grad_() for k, v in D.state_dict().items()}
    img = torch.rand(2, 3, 64, 64, generator=torch.Generator().manual_seed(3))
    fake = torch.rand(2, 3, 64, 64, generator=torch.Generator().manual_seed(4))

    # --- forward incl. aux recon (same RNG order: dropout on device then CPU randn perm)
    torch.manual_seed(5)
    lr, mr, ar = D(img, D.real_images_to_rgbs(img), calc_aux_loss=True)
    torch.manual_seed(5)
    lo, mo, ao = O.discriminator_forward(sd, plan, img, O.real_images_to_rgbs(img, plan), True, True)
    torch.testing.assert_close(lo, lr, **TOL)
    for a, b in zip(mo, mr):
        torch.testing.assert_close(a, b, **TOL)
    torch.testing.assert_close(ao[0], ar[0], **TOL)

    # --- full D-step objective with gradient penalty: loss and parameter gradients
    def ref_loss():
        r = img.clone().requires_grad_()
        f = fake.clone().requires_grad_()
        frgbs = [t.detach().requires_grad_() for t in D.real_images_to_rgbs(f)]
        torch.manual_seed(7)
        fl, fm, _ = D(f, frgbs, calc_aux_loss=False)
        rl, rm, aux = D(r, D.real_images_to_rgbs(r), calc_aux_loss=True)
        div = discriminator_hinge_loss(rl, fl)
        ms = sum(discriminator_hinge_loss(b, a) for a, b in zip(fm, rm))
        w = [1.0] + [0.1] * len(rm)
        gp = gradient_penalty(r, [rl, *rm], w) + gradient_penalty(f, [fl, *fm], w)
        return div + gp + 0.1 * ms + sum(aux)

    D.zero_grad()
    lref = ref_loss()
    lref.backward()
    r = img.clone().requires_grad_()
    f = fake.clone().requires_grad_()
    frgbs = [t.detach().requires_grad_() for t in O.real_images_to_rgbs(f, plan)]
    torch.manual_seed(7)
    lor, parts = O.discriminator_step_loss(sd, plan, r, f, frgbs, True)
    lor.backward()
    torch.testing.assert_close(lor, lref, **TOL)
    for n, p in D.named_parameters():
        if p.grad is None:
            assert sd[n].grad is None or sd[n].grad.abs().max() == 0, n
            continue
        err = (sd[n].grad - p.grad).abs().max().item()
        assert err <= 1e-3 * p.grad.abs().max().item() + 1e-6, (n, err)   # fp32 double-backward noise


def test_adamw_matches_reference_optimizer():
    ref = import_reference()
    from gigagan_pytorch.optimizer import get_optimizer
    torch.manual_seed(0)
    ps = [torch.nn.Parameter(torch.randn(4, 3)), torch.nn.Parameter(torch.randn(5))]
    opt = get_optimizer(ps, lr=2e-4, betas=(0.5, 0.9), weight_decay=0.)     # wd swallowed -> 1e-2
    mine = [(p.detach().clone(), torch.zeros_like(p), torch.zeros_like(p)) for p in ps]
    for step in range(1, 4):
        gs = [torch.randn_like(p) for p in ps]
        for p, g in zip(ps, gs):
            p.grad = g.clone()
        opt.step()
        for (p, m, v), g in zip(mine, gs):
            O.adamw_step(p, g, m, v, step)
    for (p, _, _), q in zip(mine, ps):
        torch.testing.assert_close(p, q.detach(), rtol=1e-6, atol=1e-7)


def test_unet_upsampler():
    ref = import_reference()
    torch.manual_seed(0)
    U = ref.UnetUpsampler(dim=8, image_size=64, input_image_size=16, style_network=dict(dim=64, depth=4), unconditional=True)
    plan = O.unet_plan(8, 64, 16)
    assert plan["split"] == U.style_embed_split_dims
    x = torch.rand(2, 3, 16, 16, generator=torch.Generator().manual_seed(3))
    z = rn(1, 2, 64)
    with torch.no_grad():
        r, rs = U(x, noise=z, return_all_rgbs=True)
        o, os_ = O.unet_forward(dict(U.state_dict()), plan, x, z, return_all_rgbs=True)
    torch.testing.assert_close(o, r, **TOL)
    for a, b in zip(os_, rs):
        torch.testing.assert_close(a, b, **TOL)


# ------------------------------------------------------------------ text-conditioned path
def _no_clip(ref, dim_latent=32):
    """OpenClipAdapter stand-in: with pre-encoded tokens only `.dim_latent` is consulted (the CLIP tower is third-party)."""
    from gigagan_pytorch.open_clip import OpenClipAdapter

    class NoClip(OpenClipAdapter):
        def __init__(self):
            torch.nn.Module.__init__(self)

        @property
        def dim_latent(self):
            return dim_latent

    return NoClip()


TE = dict(dim=24, depth=2, dim_head=8, heads=2)
G7 = dict(dim_capacity=2, style_network=dict(dim=16, depth=2, dim_text_latent=24), image_size=32, dim_max=16, dim_latent=16,
          num_skip_layers_excite=2, self_attn_resolutions=(16,), self_attn_dim_head=8, self_attn_heads=2,
          cross_attn_resolutions=(16, 8), cross_attn_dim_head=8, cross_attn_heads=2, unconditional=False)
D7 = dict(dim_capacity=2, dim_max=16, image_size=32, num_skip_layers_excite=2, attn_resolutions=(8,), attn_dim_head=8,
          attn_heads=2, multiscale_input_resolutions=(16, 8), unconditional=False)


def _encodings():
    enc = rn(5, 3, 7, 32)
    enc[1, 4:] = 0.
    enc[2, 1:] = 0.
    return enc


def test_text_encoder_and_cross_attention():
    ref = import_reference()
    from gigagan_pytorch.gigagan_pytorch import CrossAttentionBlock, TextEncoder
    torch.manual_seed(0)
    te = TextEncoder(clip=_no_clip(ref), **TE)
    enc = _encodings()
    g_ref, f_ref, m_ref = te(text_encodings=enc)
    g, f, m = O.text_encoder(dict(te.state_dict()), enc, TE["depth"], TE["heads"], TE["dim_head"])
    assert torch.equal(m, m_ref)
    torch.testing.assert_close(g, g_ref, **TOL)
    torch.testing.assert_close(f, f_ref, **TOL)
    torch.manual_seed(1)
    blk = CrossAttentionBlock(16, dim_context=24, dim_head=8, heads=2)
    x = rn(6, 3, 16, 8, 8)
    torch.testing.assert_close(O.cross_attention_block(dict(blk.state_dict()), x, f_ref.detach(), m_ref, 2, 8),
                               blk(x, f_ref.detach(), m_ref), **TOL)


def test_text_conditional_generator_and_discriminator():
    ref = import_reference()
    from gigagan_pytorch.gigagan_pytorch import TextEncoder
    enc = _encodings()
    torch.manual_seed(0)
    G = ref.Generator(text_encoder=TextEncoder(clip=_no_clip(ref), **TE), **G7)
    plan = O.generator_plan(32, 2, 16, 16, 2, (16,), 2, 2, 8, unconditional=False, cross_attn_resolutions=(16, 8),
                            cross_attn_heads=2, cross_attn_dim_head=8)
    sd = dict(G.state_dict())
    z = rn(1, 3, 16)
    torch.manual_seed(2)
    rgb_ref, rgbs_ref = G(noise=z, text_encodings=enc, return_all_rgbs=True)
    gt, ft, tm = O.text_encoder(O._sub(sd, "text_encoder."), enc, TE["depth"], TE["heads"], TE["dim_head"])
    torch.manual_seed(2)
    rgb, rgbs = O.generator_forward(sd, plan, z, style_depth=2, return_all_rgbs=True, global_text_tokens=gt,
                                    fine_text_tokens=ft, text_mask=tm)
    torch.testing.assert_close(rgb, rgb_ref, **TOL)
    for a, b in zip(rgbs, rgbs_ref):
        torch.testing.assert_close(a, b, **TOL)
    torch.manual_seed(1)
    D = ref.Discriminator(text_encoder=TextEncoder(clip=_no_clip(ref), **TE), **D7)
    dplan = O.discriminator_plan(32, 2, 16, 3, (8,), (16, 8), 1, (8,), num_skip_layers_excite=2, attn_heads=2, attn_dim_head=8)
    sd = dict(D.state_dict())
    img = torch.rand(3, 3, 32, 32, generator=torch.Generator().manual_seed(4))
    with torch.no_grad():
        lo_ref, ms_ref, _ = D(img, D.real_images_to_rgbs(img), text_encodings=enc, calc_aux_loss=False)
        emb, _, _ = O.text_encoder(O._sub(sd, "text_encoder."), enc, TE["depth"], TE["heads"], TE["dim_head"])
        lo, ms, _ = O.discriminator_forward(sd, dplan, img, O.real_images_to_rgbs(img, dplan), True, False, text_embeds=emb)
    torch.testing.assert_close(lo, lo_ref, **TOL)
    for a, b in zip(ms, ms_ref):
        torch.testing.assert_close(a, b, **TOL)


# ------------------------------------------------------------------ drop-in layout: same keys, shapes and seeded values
@pytest.mark.parametrize("which", ["g64", "g256", "d64", "d256", "g_text", "d_text", "unet"])
def test_dropin_classes_reproduce_reference_state_dict(which):
    """the product classes (constructed on CPU; no kernels involved) create the reference's parameters in the
    reference's order: identical state_dict keys, shapes and - with the same manual_seed - identical initial values"""
    ref = import_reference()
    import gigagan_pytorch_b200 as g
    from gigagan_pytorch.gigagan_pytorch import TextEncoder
    readme_g = dict(dim_capacity=8, style_network=dict(dim=64, depth=4), image_size=256, dim_max=512, num_skip_layers_excite=4,
                    unconditional=True)
    readme_d = dict(dim_capacity=16, dim_max=512, image_size=256, num_skip_layers_excite=4, unconditional=True)
    mine_te = dict(TE, clip_dim_latent=32)
    make = {
        "g64": (lambda: ref.Generator(**GCFG), lambda: g.Generator(**GCFG)),
        "g256": (lambda: ref.Generator(**readme_g), lambda: g.Generator(**readme_g)),
        "d64": (lambda: ref.Discriminator(**DCFG), lambda: g.Discriminator(**DCFG)),
        "d256": (lambda: ref.Discriminator(**readme_d), lambda: g.Discriminator(**readme_d)),
        "g_text": (lambda: ref.Generator(text_encoder=TextEncoder(clip=_no_clip(ref), **TE), **G7),
                   lambda: g.Generator(text_encoder=g.TextEncoder(**mine_te), **G7)),
        "d_text": (lambda: ref.Discriminator(text_encoder=TextEncoder(clip=_no_clip(ref), **TE), **D7),
                   lambda: g.Discriminator(text_encoder=g.TextEncoder(**mine_te), **D7)),
        "unet": (lambda: ref.UnetUpsampler(dim=8, image_size=64, input_image_size=16, style_network=dict(dim=64, depth=4),
                                           unconditional=True),
                 lambda: g.UnetUpsampler(dim=8, image_size=64, input_image_size=16, style_network=dict(dim=64, depth=4),
                                         unconditional=True)),
    }[which]
    torch.manual_seed(0)
    a = make[0]().state_dict()
    torch.manual_seed(0)
    b = make[1]().state_dict()
    assert list(a.keys()) == list(b.keys())
    for k in a:
        assert a[k].shape == b[k].shape, k
        assert torch.equal(a[k], b[k]), k
