"""Oracle vs the committed golden fixtures generated from the unmodified reference
(oracle/make_golden.py).  CPU only; runs everywhere (no /root/reference needed)."""
import os

import torch

from oracle import gigagan_oracle as O

G = os.path.join(os.path.dirname(__file__), "golden")
TOL = dict(rtol=1e-4, atol=1e-5)


def load(n):
    return torch.load(os.path.join(G, n), weights_only=False)


def relerr(a, b):
    return (a - b).abs().max().item() / (b.abs().max().item() + 1e-12)


def test_ka1_adaptive_conv_fwd_bwd():
    g = load("ka1_adaptive_conv.pt")
    w, x, mod, km = (g[k].clone().requires_grad_() for k in ("weights", "x", "mod", "kernel_mod"))
    y = O.adaptive_conv2d_mod(w, x, mod, km)
    (y ** 2).sum().backward()
    torch.testing.assert_close(y, g["y"], **TOL)
    for t, k in ((w, "dweights"), (x, "dx"), (mod, "dmod"), (km, "dkernel_mod")):
        assert relerr(t.grad, g[k]) < 1e-4, k
    # SURVEY 8c KA1 scalars
    assert abs(g["y"].sum().item() - 0.154600) < 1e-4 and abs(g["y"].abs().mean().item() - 0.763597) < 1e-5


def test_ka2_attention_blocks():
    for name, dot in (("l2", False), ("dot", True)):
        g = load(f"ka2_attn_block_{name}.pt")
        sd = {k: v.clone().requires_grad_() for k, v in g["sd"].items()}
        x = g["x"].clone().requires_grad_()
        y = O.self_attention_block(sd, x, dot, heads=2, dim_head=8)
        (y ** 2).sum().backward()
        torch.testing.assert_close(y, g["y"], **TOL)
        assert relerr(x.grad, g["dx"]) < 1e-4
        for k, v in g["grads"].items():
            assert relerr(sd[k].grad, v) < 1e-4, k


def test_ka3_style_network():
    g = load("ka3_style_network.pt")
    torch.testing.assert_close(O.style_network(g["sd"], g["z"], 4), g["y"], **TOL)


def test_ka4_generator():
    g = load("ka4_generator.pt")
    c = g["cfg"]
    plan = O.generator_plan(c["image_size"], c["dim_capacity"], c["dim_max"], c["dim_latent"],
                            c["num_skip_layers_excite"], c["self_attn_resolutions"], 2, 2, 8)
    sd = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in g["sd"].items()}
    torch.manual_seed(g["noise_seed"])
    rgb, rgbs = O.generator_forward(sd, plan, g["z"], style_depth=2, return_all_rgbs=True)
    torch.testing.assert_close(rgb, g["rgb"], **TOL)
    for a, b in zip(rgbs, g["rgbs"]):
        torch.testing.assert_close(a, b, **TOL)
    (rgb ** 2).mean().backward()
    for k, v in g["grads"].items():
        assert relerr(sd[k].grad, v) < 2e-4, k


def test_ka5_discriminator_step_with_gradient_penalty():
    g = load("ka5_discriminator.pt")
    c = g["cfg"]
    plan = O.discriminator_plan(c["image_size"], c["dim_capacity"], c["dim_max"], 3, c["attn_resolutions"],
                                c["multiscale_input_resolutions"], 1, c["aux_recon_resolutions"],
                                num_skip_layers_excite=c["num_skip_layers_excite"], attn_heads=2, attn_dim_head=8)
    if True:
        sd = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in g["sd"].items()}
        with torch.no_grad():
            lo, ms, _ = O.discriminator_forward(sd, plan, g["img"], O.real_images_to_rgbs(g["img"], plan), True, False)
        torch.testing.assert_close(lo, g["logits"], **TOL)
        for a, b in zip(ms, g["ms"]):
            torch.testing.assert_close(a, b, **TOL)
        r = g["img"].clone().requires_grad_()
        f = g["fake"].clone().requires_grad_()
        frgbs = [t.detach().requires_grad_() for t in O.real_images_to_rgbs(f, plan)]
        # fixture was made with calc_aux_loss=False on both passes
        fl, fm, _ = O.discriminator_forward(sd, plan, f, frgbs, True, False)
        rl, rm, _ = O.discriminator_forward(sd, plan, r, O.real_images_to_rgbs(r, plan), True, False)
        div = O.discriminator_hinge_loss(rl, fl)
        msl = sum(O.discriminator_hinge_loss(b, a) for a, b in zip(fm, rm))
        w = [1.0] + [0.1] * len(rm)
        gp = O.gradient_penalty(r, [rl, *rm], w) + O.gradient_penalty(f, [fl, *fm], w)
        total = div + gp + 0.1 * msl
        total.backward()
    torch.testing.assert_close(total.detach(), g["loss"]["total"], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(gp.detach(), g["loss"]["gradient_penalty"], rtol=1e-4, atol=1e-6)
    for k, v in g["grads"].items():
        assert relerr(sd[k].grad, v) < 1e-3, k


def test_ka6_unet_upsampler():
    g = load("ka6_unet_upsampler.pt")
    c = g["cfg"]
    plan = O.unet_plan(c["dim"], c["image_size"], c["input_image_size"], c["dim_mults"], 3, c["full_attn"],
                       c["cross_attn"], self_attn_dim_head=8, self_attn_heads=2, cross_attn_dim_head=8,
                       attn_depths=c["attn_depths"])
    sd = {k: v.clone().requires_grad_(v.is_floating_point() and not k.endswith(("filter", ".f"))) for k, v in g["sd"].items()}
    rgb, rgbs = O.unet_forward(sd, plan, g["low"], g["z"], style_depth=2, return_all_rgbs=True)
    torch.testing.assert_close(rgb, g["rgb"], **TOL)
    for a, b in zip(rgbs, g["rgbs"]):
        torch.testing.assert_close(a, b, **TOL)
    (rgb ** 2).mean().backward()
    for k, v in g["grads"].items():
        assert relerr(sd[k].grad, v) < 2e-4, k


def test_ka7_text_conditional_generator_and_discriminator():
    """text-conditioned path: TextEncoder -> StyleNetwork concat + cross attention in G,
    text-modulated predictors in D; fixture from the unmodified reference (oracle/make_golden.py KA7)."""
    g = load("ka7_text_conditional.pt")
    c, te = g["gcfg"], g["te_cfg"]
    plan = O.generator_plan(c["image_size"], c["dim_capacity"], c["dim_max"], c["dim_latent"],
                            c["num_skip_layers_excite"], c["self_attn_resolutions"], 2, 2, 8, unconditional=False,
                            cross_attn_resolutions=c["cross_attn_resolutions"], cross_attn_heads=2,
                            cross_attn_dim_head=8)
    sd = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in g["gsd"].items()}
    gt, ft, tm = O.text_encoder(O._sub(sd, "text_encoder."), g["enc"], te["depth"], te["heads"], te["dim_head"])
    assert tm.tolist() == [[True] * 6, [True] * 4 + [False] * 2]
    torch.manual_seed(g["noise_seed"])
    rgb, rgbs = O.generator_forward(sd, plan, g["z"], style_depth=2, return_all_rgbs=True, global_text_tokens=gt,
                                    fine_text_tokens=ft, text_mask=tm)
    torch.testing.assert_close(rgb, g["rgb"], **TOL)
    for a, b in zip(rgbs, g["rgbs"]):
        torch.testing.assert_close(a, b, **TOL)
    (rgb ** 2).mean().backward()
    for k, v in g["ggrads"].items():
        assert relerr(sd[k].grad, v) < 2e-4, k
    c = g["dcfg"]
    dplan = O.discriminator_plan(c["image_size"], c["dim_capacity"], c["dim_max"], 3, c["attn_resolutions"],
                                 c["multiscale_input_resolutions"], 1, (8,),
                                 num_skip_layers_excite=c["num_skip_layers_excite"], attn_heads=2, attn_dim_head=8)
    sd = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in g["dsd"].items()}
    emb, _, _ = O.text_encoder(O._sub(sd, "text_encoder."), g["enc"], te["depth"], te["heads"], te["dim_head"])
    lo, ms, _ = O.discriminator_forward(sd, dplan, g["img"], O.real_images_to_rgbs(g["img"], dplan), True, False,
                                        text_embeds=emb)
    torch.testing.assert_close(lo, g["logits"], **TOL)
    for a, b in zip(ms, g["ms"]):
        torch.testing.assert_close(a, b, **TOL)
    (lo.sum() + sum((m ** 2).sum() for m in ms)).backward()
    for k, v in g["dgrads"].items():
        assert relerr(sd[k].grad, v) < 2e-4, k
