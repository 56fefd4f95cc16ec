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
