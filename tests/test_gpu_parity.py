"""GPU parity: the CUDA path (through the C-ABI) against the oracle / golden fixtures on identical inputs.
fp32 kernels: tolerance 1e-5 relative for forwards (1e-4 of the tensor's max for gradients, which pass through
atomics and double backward); bf16: 1e-2 of the tensor's max (north_star tolerance)."""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    import gigagan_pytorch_b200 as g
    g.set_compute_dtype(torch.float32)
    yield
    g.set_compute_dtype(torch.float32)


def dev():
    return torch.device("cuda:0")


def relmax(a, b):
    return (a.float() - b.float()).abs().max().item() / (b.float().abs().max().item() + 1e-20)


def load(n):
    return torch.load(os.path.join(GOLD, n), weights_only=False)


def rn(k, *s):
    return torch.randn(*s, generator=torch.Generator().manual_seed(k))


# ------------------------------------------------------------------ primitives
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.bfloat16, 1e-2)])
@pytest.mark.parametrize("cfg", [
    dict(n=2, h=9, w=7, ci=5, co=6, k=3, s=1, p=1), dict(n=3, h=8, w=8, ci=16, co=8, k=1, s=2, p=0),
    dict(n=2, h=8, w=8, ci=4, co=8, k=2, s=2, p=0), dict(n=2, h=12, w=12, ci=3, co=10, k=7, s=1, p=3),
    dict(n=1, h=4, w=4, ci=32, co=1, k=4, s=1, p=0), dict(n=2, h=16, w=16, ci=64, co=96, k=3, s=1, p=1),
])
def test_conv_family(cfg, dtype, tol):
    from gigagan_pytorch_b200 import ops
    c = cfg
    rd = (lambda t: t.to(dtype).float())          # the reference sees the same storage rounding as the kernel
    x = rd(rn(1, c["n"], c["ci"], c["h"], c["w"])).to(dev())
    w = rd(rn(2, c["co"], c["ci"], c["k"], c["k"]) * 0.2).to(dev()).requires_grad_()
    b = rn(3, c["co"]).to(dev()).requires_grad_()
    xr = x.clone().requires_grad_()
    ref = F.leaky_relu(F.conv2d(xr, w, b, stride=c["s"], padding=c["p"]), 0.2)
    gy = rd(torch.randn_like(ref))
    gx_ref, gw_ref, gb_ref = torch.autograd.grad(ref, (xr, w, b), gy)
    xn = x.permute(0, 2, 3, 1).contiguous().to(dtype).requires_grad_()
    y = ops.conv2d(xn, w, b, stride=c["s"], pad=c["p"], act=1)
    gx, gw, gb = torch.autograd.grad(y, (xn, w, b), gy.permute(0, 2, 3, 1).contiguous().to(dtype))
    assert relmax(y.permute(0, 3, 1, 2), ref) < tol
    assert relmax(gx.permute(0, 3, 1, 2), gx_ref) < tol * 2
    assert relmax(gw, gw_ref) < tol * 2
    assert relmax(gb, gb_ref) < tol * 2


@pytest.mark.parametrize("cfg", [
    dict(n=4, h=16, w=16, ci=64, co=64, k=3, s=1, p=1, ps=False), dict(n=2, h=32, w=32, ci=128, co=256, k=3, s=1, p=1, ps=False),
    dict(n=16, h=4, w=4, ci=512, co=512, k=3, s=1, p=1, ps=False), dict(n=3, h=8, w=8, ci=64, co=128, k=1, s=1, p=0, ps=False),
    dict(n=2, h=64, w=64, ci=16, co=16, k=3, s=1, p=1, ps=False), dict(n=2, h=32, w=32, ci=32, co=64, k=3, s=1, p=1, ps=False),
    dict(n=2, h=16, w=16, ci=64, co=128, k=2, s=2, p=0, ps=False), dict(n=2, h=16, w=16, ci=64, co=128, k=1, s=2, p=0, ps=False),
    dict(n=3, h=16, w=16, ci=64, co=32, k=3, s=1, p=1, ps=True), dict(n=2, h=8, w=8, ci=256, co=1024, k=1, s=1, p=0, ps=False),
    dict(n=75, h=32, w=32, ci=256, co=256, k=3, s=1, p=1, ps=False),      # dual-M tile mode (600 pixel tiles)
    dict(n=149, h=16, w=16, ci=256, co=512, k=3, s=1, p=1, ps=False),     # dual-M, odd tile count (298 tiles x 2)
    # thin high-resolution layers (conv_thin_tc.cu: rows staged once in a shared-memory ring, resident filters)
    dict(n=2, h=128, w=128, ci=16, co=32, k=3, s=1, p=1, ps=False), dict(n=1, h=256, w=256, ci=32, co=16, k=3, s=1, p=1, ps=False),
    dict(n=3, h=128, w=128, ci=64, co=64, k=3, s=1, p=1, ps=False), dict(n=5, h=128, w=128, ci=32, co=16, k=3, s=1, p=1, ps=True),
    dict(n=3, h=40, w=256, ci=16, co=16, k=1, s=1, p=0, ps=False), dict(n=7, h=24, w=128, ci=64, co=32, k=3, s=1, p=1, ps=True),
    dict(n=3, h=128, w=128, ci=48, co=48, k=3, s=1, p=1, ps=False),
])
def test_tcgen05_conv_matches_ffma(cfg):
    """bf16 tensor-core implicit GEMM (TMA taps, TMEM accumulators) vs the FFMA kernel on identical bf16 inputs;
    both accumulate in fp32, so they agree to accumulation-order noise (bit-exactness is not defined for fp)."""
    from gigagan_pytorch_b200 import ops, _lib
    c = cfg
    dt = torch.bfloat16
    x = rn(1, c["n"], c["h"], c["w"], c["ci"]).to(dev()).to(dt)
    wshape = ((c["n"],) if c["ps"] else ()) + (c["co"], c["k"], c["k"], c["ci"])
    w = (rn(2, *wshape) * (c["ci"] * c["k"] * c["k"]) ** -0.5).to(dev()).to(dt)
    bias = rn(3, c["co"]).to(dev())
    g = ops.ConvGeom(c["k"], c["k"], c["s"], c["p"], c["ps"], act=1)
    L = _lib.lib()
    y_tc = ops._conv_fprop_raw(x, w, bias, None, g, c["co"])
    old = L.gg_set_flags(1)
    try:
        y_ff = ops._conv_fprop_raw(x, w, bias, None, g, c["co"])
    finally:
        L.gg_set_flags(old)
    torch.cuda.synchronize()
    assert relmax(y_tc, y_ff) < 2e-2 and (y_tc.float() - y_ff.float()).abs().mean().item() < 2e-3 * y_ff.float().abs().mean().item()
    # residual + gain epilogue
    res = rn(4, *y_ff.shape).to(dev()).to(dt)
    g2 = ops.ConvGeom(c["k"], c["k"], c["s"], c["p"], c["ps"], act=0, gain=2 ** -0.5)
    a = ops._conv_fprop_raw(x, w, bias, res, g2, c["co"])
    old = L.gg_set_flags(1)
    try:
        b = ops._conv_fprop_raw(x, w, bias, res, g2, c["co"])
    finally:
        L.gg_set_flags(old)
    assert relmax(a, b) < 2e-2
    # weight gradient (MN-major tcgen05 operands, split over pixel ranges, fp32 red.add)
    gy = rn(5, *y_ff.shape).to(dev()).to(dt)
    gp = ops.ConvGeom(c["k"], c["k"], c["s"], c["p"], c["ps"])
    dw_tc = ops._conv_wgrad_raw(x, gy, gp)
    old = L.gg_set_flags(1)
    try:
        dw_ff = ops._conv_wgrad_raw(x, gy, gp)
    finally:
        L.gg_set_flags(old)
    assert relmax(dw_tc, dw_ff) < 1e-3, relmax(dw_tc, dw_ff)


def test_conv_double_backward_fp32():
    """gradient-penalty shaped check: d/dw of |d y / d x|^2 through Conv2dFn/ConvDgradFn/ConvWgradFn."""
    from gigagan_pytorch_b200 import ops
    x = rn(1, 2, 4, 6, 6).to(dev())
    w1 = (rn(2, 8, 4, 3, 3) * 0.3).to(dev()).requires_grad_()
    w2 = (rn(3, 1, 8, 3, 3) * 0.3).to(dev()).requires_grad_()

    def penalty(conv, xin, to_out):
        h = conv(xin, w1, 1, True)
        o = conv(h, w2, 1, False)
        g, = torch.autograd.grad(o.sum(), xin, create_graph=True)
        return (g.float() ** 2).sum()

    xr = x.clone().requires_grad_()
    pr = penalty(lambda t, w, p, a: F.leaky_relu(F.conv2d(t, w, padding=p), 0.2) if a else F.conv2d(t, w, padding=p), xr, None)
    gr = torch.autograd.grad(pr, (w1, w2))
    xn = x.permute(0, 2, 3, 1).contiguous().requires_grad_()
    h = ops.conv2d(xn, w1, pad=1, act=1)
    o = ops.conv2d(h, w2, pad=1)
    g, = torch.autograd.grad(ops.sum_all(o), xn, create_graph=True)
    pm = ops.sum_all(ops.mul(g, g))
    gm = torch.autograd.grad(pm, (w1, w2))
    assert relmax(pm, pr) < 1e-4
    for a, b in zip(gm, gr):
        assert relmax(a, b) < 1e-4


def test_bmm_and_linear():
    from gigagan_pytorch_b200 import ops
    a = rn(1, 2, 3, 17, 9).to(dev()).requires_grad_()
    b = rn(2, 2, 3, 9, 21).to(dev()).requires_grad_()
    c = ops.bmm(a, b, alpha=0.5)
    ref = 0.5 * (a @ b)
    assert relmax(c, ref) < 1e-5
    g = torch.randn_like(ref)
    for m, r in zip(torch.autograd.grad(c, (a, b), g), torch.autograd.grad(ref, (a, b), g)):
        assert relmax(m, r) < 1e-5
    at = rn(3, 2, 9, 3, 17).to(dev()).permute(0, 2, 3, 1)          # strided view
    assert relmax(ops.bmm(at, b, out_bmhn=True), at @ b) < 1e-5
    x, w, bias = rn(4, 5, 7).to(dev()), rn(5, 11, 7).to(dev()), rn(6, 11).to(dev())
    assert relmax(ops.linear(x, w, bias), F.linear(x, w, bias)) < 1e-5


@pytest.mark.parametrize("kind", ["lrelu", "relu", "gelu", "silu", "sigmoid", "invnorm"])
def test_unary_levels(kind):
    from gigagan_pytorch_b200 import ops
    fns = dict(lrelu=(ops.U_LRELU, lambda t: F.leaky_relu(t, 0.2)), relu=(ops.U_RELU, F.relu),
               gelu=(ops.U_GELU, F.gelu), silu=(ops.U_SILU, F.silu), sigmoid=(ops.U_SIGMOID, torch.sigmoid),
               invnorm=(ops.U_INVNORM, lambda t: 1.0 / t.sqrt().clamp(min=1e-12)))
    k, f = fns[kind]
    x = rn(1, 64, 33).to(dev())
    if kind == "invnorm":
        x = x.abs() + 0.1
    x1, x2 = x.clone().requires_grad_(), x.clone().requires_grad_()
    y1, y2 = ops.unary(k, x1), f(x2)
    assert relmax(y1, y2) < 1e-5
    g1, = torch.autograd.grad(y1.sum() if False else ops.sum_all(ops.mul(y1, y1)), x1, create_graph=True)
    g2, = torch.autograd.grad((y2 * y2).sum(), x2, create_graph=True)
    assert relmax(g1, g2) < 1e-4
    h1, = torch.autograd.grad(ops.sum_all(ops.mul(g1, g1)), x1)
    h2, = torch.autograd.grad((g2 * g2).sum(), x2)
    assert relmax(h1, h2) < 1e-3


def test_broadcasts_reductions_softmax():
    from gigagan_pytorch_b200 import ops
    x = rn(1, 6, 5, 5, 12).to(dev()).requires_grad_()             # (N,H,W,C): 2 samples x 3 scales
    s = rn(2, 2, 12).to(dev()).requires_grad_()
    y = ops.scale_channels(x, s, 25, 2)
    ref = x * s.repeat(3, 1)[:, None, None, :]
    assert relmax(y, ref) < 1e-6
    g = torch.randn_like(ref)
    for m, r in zip(torch.autograd.grad(y, (x, s), g), torch.autograd.grad(ref, (x, s), g)):
        assert relmax(m, r) < 1e-5
    assert relmax(ops.mean_hw(x), x.mean(dim=(1, 2))) < 1e-5
    assert relmax(ops.rowdot(x, x), (x * x).sum(-1)) < 1e-5
    sm = rn(3, 7, 4, 33).to(dev()).requires_grad_()
    p = ops.softmax(sm)
    pr = sm.softmax(-1)
    assert relmax(p, pr) < 1e-5
    gp = torch.randn_like(pr)
    assert relmax(torch.autograd.grad(p, sm, gp)[0], torch.autograd.grad(pr, sm, gp)[0]) < 1e-5


def test_resample_and_layout():
    from gigagan_pytorch_b200 import ops
    from oracle import gigagan_oracle as O
    x = rn(1, 2, 5, 8, 8).to(dev()).requires_grad_()
    xn = ops.to_nhwc(x, 5, torch.float32)
    y = ops.to_nchw(ops.upsample2x_blur(xn), 5)
    ref = O.upsample2x(x)
    assert relmax(y, ref) < 1e-5
    g = torch.randn_like(ref)
    assert relmax(torch.autograd.grad(y, x, g)[0], torch.autograd.grad(ref, x, g)[0]) < 1e-5
    img = torch.rand(2, 3, 32, 32, device=dev())
    r = ops.to_nchw(ops.resize_bilinear(ops.to_nhwc(img, 3, torch.float32), 8), 3)
    assert relmax(r, F.interpolate(img, 8, mode="bilinear")) < 1e-5


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-4), (torch.bfloat16, 2e-2)])
def test_ka1_adaptive_conv(dtype, tol):
    import gigagan_pytorch_b200 as g
    fx = load("ka1_adaptive_conv.pt")
    g.set_compute_dtype(dtype)
    m = g.AdaptiveConv2DMod(8, 6, 3, num_conv_kernels=2).to(dev())
    with torch.no_grad():
        m.weights.copy_(fx["weights"])
    x, mod, km = (fx[k].to(dev()).requires_grad_() for k in ("x", "mod", "kernel_mod"))
    y = m(x, mod=mod, kernel_mod=km)
    (y ** 2).sum().backward()
    assert relmax(y, fx["y"].to(dev())) < tol
    for t, k in ((m.weights, "dweights"), (x, "dx"), (mod, "dmod"), (km, "dkernel_mod")):
        assert relmax(t.grad, fx[k].to(dev())) < tol * 3, k


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("name,dot", [("l2", False), ("dot", True)])
def test_ka2_attention_block(name, dot, fused):
    import gigagan_pytorch_b200 as g
    from gigagan_pytorch_b200 import ops
    fx = load(f"ka2_attn_block_{name}.pt")
    blk = g.SelfAttentionBlock(16, dim_head=8, heads=2, dot_product=dot).to(dev())
    blk.load_state_dict(fx["sd"])
    x = fx["x"].to(dev()).requires_grad_()
    xn = ops.to_nhwc(x, 16, torch.float32)
    y = ops.to_nchw(blk.forward_nhwc(xn, fused=fused), 16)
    (y ** 2).sum().backward()
    assert relmax(y, fx["y"].to(dev())) < 1e-4
    assert relmax(x.grad, fx["dx"].to(dev())) < 2e-4
    for k, v in fx["grads"].items():
        assert relmax(dict(blk.named_parameters())[k].grad, v.to(dev())) < 2e-4, k


def test_ka3_style_network():
    import gigagan_pytorch_b200 as g
    fx = load("ka3_style_network.pt")
    sn = g.StyleNetwork(dim=64, depth=4).to(dev())
    sn.load_state_dict(fx["sd"])
    assert relmax(sn(fx["z"].to(dev())), fx["y"].to(dev())) < 1e-5


# ------------------------------------------------------------------ whole models vs fixtures from the reference
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-4), (torch.bfloat16, 4e-2)])
def test_ka4_generator(dtype, tol):
    import gigagan_pytorch_b200 as g
    fx = load("ka4_generator.pt")
    g.set_compute_dtype(dtype)
    G = g.Generator(**fx["cfg"]).to(dev())
    G.load_state_dict(fx["sd"])
    # the fixture drew its layer noises from CPU randn under manual_seed(noise_seed); reproduce them exactly
    torch.manual_seed(fx["noise_seed"])
    res = [4, 8, 16, 32]
    noises = []
    for r in res:
        for _ in range(2):
            noises.append(torch.randn(2, 1, r, r).to(dev()))
    rgb, rgbs = G.forward_nhwc(noise=fx["z"].to(dev()), layer_noises=noises)
    from gigagan_pytorch_b200 import ops
    out = ops.to_nchw(rgb, 3)
    assert relmax(out, fx["rgb"].to(dev())) < tol
    for a, b in zip(rgbs, fx["rgbs"]):
        assert relmax(ops.to_nchw(a, 3), b.to(dev())) < tol
    (out ** 2).mean().backward()
    named = dict(G.named_parameters())
    worst = max((relmax(named[k].grad, v.to(dev())), k) for k, v in fx["grads"].items())
    assert worst[0] < tol * 5, worst


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-4), (torch.bfloat16, 4e-2)])
def test_ka5_discriminator_step_with_gradient_penalty(dtype, tol):
    import gigagan_pytorch_b200 as g
    from gigagan_pytorch_b200 import ops
    from gigagan_pytorch_b200.trainer import discriminator_hinge_loss, gradient_penalty
    fx = load("ka5_discriminator.pt")
    g.set_compute_dtype(dtype)
    D = g.Discriminator(**fx["cfg"]).to(dev())
    D.load_state_dict(fx["sd"])
    D.eval()
    with torch.no_grad():
        lo, ms, _ = D(fx["img"].to(dev()), D.real_images_to_rgbs(fx["img"].to(dev())), calc_aux_loss=False)
    assert relmax(lo, fx["logits"].to(dev())) < tol
    for a, b in zip(ms, fx["ms"]):
        assert relmax(a, b.to(dev())) < tol
    D.train()
    r = fx["img"].to(dev()).requires_grad_()
    f = fx["fake"].to(dev()).requires_grad_()
    frgbs = [t.detach().requires_grad_() for t in D.real_images_to_rgbs(f)]
    # gradient penalty needs the composed (any-order differentiable) attention
    dt = dtype
    fn = ops.to_nhwc(f, 3, dt)
    rn_ = ops.to_nhwc(r, 3, dt)
    fl, fm, _ = D.forward_nhwc(fn, [ops.to_nhwc(t, 3, dt) for t in frgbs], True, False, fused_attention=False)
    rl, rm, _ = D.forward_nhwc(rn_, D.real_images_to_rgbs_nhwc(rn_), True, False, fused_attention=False)
    div = discriminator_hinge_loss(rl, fl)
    msl = sum(discriminator_hinge_loss(b, a) for a, b in zip(fm, rm))
    w = [1.0] + [0.1] * len(rm)
    gp = gradient_penalty(r, [rl, *rm], w) + gradient_penalty(f, [fl, *fm], w)
    total = div + gp + 0.1 * msl
    total.backward()
    assert relmax(gp, fx["loss"]["gradient_penalty"].to(dev())) < tol * 5
    assert relmax(total, fx["loss"]["total"].to(dev())) < tol * 5
    named = dict(D.named_parameters())
    worst = max((relmax(named[k].grad, v.to(dev())), k) for k, v in fx["grads"].items())
    assert worst[0] < tol * 10, worst


@pytest.mark.parametrize("case", ["kk", "kmn", "mnk", "mnmn"])
def test_tcgen05_bmm_matches_ffma(case):
    """all four operand-majorness combinations of the tensor-core batched GEMM, on strided attention-shaped views"""
    from gigagan_pytorch_b200 import ops, _lib
    n, heads, seq, d, Lp = 3, 2, 256, 64, 320
    dt = torch.bfloat16
    q = rn(1, n, seq, heads, d).to(dev()).to(dt)
    kf = rn(2, n, Lp, heads, d).to(dev()).to(dt)
    pm = rn(3, n, heads, seq, Lp).to(dev()).to(dt)
    if case == "kk":      # S = Q K^T
        a, b = q.permute(0, 2, 1, 3), kf.permute(0, 2, 3, 1)
    elif case == "kmn":   # O = P V
        a, b = pm, kf.permute(0, 2, 1, 3)
    elif case == "mnk":   # dKt = Q^T dS  -> A MN-major (M = d contiguous), B MN-major (keys contiguous)
        a, b = q.permute(0, 2, 3, 1), pm
    else:                 # dV = P^T dO
        a, b = pm.transpose(-1, -2), q.permute(0, 2, 1, 3)
    L = _lib.lib()
    c_tc = ops.bmm(a, b, alpha=0.5)
    old = L.gg_set_flags(1)
    try:
        c_ff = ops.bmm(a, b, alpha=0.5)
    finally:
        L.gg_set_flags(old)
    ref = 0.5 * (a.float() @ b.float())
    assert relmax(c_ff, ref) < 1e-2
    assert relmax(c_tc, ref) < 1e-2, relmax(c_tc, ref)
    o = ops.bmm(a, b, out_bmhn=True)
    assert relmax(o, a.float() @ b.float()) < 1e-2


def test_composed_attention_padded_bf16_vs_fp32():
    import gigagan_pytorch_b200 as g
    torch.manual_seed(0)
    for dot in (False, True):
        blk = g.SelfAttentionBlock(64, dim_head=64, heads=2, dot_product=dot).to(dev())
        x = torch.randn(2, 16, 16, 64, device=dev())
        ref = blk.forward_nhwc(x, fused=False)
        out = blk.forward_nhwc(x.to(torch.bfloat16), fused=False)
        assert relmax(out, ref) < 3e-2
        fz = blk.forward_nhwc(x.to(torch.bfloat16), fused=True)
        assert relmax(fz, ref) < 3e-2


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-4), (torch.bfloat16, 2e-2)])
@pytest.mark.parametrize("dot", [False, True])
def test_attention_node_matches_primitive_composition_through_double_backward(dot, dtype, tol):
    """gradient-penalty attention as ONE any-order node (ops.ComposedAttnFn: K-concatenated products, deferred rank-d
    gradient of the probabilities, softmax backward with on-the-fly addend) against the composition from closed
    primitives, on a penalty-shaped objective (first-order term + squared input gradient) at 32x32 tokens, both logit
    forms (gigagan_pytorch.py:562-592 under :138-155)."""
    import gigagan_pytorch_b200 as g
    from gigagan_pytorch_b200 import modules
    g.set_compute_dtype(dtype)
    torch.manual_seed(0)
    blk = g.SelfAttention(64, dim_head=64, heads=2, dot_product=dot).to(dev())
    x0 = torch.randn(2, 32, 32, 64, device=dev())
    w = torch.randn(2, 32, 32, 64, device=dev()).to(dtype)
    res = []
    for node in (False, True, "concat"):                      # "concat": the node fed by concatenations (no operand-builder kernel)
        old = modules._COMPUTE["attention_node"], modules._COMPUTE["attention_augment"]
        modules._COMPUTE["attention_node"], modules._COMPUTE["attention_augment"] = bool(node), node is True
        try:
            blk.zero_grad(set_to_none=True)
            x = x0.clone().to(dtype).requires_grad_()
            o = blk.forward_nhwc(x, fused=False)
            gx, = torch.autograd.grad(ops_sum(o, w), x, create_graph=True, retain_graph=True)
            total = ops_sum(o, w) + 10.0 * ops_sum(gx, gx)
            total.backward()
            res.append([o.detach(), gx.detach(), x.grad] + [p.grad for p in blk.parameters()])
        finally:
            modules._COMPUTE["attention_node"], modules._COMPUTE["attention_augment"] = old
    for other in res[1:]:
        for a, b in zip(other, res[0]):
            assert torch.isfinite(a).all()
            assert relmax(a, b) < tol, (a.shape, relmax(a, b))


def ops_sum(a, b):
    from gigagan_pytorch_b200 import ops
    a2, b2 = a.reshape(-1, a.shape[-1]), b.reshape(-1, b.shape[-1])
    return ops.sum_all(ops.dot_sc(a2, b2, a2.shape[0], 1))


def test_attn_augment_kernels_match_formulas():
    """csrc/attn_augment.cu (operand builder of the gradient-penalty attention node, its first and second derivative)
    against the plain formulas on identical bf16 inputs: layouts, the hi/lo split of -|k|^2/2, the padding rows, the null
    key/value rows and gradients."""
    from gigagan_pytorch_b200 import ops
    torch.manual_seed(0)
    n, seq, h, d, Lp = 3, 100, 2, 64, 128
    bf = torch.bfloat16
    q4 = (torch.randn(n, seq, h, d, device=dev()) * 0.5).to(bf)
    v4 = torch.randn(n, seq, h, d, device=dev()).to(bf)
    nk = torch.randn(2, h, d, device=dev())
    nkr = nk.to(bf).float()                                   # the key / value rows hold the bf16-rounded parameter
    qa, ka, vf = ops._k_aug_fwd(q4, v4, nk, Lp)
    kf = torch.cat([nkr[0][None, None].expand(n, 1, h, d), q4.float(), torch.zeros(n, Lp - seq - 1, h, d, device=dev())], 1)
    t = -0.5 * (kf * kf).sum(-1)
    assert torch.equal(qa[..., :64], q4) and torch.equal(ka[:, 1:seq + 1, :, :64], q4)
    assert (qa[..., 64:66].float() == 1).all() and (qa[..., 66:] == 0).all() and (ka[..., 66:] == 0).all()
    assert torch.equal(ka[:, 0, :, :64].float(), nkr[0][None].expand(n, h, d)) and (ka[:, seq + 1:, :, :64] == 0).all()
    hl = ka[..., 64].float() + ka[..., 65].float()
    assert ((hl[:, :seq + 1] - t[:, :seq + 1]).abs() <= 2.0 ** -15 * t[:, :seq + 1].abs() + 1e-6).all()
    assert (ka[:, seq + 1:, :, 64].float() < -1e29).all()
    assert torch.equal(vf[:, 1:seq + 1], v4) and (vf[:, seq + 1:] == 0).all()
    assert torch.equal(vf[:, 0].float(), nkr[1][None].expand(n, h, d))
    # first derivative
    dqa = torch.randn(n, seq, h, 80, device=dev()).to(bf)
    dka = torch.randn(n, Lp, h, 80, device=dev()).to(bf)
    dvf = torch.randn(n, Lp, h, d, device=dev()).to(bf)
    dq, dv, dnull = ops._k_aug_bwd(dqa, dka, dvf, q4, nk)
    ghi = dka[:, 1:seq + 1, :, 64:65].float()
    ref_dq = dqa[..., :64].float() + dka[:, 1:seq + 1, :, :64].float() - ghi * q4.float()
    assert relmax(dq, ref_dq) < 1e-2 and torch.equal(dv, dvf[:, 1:seq + 1])
    ref_dn = torch.stack([(dka[:, 0, :, :64].float() - dka[:, 0, :, 64:65].float() * nkr[0][None]).sum(0), dvf[:, 0].float().sum(0)])
    assert relmax(dnull, ref_dn) < 1e-5
    # second derivative
    wq = torch.randn(n, seq, h, d, device=dev()).to(bf)
    wv = torch.randn(n, seq, h, d, device=dev()).to(bf)
    wn = torch.randn(2, h, d, device=dev())
    g_dqa, g_dka, g_dvf, g_q, g_null = ops._k_aug_bwd2(wq, wv, wn, q4, nk, dka)
    assert torch.equal(g_dqa[..., :64], wq) and (g_dqa[..., 64:] == 0).all()
    assert torch.equal(g_dka[:, 1:seq + 1, :, :64], wq) and (g_dka[:, seq + 1:] == 0).all() and (g_dka[..., 65:] == 0).all()
    assert relmax(g_dka[:, 1:seq + 1, :, 64], -(wq.float() * q4.float()).sum(-1)) < 1e-2
    assert relmax(g_dka[:, 0, :, :64], wn[0][None].expand(n, h, d)) < 1e-2
    assert relmax(g_dka[:, 0, :, 64], -(wn[0] * nkr[0]).sum(-1)[None].expand(n, h)) < 1e-2
    assert torch.equal(g_dvf[:, 1:seq + 1], wv) and (g_dvf[:, seq + 1:] == 0).all()
    assert relmax(g_dvf[:, 0], wn[1][None].expand(n, h, d)) < 1e-2
    assert relmax(g_q, -ghi * wq.float()) < 1e-2
    ref_gn = torch.stack([-(dka[:, 0, :, 64:65].float() * wn[0][None]).sum(0), torch.zeros(h, d, device=dev())])
    assert relmax(g_null, ref_gn) < 1e-5


def test_fused_attention_matches_composed_large():
    """size beyond the oracle's reach: fused (online softmax) vs composed (materialised) on 32x32 tokens."""
    import gigagan_pytorch_b200 as g
    from gigagan_pytorch_b200 import ops
    torch.manual_seed(0)
    blk = g.SelfAttentionBlock(64, dim_head=64, heads=2, dot_product=False).to(dev())
    x = torch.randn(2, 32, 32, 64, device=dev())
    a = blk.forward_nhwc(x, fused=True)
    b = blk.forward_nhwc(x, fused=False)
    assert relmax(a, b) < 1e-4


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("nk,hw,ci,co", [(2, 8, 64, 32), (2, 4, 128, 128), (1, 8, 64, 48), (3, 4, 512, 512), (2, 16, 128, 64)])
def test_adaptive_conv_shared_bank_identity_matches_per_sample(fused, nk, hw, ci, co):
    """low-resolution AdaptiveConv2DMod: shared-bank dense formulation (bf16, tcgen05) - the any-order composition and the
    single fused autograd node (ops.SharedBankConvFn) - vs the reference algorithm (oracle, fp32), forward and every
    gradient.  mod / kernel_mod are column slices of one wider tensor, as the generator passes them."""
    import gigagan_pytorch_b200 as g
    from oracle import gigagan_oracle as O
    torch.manual_seed(0)
    B = 4
    m = g.AdaptiveConv2DMod(ci, co, 3, num_conv_kernels=nk).to(dev())
    x = rn(1, B, ci, hw, hw).to(dev())
    wide = torch.cat(((rn(2, B, ci) * 0.5), rn(3, B, max(nk, 1)), rn(4, B, 5)), dim=1).to(dev())
    mod, km = wide[:, :ci], (wide[:, ci:ci + nk] if nk > 1 else None)
    xr, wr, mr = (t.detach().clone().requires_grad_() for t in (x, m.weights, mod))
    kr = km.detach().clone().requires_grad_() if nk > 1 else None
    ref = O.adaptive_conv2d_mod(wr, xr, mr, kr)
    gy = torch.randn_like(ref)
    ins_r = (xr, wr, mr) + ((kr,) if nk > 1 else ())
    gref = torch.autograd.grad(ref, ins_r, gy)
    g.set_compute_dtype(torch.bfloat16)
    xn = x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).requires_grad_()
    wide2 = wide.clone().requires_grad_()
    mod2, km2 = wide2[:, :ci], (wide2[:, ci:ci + nk] if nk > 1 else None)
    y = m.forward_nhwc(xn, mod2, km2, fused=fused)
    assert relmax(y.permute(0, 3, 1, 2), ref) < 2e-2
    gm = torch.autograd.grad(y, (xn, m.weights, wide2), gy.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16))
    assert relmax(gm[0].permute(0, 3, 1, 2), gref[0]) < 3e-2
    assert relmax(gm[1], gref[1]) < 3e-2
    assert relmax(gm[2][:, :ci], gref[2]) < 3e-2
    if nk > 1:
        assert relmax(gm[2][:, ci:ci + nk], gref[3]) < 3e-2
    assert gm[2][:, ci + max(nk, 1):].abs().max().item() == 0


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 1e-2)])
@pytest.mark.parametrize("r,k", [(300, 64), (4096, 256), (1000, 512)])
def test_row_linear_head_matches_torch(dtype, tol, r, k):
    """one-output-channel logit heads (ref gigagan_pytorch.py:1470, :1497): the bandwidth kernel pair against torch fp32 on
    the same (storage-rounded) activations: y, dx, dw, dbias"""
    from gigagan_pytorch_b200 import ops
    x = rn(1, r, k).to(dtype).to(dev()).requires_grad_()
    w = (rn(2, 1, k) * k ** -0.5).to(dev()).requires_grad_()
    b = rn(3, 1).to(dev()).requires_grad_()
    y = ops.linear_rows(x, w, b, fused=True)
    assert y.dtype == torch.float32 and y.shape == (r, 1)
    xr = x.detach().float().requires_grad_()
    yr = xr @ w.t() + b
    assert relmax(y, yr) < tol
    gy = torch.randn_like(yr)
    g = torch.autograd.grad(y, (x, w, b), gy)
    gr = torch.autograd.grad(yr, (xr, w, b), gy)
    for a_, b_ in zip(g, gr):
        assert relmax(a_, b_) < tol


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_patch_select_matches_rearrange(dtype):
    """the aux decoder's patch subset (ref gigagan_pytorch.py:1300-1312: 'b c (p1 h) (p2 w) -> b (p1 p2) c h w', per-image
    index selection, '(b p)') as one gather launch, and its adjoint"""
    from gigagan_pytorch_b200 import ops
    B, pd, hh, ww, C, nsel = 3, 2, 4, 5, 8, 2
    t = rn(1, B, C, pd * hh, pd * ww)
    perm = torch.stack([torch.randperm(pd * pd, generator=torch.Generator().manual_seed(5 + i))[:nsel] for i in range(B)])
    patches = t.view(B, C, pd, hh, pd, ww).permute(0, 2, 4, 1, 3, 5).reshape(B, pd * pd, C, hh, ww)
    ref = patches[torch.arange(B)[:, None], perm].reshape(B * nsel, C, hh, ww)                  # the reference's indexing
    tn = t.permute(0, 2, 3, 1).contiguous().to(dtype).to(dev()).requires_grad_()
    sel = perm.to(torch.int32).contiguous().to(dev())
    out = ops.patch_select(tn, sel, pd)
    assert torch.equal(out.detach().float().cpu().permute(0, 3, 1, 2), ref.to(dtype).float())
    gy = torch.randn_like(out)
    gt, = torch.autograd.grad(out, tn, gy)
    tr = t.to(dtype).float().requires_grad_()
    pr = tr.view(B, C, pd, hh, pd, ww).permute(0, 2, 4, 1, 3, 5).reshape(B, pd * pd, C, hh, ww)
    gr, = torch.autograd.grad(pr[torch.arange(B)[:, None], perm].reshape(B * nsel, C, hh, ww), tr,
                              gy.float().cpu().permute(0, 3, 1, 2))
    assert torch.equal(gt.float().cpu().permute(0, 3, 1, 2), gr)


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("n", [256, 1024])
@pytest.mark.parametrize("variant", [0, 8, 16, 32, 64, 128])
def test_tcgen05_fused_attention(mode, n, variant):
    """tcgen05 fused attention forward + backward (TMEM accumulators, softmax out of TMEM, P/dS through swizzled
    smem) vs the FFMA flash kernels on identical bf16 inputs (strided q/k/v views, null key/value, dot and L2).
    variant = gg_set_flags bits: 0 default (second generation; single-pass forward for the shared-QK L2 form),
    8 first generation, 16 / 32 second generation with 8 / 16 softmax warps everywhere, 64 two-pass L2 forward,
    128 forward with one CTA per SM (the default runs two)."""
    from gigagan_pytorch_b200 import _lib, ops
    B, heads, d = 3, 2, 64
    dt = torch.bfloat16
    L = _lib.lib()
    null_kv = rn(2, 2, heads, d).to(dev())
    go = (rn(3, B, n, heads * d)).to(dev()).to(dt)
    res = []
    for force_ffma in (0, 1):
        qkv = (rn(1, B, n, 3 * heads * d) * 0.7).to(dev()).to(dt).requires_grad_()
        nk = null_kv.clone().requires_grad_()
        q, k, v = qkv[..., : heads * d], qkv[..., heads * d: 2 * heads * d], qkv[..., 2 * heads * d:]
        old = L.gg_set_flags(1 if force_ffma else variant)
        try:
            o = ops.fused_attention(q, q if mode == 1 else k, v, nk, heads, d ** -0.5, l2=(mode == 1))
            g1, g2 = torch.autograd.grad(o, (qkv, nk), go)
        finally:
            L.gg_set_flags(old)
        res.append((o.detach(), g1, g2))
    torch.cuda.synchronize()
    (o_tc, gq_tc, gn_tc), (o_ff, gq_ff, gn_ff) = res
    assert relmax(o_tc, o_ff) < 2e-2, relmax(o_tc, o_ff)
    assert relmax(gq_tc, gq_ff) < 3e-2, relmax(gq_tc, gq_ff)
    assert relmax(gn_tc, gn_ff) < 3e-2, relmax(gn_tc, gn_ff)


def test_fused_adamw_matches_torch_adamw():
    """gg_adamw over the flat buffer == torch.optim.AdamW as the reference configures it (wd 1e-2 on ndim>=2 only)."""
    from gigagan_pytorch_b200.trainer import FlatAdamW
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3), torch.nn.Linear(8, 5)).to(dev())
    ref = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3), torch.nn.Linear(8, 5)).to(dev())
    ref.load_state_dict(net.state_dict())
    opt = FlatAdamW(net, lr=2e-4, betas=(0.5, 0.9))
    wd, nwd = [p for p in ref.parameters() if p.ndim >= 2], [p for p in ref.parameters() if p.ndim < 2]
    ropt = torch.optim.AdamW([dict(params=wd), dict(params=nwd, weight_decay=0.0)], lr=2e-4, betas=(0.5, 0.9), weight_decay=1e-2)
    for step in range(5):
        opt.zero_grad()
        for p, q in zip(net.parameters(), ref.parameters()):
            g = torch.randn_like(p)
            p.grad.copy_(g)
            q.grad = g.clone()
        opt.step()
        ropt.step()
    for p, q in zip(net.parameters(), ref.parameters()):
        assert (p - q).abs().max().item() < 1e-6, (p - q).abs().max().item()


@pytest.mark.parametrize("dtype,tol,n", [(torch.float32, 1e-4, 96), (torch.bfloat16, 2e-2, 256)])
def test_attend_matches_reference_formula(dtype, tol, n):
    """attend.py:99-108: softmax(q k^T * d^-0.5) v, forward and gradients."""
    import gigagan_pytorch_b200 as g
    g.set_compute_dtype(dtype)
    att = g.Attend()
    q, k, v = (rn(s, 2, 4, n, 64).to(dev()).requires_grad_() for s in (1, 2, 3))
    out = att(q, k, v)
    ref = torch.softmax((q @ k.transpose(-1, -2)) * 64 ** -0.5, dim=-1) @ v
    go = torch.randn_like(ref)
    assert relmax(out, ref) < tol
    for a, b in zip(torch.autograd.grad(out, (q, k, v), go), torch.autograd.grad(ref, (q, k, v), go)):
        assert relmax(a, b) < tol * 3


# bf16 is compared with the reference's FP32 output here (the fixture), hence the looser bound for this deep network
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-4), (torch.bfloat16, 8e-2)])
def test_ka6_unet_upsampler(dtype, tol):
    import gigagan_pytorch_b200 as g
    fx = load("ka6_unet_upsampler.pt")
    g.set_compute_dtype(dtype)
    U = g.UnetUpsampler(**fx["cfg"]).to(dev())
    U.load_state_dict(fx["sd"])
    rgb, rgbs = U(fx["low"].to(dev()), noise=fx["z"].to(dev()), return_all_rgbs=True)
    assert relmax(rgb, fx["rgb"].to(dev())) < tol
    for a, b in zip(rgbs, fx["rgbs"]):
        assert relmax(a, b.to(dev())) < tol
    (rgb ** 2).mean().backward()
    named = dict(U.named_parameters())
    worst = max((relmax(named[k].grad, v.to(dev())), k) for k, v in fx["grads"].items())
    assert worst[0] < tol * 6, worst


def test_maxpool_and_token_softmax():
    from gigagan_pytorch_b200 import ops
    x = rn(1, 2, 8, 6, 5).to(dev()).requires_grad_()
    y = ops.maxpool2(x)
    ref = F.max_pool2d(x.permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1)
    assert relmax(y, ref) < 1e-6
    g = torch.randn_like(ref)
    assert relmax(torch.autograd.grad(y, x, g)[0], torch.autograd.grad(ref, x, g)[0]) < 1e-6
    t = rn(2, 3, 50, 7).to(dev()).requires_grad_()
    p, pr = ops.softmax_tokens(t), t.softmax(dim=1)
    assert relmax(p, pr) < 1e-5
    gp = torch.randn_like(pr)
    assert relmax(torch.autograd.grad(p, t, gp)[0], torch.autograd.grad(pr, t, gp)[0]) < 1e-5


@pytest.mark.parametrize("upsampler", [False, True])
@pytest.mark.parametrize("amp", [False, True])
def test_trainer_steps_run_and_cuda_graph_replay_matches_eager(upsampler, amp):
    """Eager vs CUDA-graph trainer in LOCKSTEP on a tiny config (plain generator and train_upsampler=True): before every
    step the graph trainer is given the eager trainer's exact parameters and AdamW state and both draw the same noise
    (same device / host seeds), so one step from identical state must give the same losses and the same flat gradient
    buffers - whether the graph trainer ran it eagerly (first sight of a step variant), captured it, or replayed it.
    Gradients are compared rather than parameter updates: Adam's first steps are ~lr*sign(g), which turns rounding
    noise on near-zero gradients into full-size update differences (two EAGER runs with the same seeds already differ
    by 0.5-0.7 of the largest update after 5 steps - profiles/r02_parity_diagnostics.txt)."""
    import gigagan_pytorch_b200 as g
    from gigagan_pytorch_b200.trainer import cycle
    g.set_compute_dtype(torch.float32)
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
    reals = [torch.rand(4, 3, 64, 64, generator=torch.Generator().manual_seed(10 + s)).to(dev()) for s in range(8)]

    class Pool:
        batch_size = 4

        def __iter__(self):
            return iter(reals)

    gans, its = [], []
    for graphs in (False, True):
        torch.manual_seed(0)
        gan = g.GigaGAN(generator=gen, discriminator=disc, train_upsampler=upsampler, amp=amp, mixed_precision_type="bf16",
                        log_steps_every=10 ** 9, create_ema_generator_at_init=False, save_and_sample_every=0).to(dev())
        gan.use_cuda_graphs = graphs
        gan._ensure_optimizers()
        gans.append(gan)
        its.append(cycle(Pool()))
    eager, graph = gans
    # fp32 is the check of the replay logic.  bf16: both trainers sum through fp32 atomics in arbitrary order and every
    # difference passes bf16 roundings and LeakyReLU slopes; typical deviation 2e-3, one run in ~15 full-suite runs exceeded
    # the former 2e-2 gradient bound, hence the margin
    tol_loss, tol_grad = (2e-2, 5e-2) if amp else (1e-4, 1e-3)
    p_start = torch.cat([eager.G_opt.flat, eager.D_opt.flat]).clone()
    for step in range(1, 7):        # plain variant: eager warm-up, capture, replay; then the same for the penalty variant
        gp = step > 3
        for name in ("G_opt", "D_opt"):
            src, dst = getattr(eager, name), getattr(graph, name)
            for f in ("flat", "m", "v", "step_t"):
                getattr(dst, f).copy_(getattr(src, f))
        for gan in gans:
            for bank in gan._banks:
                bank.dirty = True                  # same step-variant key every time: (True, True)
        res = []
        for gan, it in zip(gans, its):
            torch.manual_seed(100 + step)          # device noise (also inside graph replays) and host patch selection
            d = gan.train_discriminator_step(dl_iter=it, apply_gradient_penalty=gp)
            gd = gan.D_opt.grad.clone()
            torch.manual_seed(200 + step)          # the generator step's noise must not depend on how many Philox offsets
            gl = gan.train_generator_step(batch_size=4, dl_iter=it)   # the (eager / captured / replayed) D step consumed
            gg = gan.G_opt.grad.clone()
            torch.cuda.synchronize()
            res.append((torch.stack([d.divergence.float(), d.multiscale_divergence.float(), d.gradient_penalty.float(),
                                     d.aux_reconstruction.float(), gl.divergence.float(),
                                     gl.multiscale_divergence.float()]).clone(), gd, gg))
        (la, gda, gga), (lb, gdb, ggb) = res
        assert torch.isfinite(la).all() and torch.isfinite(lb).all()
        if gp:
            assert la[2].item() > 0
        rel = ((la - lb).abs() / la.abs().clamp_min(0.5 if amp else 0.1)).max().item()
        assert rel < tol_loss, (step, rel, la.tolist(), lb.tolist())
        for what, x, y in (("D", gda, gdb), ("G", gga, ggb)):
            assert x.abs().max().item() > 0
            r = (x - y).abs().max().item() / x.abs().max().item()
            assert r < tol_grad, (step, what, r)
    moved = (torch.cat([eager.G_opt.flat, eager.D_opt.flat]) - p_start).abs().max().item()
    assert moved > 1e-4                            # the parameters did move
    assert graph.graph_kernel_launches > 0         # and the graph trainer did replay captured steps


# ------------------------------------------------------------------ text-conditioned path
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-4), (torch.bfloat16, 5e-2)])
def test_ka7_text_conditional(dtype, tol):
    """TextEncoder + cross attention in G, text-modulated predictors in D, against the reference's outputs and grads"""
    import gigagan_pytorch_b200 as g
    from gigagan_pytorch_b200 import ops
    fx = load("ka7_text_conditional.pt")
    g.set_compute_dtype(dtype)
    G = g.Generator(text_encoder=g.TextEncoder(**fx["te_cfg"]), **fx["gcfg"]).to(dev())
    G.load_state_dict(fx["gsd"])
    enc = fx["enc"].to(dev())
    gt, ft, tm = G.encode_text(text_encodings=enc)
    assert tm.tolist() == [[True] * 6, [True] * 4 + [False] * 2]
    torch.manual_seed(fx["noise_seed"])
    noises = []
    for r in [4, 8, 16, 32]:
        for _ in range(2):
            noises.append(torch.randn(2, 1, r, r).to(dev()))
    rgb, rgbs = G.forward_nhwc(noise=fx["z"].to(dev()), layer_noises=noises, global_text_tokens=gt,
                               fine_text_tokens=ft, text_mask=tm)
    out = ops.to_nchw(rgb, 3)
    assert relmax(out, fx["rgb"].to(dev())) < tol
    for a, b in zip(rgbs, fx["rgbs"]):
        assert relmax(ops.to_nchw(a, 3), b.to(dev())) < tol
    (out ** 2).mean().backward()
    named = dict(G.named_parameters())
    # Noise.weight gradients (shape (C,1,1), zero-initialised) are sums of +-gy*noise terms that cancel almost fully:
    # in bf16 the rounding of gy dominates them, so they are held to the fp32 bound only
    skip = (lambda k, v: dtype == torch.bfloat16 and v.ndim == 3 and v.shape[1:] == (1, 1))
    worst = max((relmax(named[k].grad, v.to(dev())), k) for k, v in fx["ggrads"].items() if not skip(k, v))
    assert worst[0] < tol * (5 if dtype == torch.float32 else 10), worst       # 16-channel toy model: bf16 noise is large
    D = g.Discriminator(text_encoder=g.TextEncoder(**fx["te_cfg"]), **fx["dcfg"]).to(dev())
    D.load_state_dict(fx["dsd"])
    img = fx["img"].to(dev())
    lo, ms, _ = D(img, D.real_images_to_rgbs(img), text_encodings=enc, calc_aux_loss=False)
    assert relmax(lo, fx["logits"].to(dev())) < tol
    for a, b in zip(ms, fx["ms"]):
        assert relmax(a, b.to(dev())) < tol
    (lo.sum() + sum((m ** 2).sum() for m in ms)).backward()
    named = dict(D.named_parameters())
    worst = max((relmax(named[k].grad, v.to(dev())), k) for k, v in fx["dgrads"].items())
    assert worst[0] < tol * 10, worst


def test_weight_bank_layouts():
    """one-launch re-layout of every conv weight of a flat parameter buffer: forward (O,KH,KW,Ipad) and flipped /
    swapped (Ipad,KH,KW,O) kernel layouts against torch permutes, including odd channel counts and 5-D filter banks"""
    from gigagan_pytorch_b200 import ops
    shapes = [(3, 64, 1, 1), (32, 3, 3, 3), (40, 3, 7, 7), (64, 32, 3, 3), (1, 48, 1, 1), (96, 160, 2, 2),
              (2, 33, 16, 3, 3), (16, 16, 4, 4)]
    import math
    n = sum(math.prod(s) for s in shapes)
    flat = torch.randn(n + 8, device=dev())
    params, off = [], 0
    for s in shapes:
        k = math.prod(s)
        params.append(flat[off:off + k].view(s))
        off += k
    pad = lambda c: 16 if c < 16 else (c + 15) // 16 * 16
    for dt in (torch.bfloat16, torch.float32):
        bank = ops.WeightBank(flat, params, dt, pad)
        bank.refresh()
        for p in params:
            for w in (p.unbind(0) if p.ndim == 5 else [p]):
                f, b = bank.lookup(w, pad(w.shape[1]), dt)
                ref = torch.nn.functional.pad(w.permute(0, 2, 3, 1), (0, pad(w.shape[1]) - w.shape[1])).to(dt)
                assert torch.equal(f, ref), (tuple(w.shape), dt)
                assert torch.equal(b, ref.flip((1, 2)).permute(3, 1, 2, 0)), (tuple(w.shape), dt)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape", [(32, 3, 3, 3), (48, 160, 1, 1), (16, 200, 3, 3), (8, 16, 7, 7)])
def test_wgrad_sink_equals_autograd_accumulation(dtype, shape):
    """conv weights whose .grad lives in the optimiser's flat buffer receive the wgrad kernel's result through
    gg_wgrad_sink (in-place +=) instead of autograd's accumulation: same numbers, including a second accumulation"""
    from gigagan_pytorch_b200 import ops
    O, I, k, _ = shape
    cin = 16 if I < 16 else I
    x = rn(1, 2, 16, 16, cin).to(dev()).to(dtype)
    if I < cin:
        x[..., I:] = 0
    w_ref = (rn(2, *shape) * 0.1).to(dev()).requires_grad_()
    w_snk = w_ref.detach().clone().requires_grad_()
    w_snk.grad = torch.zeros_like(w_snk)
    w_snk._gg_sink = True
    for w in (w_ref, w_snk):
        for rep in range(2):                                   # two backward passes: accumulation semantics
            y = ops.conv2d(x, w, None, pad=k // 2)
            (y.float() ** 2).sum().backward()
    assert relmax(w_snk.grad, w_ref.grad) < 1e-6


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 2e-2)])
@pytest.mark.parametrize("c", [64, 256, 512])
def test_fused_rmsnorm_matches_composed(dtype, tol, c):
    """ChannelRMSNorm: fused first-order kernels vs the any-order differentiable composition (and torch's formula)"""
    from gigagan_pytorch_b200 import modules as M
    x0 = rn(1, 3, 9, 7, c).to(dev())
    x0[0, 0, 0] = 0                                             # a zero row: F.normalize's eps clamp
    gamma0 = (1 + 0.1 * rn(2, c, 1, 1)).to(dev())
    gy = rn(3, 3, 9, 7, c).to(dev()).to(dtype)
    outs = []
    for fused in (True, False):
        x = x0.to(dtype).requires_grad_()
        gamma = gamma0.clone().requires_grad_()
        y = M.channel_rmsnorm(x, gamma, fused)
        gx, gg = torch.autograd.grad(y, (x, gamma), gy)
        outs.append((y, gx, gg))
    for a, b in zip(*outs):
        assert relmax(a, b) < tol
    ref = F.normalize(x0.to(dtype).float(), dim=-1) * c ** 0.5 * gamma0.reshape(-1)
    assert relmax(outs[0][0], ref) < tol


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 1e-2)])
@pytest.mark.parametrize("c,ns,p", [(8, 1, 64), (16, 2, 100), (32, 1, 4096), (64, 3, 25), (256, 2, 7), (48, 2, 9), (512, 1, 33)])
def test_dot_sc_channel_reductions(dtype, tol, c, ns, p):
    """per-(sample, channel) reductions sum_rows a*b (bias gradients, squeeze-excite means, modulation gradients):
    the narrow-map, 32-vector and generic kernels against torch"""
    from gigagan_pytorch_b200 import ops
    reps = 2
    a = rn(1, reps * ns * p, c).to(dev()).to(dtype)
    b = rn(2, reps * ns * p, c).to(dev()).to(dtype)
    out = ops.dot_sc(a, b, p, ns)
    ref = (a.float() * b.float()).view(reps, ns, p, c).sum(dim=(0, 2))
    assert relmax(out.view(ns, c), ref) < tol
    out1 = ops.dot_sc(a, None, p, ns)
    assert relmax(out1.view(ns, c), a.float().view(reps, ns, p, c).sum(dim=(0, 2))) < tol


def test_generate_after_training_uses_the_updated_weights():
    """the bf16 kernel-layout weight banks are rebuilt lazily (only for the model whose parameters changed): sampling
    right after an optimiser step must see the new generator weights, i.e. equal a bank-less generator with the same
    state_dict"""
    import gigagan_pytorch_b200 as g
    g.set_compute_dtype(torch.bfloat16)
    torch.manual_seed(0)
    gen = dict(dim_capacity=2, style_network=dict(dim=16, depth=2), image_size=64, dim_max=16, dim_latent=16,
               num_skip_layers_excite=2, unconditional=True, self_attn_resolutions=(16,), self_attn_dim_head=8, self_attn_heads=2)
    disc = dict(dim_capacity=2, dim_max=16, image_size=64, num_skip_layers_excite=2, unconditional=True,
                attn_resolutions=(8,), attn_dim_head=8, attn_heads=2, multiscale_input_resolutions=(32, 16, 8))
    gan = g.GigaGAN(generator=gen, discriminator=disc, amp=True, mixed_precision_type="bf16", log_steps_every=10 ** 9,
                    create_ema_generator_at_init=False).to(dev())
    reals = [torch.rand(4, 3, 64, 64, generator=torch.Generator().manual_seed(10 + s)).to(dev()) for s in range(4)]

    class Pool:
        batch_size = 4

        def __iter__(self):
            return iter(reals)

    gan.set_dataloader(Pool())
    gan(steps=3)
    z = rn(7, 2, 16).to(dev())
    torch.manual_seed(11)
    a = gan.generate(noise=z)
    fresh = g.Generator(**gen).to(dev())
    fresh.load_state_dict(gan.G.state_dict())
    fresh.eval()
    torch.manual_seed(11)
    b = fresh(noise=z)
    assert relmax(a, b) < 1e-6
    g.set_compute_dtype(torch.float32)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 2e-2)])
@pytest.mark.parametrize("co", [16, 32, 512])
def test_conv_bias_lrelu_backward_fused_matches_composed(dtype, tol, co):
    """first-order backward of conv + bias + LeakyReLU: the fused activation-gradient / bias-gradient kernel against
    the composed (any-order differentiable) route that create_graph=True selects"""
    from gigagan_pytorch_b200 import ops
    x = rn(1, 2, 12, 12, 16).to(dev()).to(dtype)
    gy = rn(4, 2, 12, 12, co).to(dev()).to(dtype)
    res = []
    for cg in (False, True):
        w = (rn(2, co, 16, 3, 3) * 0.1).to(dev()).requires_grad_()
        b = rn(3, co).to(dev()).requires_grad_()
        xx = x.clone().requires_grad_()
        y = ops.conv2d(xx, w, b, pad=1, act=1)
        res.append(torch.autograd.grad(y, (xx, w, b), gy, create_graph=cg))
    for a, b_ in zip(*res):
        assert relmax(a.detach(), b_.detach()) < tol


def test_flat_ema_update_matches_lerp():
    """EMA generator kept in one flat buffer: copy until update_after_step, then one axpby launch == per-parameter lerp"""
    import gigagan_pytorch_b200 as g
    g.set_compute_dtype(torch.float32)
    torch.manual_seed(0)
    gen = dict(dim_capacity=2, style_network=dict(dim=16, depth=2), image_size=32, dim_max=16, dim_latent=16,
               num_skip_layers_excite=2, unconditional=True, self_attn_resolutions=(16,), self_attn_dim_head=8, self_attn_heads=2)
    disc = dict(dim_capacity=2, dim_max=16, image_size=32, num_skip_layers_excite=2, unconditional=True,
                attn_resolutions=(8,), attn_dim_head=8, attn_heads=2, multiscale_input_resolutions=(16, 8))
    gan = g.GigaGAN(generator=gen, discriminator=disc, amp=False, log_steps_every=10 ** 9, create_ema_generator_at_init=False).to(dev())
    from gigagan_pytorch_b200.trainer import ema_current_decay
    gan._ensure_optimizers()
    gan.create_ema_generator(update_every=1, update_after_step=0, decay=0.9)
    gan._ema_update()                                              # step 0 <= update_after_step: plain copy
    for pe, p in zip(gan.G_ema.parameters(), gan.G.parameters()):
        assert torch.equal(pe, p)
    for k in range(1, 4):                                          # ema_pytorch warms the decay up: 0.37, 0.52, 0.60 ... -> 0.9
        before = [pe.detach().clone() for pe in gan.G_ema.parameters()]
        gan.G_opt.flat.add_(torch.randn_like(gan.G_opt.flat) * 0.1)    # "an optimiser step"
        gan._ema_update()
        d = ema_current_decay(k + 1, 0, 0.9)
        assert abs(d - min(0.9, 1 - (1 + k) ** (-2 / 3))) < 1e-12
        for b, pe, p in zip(before, gan.G_ema.parameters(), gan.G.parameters()):
            assert relmax(pe, torch.lerp(b, p.detach(), 1.0 - d)) < 1e-6
    assert gan._ema_flat is not None and gan._ema_flat.numel() == gan.G_opt.flat.numel()


@pytest.mark.parametrize("merge", [True, False])
def test_ka9_text_conditional_trainer_step(merge):
    """One text-conditional D step objective (hinge + multiscale + gradient penalty) and one G step objective through the
    TRAINER (GigaGAN._d_objective / _g_objective with pre-encoded text_encodings, ref :2263-2417 / :2518-2551) against
    the same quantities built from the reference's modules (fixture ka9)."""
    import gigagan_pytorch_b200 as g
    fx = load("ka9_text_step.pt")
    g.set_compute_dtype(torch.float32)
    G = g.Generator(text_encoder=g.TextEncoder(**fx["te_cfg"]), **fx["gcfg"])
    D = g.Discriminator(text_encoder=g.TextEncoder(**fx["te_cfg"]), **fx["dcfg"])
    G.load_state_dict(fx["gsd"]); D.load_state_dict(fx["dsd"])
    gan = g.GigaGAN(generator=G, discriminator=D, amp=False, log_steps_every=10 ** 9, create_ema_generator_at_init=False,
                    discr_aux_recon_loss_weight=0., matching_awareness_loss_weight=0.,
                    generator_contrastive_loss_weight=0.).to(dev())
    gan.merge_real_fake = merge
    gan.G.train(); gan.D.train()
    enc, z, real = fx["enc"].to(dev()), fx["z"].to(dev()), fx["real"].to(dev())
    total, (div, ms, gp, _) = gan._d_objective(real, z, True, True, text=enc)
    total.backward(inputs=list(gan.D.parameters()))
    tol = 2e-4
    for k, v in (("total", total), ("divergence", div), ("multiscale", ms), ("gradient_penalty", gp)):
        assert relmax(v.detach(), fx["dloss"][k].to(dev())) < tol * 5, (k, v.item(), fx["dloss"][k].item())
    named = dict(gan.D.named_parameters())
    worst = max((relmax(named[k].grad, v.to(dev())), k) for k, v in fx["dgrads"].items())
    assert worst[0] < tol * 10, worst
    for p in gan.D.parameters():
        p.grad = None
        p.requires_grad_(False)
    total, (gdiv, gms) = gan._g_objective(z, True, enc)
    total.backward(inputs=list(gan.G.parameters()))
    for k, v in (("total", total), ("divergence", gdiv), ("multiscale", gms)):
        assert relmax(v.detach(), fx["gloss"][k].to(dev())) < tol * 5, (k, v.item(), fx["gloss"][k].item())
    named = dict(gan.G.named_parameters())
    # Noise.weight gradients depend on the per-layer noise images (device RNG here, CPU RNG in the fixture): not compared
    skip = lambda k, v: v.ndim == 3 and v.shape[1:] == (1, 1) and ".1." in k
    worst = max((relmax(named[k].grad, v.to(dev())), k) for k, v in fx["ggrads"].items() if not skip(k, v))
    assert worst[0] < tol * 10, worst


def test_text_conditional_trainer_runs_with_graphs():
    """GigaGAN(steps=...) on a conditional dataset yielding (images, text_encodings): eager warm-up, capture, replay"""
    import gigagan_pytorch_b200 as g
    fx = load("ka9_text_step.pt")
    g.set_compute_dtype(torch.float32)
    torch.manual_seed(0)
    gan = g.GigaGAN(generator=dict(fx["gcfg"], text_encoder=dict(fx["te_cfg"])),
                    discriminator=dict(fx["dcfg"], text_encoder=dict(fx["te_cfg"])), amp=True, mixed_precision_type="bf16",
                    log_steps_every=10 ** 9, create_ema_generator_at_init=True, matching_awareness_loss_weight=0.,
                    generator_contrastive_loss_weight=0., save_and_sample_every=0).to(dev())
    gan.use_cuda_graphs = True
    items = []
    for s in range(4):
        enc = torch.randn(4, 6, 32, generator=torch.Generator().manual_seed(s))
        enc[1, 3:] = 0.
        items.append((torch.rand(4, 3, 32, 32, generator=torch.Generator().manual_seed(10 + s)), enc))

    class Pool:
        batch_size = 4

        def __iter__(self):
            return iter(items)

    gan.set_dataloader(Pool())
    p0 = torch.cat([p.detach().flatten().float().clone() for p in gan.G.parameters()])
    gan(steps=6)
    torch.cuda.synchronize()
    p1 = torch.cat([p.detach().flatten().float() for p in gan.G.parameters()])
    assert torch.isfinite(p1).all() and (p1 - p0).abs().max().item() > 1e-5
    out = gan.generate(batch_size=2, text_encodings=items[0][1][:2].to(dev()))
    assert out.shape == (2, 3, 32, 32) and torch.isfinite(out).all()
