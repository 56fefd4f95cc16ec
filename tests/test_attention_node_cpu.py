"""Derivative plumbing of the any-order attention node (ops.ComposedAttnFn / ComposedAttnBwdFn) on a machine without a
GPU: the four raw kernel calls are replaced by torch stand-ins (in THIS test only) and first and second derivatives are
compared with torch autograd of the plain formula (reference gigagan_pytorch.py:562-592 under gradient_penalty :138-155).
What is checked is the hand-written chain rule: the K-concatenated d/d(dS) product, the deferred rank-d gradient of the
probabilities, the on-the-fly addend of the softmax backward and the order in which autograd runs the two nodes."""
import pytest
import torch

from gigagan_pytorch_b200 import ops


@pytest.fixture
def torch_kernels(monkeypatch):
    def k_bmm(a, b, alpha=1.0, out=False):
        return alpha * (a @ b)

    def k_softmax(s, bias, P, Ns):
        return torch.softmax(s if bias is None else s + bias, dim=-1)

    def k_softmax_bwd(p, gp, gp2=None):
        g = gp if gp2 is None else gp + gp2
        return p * (g - (p * g).sum(-1, keepdim=True))

    def k_softmax_bwd2(p, gp, G):
        r_pg = (p * gp).sum(-1, keepdim=True)
        r_Gp = (G * p).sum(-1, keepdim=True)
        return G * (gp - r_pg) - gp * r_Gp, p * (G - r_Gp)

    monkeypatch.setattr(ops, "_k_bmm", k_bmm)
    monkeypatch.setattr(ops, "_k_softmax", k_softmax)
    monkeypatch.setattr(ops, "_k_softmax_bwd", k_softmax_bwd)
    monkeypatch.setattr(ops, "_k_softmax_bwd2", k_softmax_bwd2)
    monkeypatch.setattr(ops, "axpby", lambda a, x, b=0.0, y=None: a * x + (0 if y is None else b * y))


def _ref(qa, ka, v, mask, alpha):
    s = alpha * (qa @ ka.transpose(-1, -2))
    if mask is not None:
        s = s + mask
    return torch.softmax(s, dim=-1) @ v


def _objective(fn, qa, ka, v, mask, alpha, w_o, w_g, penalty):
    """hinge-like first-order term + (optionally) a gradient-penalty-like term on d(out)/d(inputs)."""
    o = fn(qa, ka, v, mask, alpha)
    total = (o * w_o).sum()
    if penalty:
        gq, gk, gv = torch.autograd.grad((o * w_g).sum(), (qa, ka, v), create_graph=True, retain_graph=True)
        total = total + (gq ** 2).sum() + 0.5 * (gk ** 2).sum() + 0.25 * (gv ** 2).sum()
    return total


@pytest.mark.parametrize("penalty", [False, True])
@pytest.mark.parametrize("masked", [False, True])
def test_attention_node_matches_torch_autograd(torch_kernels, penalty, masked):
    torch.manual_seed(0)
    b, h, n, m, D, d = 2, 3, 5, 7, 6, 4
    base = [torch.randn(b, n, h, D, dtype=torch.float64), torch.randn(b, m, h, D, dtype=torch.float64),
            torch.randn(b, m, h, d, dtype=torch.float64)]
    w_o, w_g = torch.randn(b, h, n, d, dtype=torch.float64), torch.randn(b, h, n, d, dtype=torch.float64)
    mask = None
    if masked:
        mask = torch.zeros(1, m, dtype=torch.float64)
        mask[:, -2:] = -1e30
    grads = []
    for fn in (_ref, ops.composed_attention):
        leaves = [t.clone().requires_grad_() for t in base]
        # permuted views, as the modules pass them ((b, tokens, heads, dim) storage)
        qa, ka, v = (t.permute(0, 2, 1, 3) for t in leaves)
        total = _objective(fn, qa, ka, v, mask, 0.37, w_o, w_g, penalty)
        total.backward()
        grads.append([t.grad.clone() for t in leaves] + [total.detach()])
    for a, r in zip(grads[1], grads[0]):
        assert torch.allclose(a, r, rtol=1e-9, atol=1e-11), (a - r).abs().max()


def test_attention_node_partial_cotangents(torch_kernels):
    """only some of (dqa, dka, dv) enter the second-order objective: the None branches of the second-order node"""
    torch.manual_seed(1)
    b, h, n, m, D = 1, 2, 4, 6, 5
    base = [torch.randn(b, h, n, D, dtype=torch.float64), torch.randn(b, h, m, D, dtype=torch.float64),
            torch.randn(b, h, m, D, dtype=torch.float64)]
    for pick in ((0,), (1,), (2,), (0, 2), (1, 2)):
        out = []
        for fn in (_ref, ops.composed_attention):
            qa, ka, v = (t.clone().requires_grad_() for t in base)
            o = fn(qa, ka, v, None, 0.5)
            g = torch.autograd.grad(o.sum(), (qa, ka, v), create_graph=True, retain_graph=True)
            total = sum((g[i] ** 2).sum() for i in pick) + (o ** 2).sum()
            total.backward()
            out.append([qa.grad, ka.grad, v.grad])
        for a, r in zip(out[1], out[0]):
            assert torch.allclose(a, r, rtol=1e-9, atol=1e-11), (pick, (a - r).abs().max())


# ------------------------------------------------------------------ operand builder (ops.AttnAugmentFn)
PADV = -40.0          # the -1e30 of the padding rows, finite here so that float64 softmax stays exact


def _aug_ref(q4, v4, null_kv, Lp):
    """plain torch: what csrc/attn_augment.cu builds (hi = t exactly, lo = 0), differentiable by autograd"""
    n, seq, h, d = q4.shape
    L = seq + 1
    kf = torch.cat([null_kv[0][None, None].expand(n, 1, h, d), q4, q4.new_zeros(n, Lp - L, h, d)], dim=1)
    vf = torch.cat([null_kv[1][None, None].expand(n, 1, h, d), v4, q4.new_zeros(n, Lp - L, h, d)], dim=1)
    pad = q4.new_zeros(1, Lp, 1)
    pad[:, L:] = PADV
    t = -0.5 * (kf * kf).sum(-1) + pad
    ka = torch.cat([kf, t.unsqueeze(-1), q4.new_zeros(n, Lp, h, 15)], dim=-1)
    qa = torch.cat([q4, q4.new_ones(n, seq, h, 2), q4.new_zeros(n, seq, h, 14)], dim=-1)
    return qa, ka, vf


@pytest.fixture
def torch_aug_kernels(monkeypatch):
    def fwd(q4, v4, null_kv, Lp):
        with torch.no_grad():
            return tuple(t.clone() for t in _aug_ref(q4, v4, null_kv, Lp))

    def bwd(dqa, dka, dvf, q4, null_kv):
        seq = q4.shape[1]
        ghi = dka[:, 1:seq + 1, :, 64:65]
        dq = dqa[..., :64] + dka[:, 1:seq + 1, :, :64] - ghi * q4
        dv = dvf[:, 1:seq + 1]
        dnull = torch.stack([(dka[:, 0, :, :64] - dka[:, 0, :, 64:65] * null_kv[0][None]).sum(0), dvf[:, 0].sum(0)])
        return dq, dv, dnull

    def bwd2(wq, wv, wnull, q4, null_kv, dka):
        n, seq, h, d = q4.shape
        Lp = dka.shape[1]
        g_dqa = torch.cat([wq, wq.new_zeros(n, seq, h, 16)], dim=-1)
        g_dka = dka.new_zeros(dka.shape)
        g_dka[:, 1:seq + 1, :, :64] = wq
        g_dka[:, 1:seq + 1, :, 64] = -(wq * q4).sum(-1)
        g_dka[:, 0, :, :64] = wnull[0][None]
        g_dka[:, 0, :, 64] = -(wnull[0] * null_kv[0]).sum(-1)[None]
        g_dvf = dka.new_zeros(n, Lp, h, d)
        g_dvf[:, 1:seq + 1] = wv
        g_dvf[:, 0] = wnull[1][None]
        g_q = -dka[:, 1:seq + 1, :, 64:65] * wq
        g_null = torch.stack([-(dka[:, 0, :, 64:65] * wnull[0][None]).sum(0), torch.zeros_like(wnull[1])])
        return g_dqa, g_dka, g_dvf, g_q, g_null

    monkeypatch.setattr(ops, "_k_aug_fwd", fwd)
    monkeypatch.setattr(ops, "_k_aug_bwd", bwd)
    monkeypatch.setattr(ops, "_k_aug_bwd2", bwd2)
    monkeypatch.setattr(ops, "_c", lambda t: t.contiguous())


def test_attention_operand_builder_derivatives(torch_kernels, torch_aug_kernels):
    """first and second derivative of the operand builder (ops.AttnAugmentFn / AttnAugmentBwdFn, the maths of
    csrc/attn_augment.cu) chained with the attention node, against torch autograd of the plain formulas on a
    penalty-shaped objective; the null key/value parameter takes part"""
    torch.manual_seed(2)
    n, seq, h, d, Lp = 2, 5, 2, 64, 8
    base = [torch.randn(n, seq, h, d, dtype=torch.float64) * 0.3, torch.randn(n, seq, h, d, dtype=torch.float64),
            torch.randn(2, h, d, dtype=torch.float64) * 0.3]
    w_o, w_g = torch.randn(n, h, seq, d, dtype=torch.float64), torch.randn(n, h, seq, d, dtype=torch.float64)
    res = []
    for custom in (False, True):
        q4, v4, nk = (t.clone().requires_grad_() for t in base)
        if custom:
            qa, ka, vf = ops.attn_augment(q4, v4, nk, Lp)
            o = ops.composed_attention(qa.permute(0, 2, 1, 3), ka.permute(0, 2, 1, 3), vf.permute(0, 2, 1, 3), None, 0.25)
        else:
            qa, ka, vf = _aug_ref(q4, v4, nk, Lp)
            o = _ref(qa.permute(0, 2, 1, 3), ka.permute(0, 2, 1, 3), vf.permute(0, 2, 1, 3), None, 0.25)
        gq, gv = torch.autograd.grad((o * w_g).sum(), (q4, v4), create_graph=True, retain_graph=True)
        total = (o * w_o).sum() + (gq ** 2).sum() + 0.5 * (gv ** 2).sum()
        total.backward()
        res.append([q4.grad, v4.grad, nk.grad, total.detach()])
    for a, r in zip(res[1], res[0]):
        assert torch.allclose(a, r, rtol=1e-8, atol=1e-10), (a - r).abs().max()
