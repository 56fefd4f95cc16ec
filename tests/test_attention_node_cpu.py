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
