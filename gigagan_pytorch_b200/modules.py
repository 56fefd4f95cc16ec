"""Drop-in GigaGAN modules: same constructor keywords, parameter names/shapes/creation order (so state_dicts and
seeded initialisation match lucidrains/gigagan-pytorch @ 0806433f), forward passes computed by the sm_100a kernels
behind ``ops``.  Reference: gigagan_pytorch/gigagan_pytorch.py (line numbers cited per class).

Internally feature maps are NHWC in the compute dtype; the public ``forward`` of every class takes and returns the
reference's NCHW fp32 tensors.  ``forward_nhwc`` is the layout-preserving entry the trainer uses.
"""
from __future__ import annotations

import math
import os
from functools import partial
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F
from torch import nn

from . import ops
from .ops import U_GELU, U_INVNORM, U_RSQRT_EPS8, U_SIGMOID, U_SILU

_COMPUTE = {"dtype": torch.float32, "fused_attention": True,
            # gradient-penalty attention as one any-order autograd node (ops.ComposedAttnFn); GG_ATTN_NODE=0 keeps the
            # composition from primitives (A/B measurements, parity cross-checks)
            "attention_node": os.environ.get("GG_ATTN_NODE", "1") != "0",
            # ... and its augmented operands (shared-QK L2 form) from one kernel; GG_ATTN_AUGMENT=0: concatenations
            "attention_augment": os.environ.get("GG_ATTN_AUGMENT", "1") != "0"}


def set_compute_dtype(dtype):
    """fp32 (FFMA kernels, 1e-5 parity) or bf16 (tcgen05 where dense) for all modules created/run afterwards."""
    assert dtype in (torch.float32, torch.bfloat16)
    _COMPUTE["dtype"] = dtype


def compute_dtype():
    return _COMPUTE["dtype"]


def exists(v):
    return v is not None


def is_power_of_two(n):
    return math.log2(n).is_integer()


def img_cpad(c):
    """channel count of image-like NHWC tensors (rgb, images): padded to 16 in bf16 so that every convolution that
    touches them is tcgen05-eligible (zero channels, zero weights); exact in fp32."""
    return 16 if (_COMPUTE["dtype"] == torch.bfloat16 and c < 16) else c


# ----------------------------------------------------------------------------- small functional blocks (NHWC)
def channel_rmsnorm(x, gamma, fused=None):
    """ref :224-232.  x NHWC; gamma (C,1,1).  ``fused`` (default: the fused-attention switch): one first-order
    kernel pair instead of the any-order differentiable composition."""
    c = x.shape[-1]
    fused = _COMPUTE["fused_attention"] if fused is None else fused
    if fused and ops.rmsnorm_fused_ok(x):
        return ops.rmsnorm_fused(x, gamma, c ** 0.5)
    ss = ops.rowdot(x, x)
    inv = ops.unary(U_INVNORM, ss)
    y = ops.scale_rows(x, inv)
    g = ops.axpby(c ** 0.5, gamma.reshape(1, c))
    return ops.scale_channels(y, g, x.numel() // c, 1)


def squeeze_excite(seq, x):
    """ref :297-307 -> fp32 (N, C_out) gates."""
    h = ops.mean_hw(x)
    h = ops.linear(h, seq[1].weight, seq[1].bias)
    h = ops.unary(U_SILU, h)
    h = ops.linear(h, seq[3].weight, seq[3].bias)
    return ops.unary(U_SIGMOID, h)


def apply_excite(x, gates):
    n, h, w, c = x.shape
    return ops.scale_channels(x, gates, h * w, gates.shape[0])


def SqueezeExciteParams(dim, dim_out, reduction=4, dim_min=32):
    dim_hidden = max(dim_out // reduction, dim_min)
    return nn.Sequential(nn.Identity(), nn.Linear(dim, dim_hidden), nn.SiLU(), nn.Linear(dim_hidden, dim_out),
                         nn.Sigmoid(), nn.Identity())


class Blur(nn.Module):
    """Parameter-less; keeps the reference's ``f`` buffer for state_dict compatibility (ref :246-255)."""

    def __init__(self):
        super().__init__()
        self.register_buffer("f", torch.Tensor([1, 2, 1]))


def UpsampleParams(*_):
    return nn.Sequential(nn.Identity(), Blur())


class ChannelRMSNorm(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.scale = dim ** 0.5
        self.gamma = nn.Parameter(torch.ones(dim, 1, 1))

    def forward_nhwc(self, x, fused=None):
        return channel_rmsnorm(x, self.gamma, fused)


# ----------------------------------------------------------------------------- AdaptiveConv2DMod (ref :315-409)
class AdaptiveConv2DMod(nn.Module):
    def __init__(self, dim, dim_out, kernel, *, demod=True, stride=1, dilation=1, eps=1e-8, num_conv_kernels=1):
        super().__init__()
        assert stride == 1 and dilation == 1, "the training hot path only uses stride 1 / dilation 1"
        self.eps, self.dim_out, self.kernel = eps, dim_out, kernel
        self.stride, self.dilation = stride, dilation
        self.adaptive = num_conv_kernels > 1
        self.weights = nn.Parameter(torch.randn((num_conv_kernels, dim_out, dim, kernel, kernel)))
        self.demod = demod
        nn.init.kaiming_normal_(self.weights, a=0, mode="fan_in", nonlinearity="leaky_relu")

    def forward_nhwc(self, x, mod, kernel_mod=None, out_pad=0, fused=None):
        """``fused`` (default: the first-order switch shared with the attention blocks): low-resolution layers run as
        one fused autograd node (ops.SharedBankConvFn); False keeps the any-order differentiable composition."""
        b = x.shape[0]
        if mod.shape[0] != b:                       # scale-major repeat, ref :365-366
            mod = mod.repeat(b // mod.shape[0], 1)
        if self.adaptive:
            assert exists(kernel_mod) and kernel_mod.numel() > 0
            if kernel_mod.shape[0] != b:
                kernel_mod = kernel_mod.repeat(b // kernel_mod.shape[0], 1)
        else:
            assert not exists(kernel_mod) or kernel_mod.numel() == 0
            kernel_mod = None
        hw = x.shape[1] * x.shape[2]
        # shared-bank form: always for the 4x4 / 8x8 maps (B private filters leave the 128-row MMA tile empty); up to
        # 16x16 for the demodulated 3x3 layers, where B private 512x512x9 filters cost more HBM traffic (filter write,
        # two reads, a flipped copy, an fp32 gradient of twice that size) than the n-fold convolution costs tensor time
        if x.dtype == torch.bfloat16 and self.eps == 1e-8 and (hw < 128 or (self.demod and hw <= 256)):
            fused = _COMPUTE["fused_attention"] if fused is None else fused
            if fused and self.demod and out_pad <= self.dim_out:
                return ops.shared_bank_conv(x, self.weights, mod, kernel_mod, self.eps)
            return self._forward_shared_bank(x, mod, kernel_mod, out_pad)
        w = ops.AdaConvWeightsFn.apply(self.weights, mod, kernel_mod, self.demod, self.eps, x.dtype, out_pad)
        return ops.conv2d_prepared(x, w, pad=(self.kernel - 1) // 2, per_sample=True)

    def _forward_shared_bank(self, x, mod, kernel_mod, out_pad=0):
        """Low-resolution layers (4x4, 8x8): B private 512x512x9 filters are weight-bandwidth bound and leave the
        128-row MMA tile empty.  Use  y_b = d_b * sum_n a_bn conv(x_b * s_b, W_n)  with the SHARED bank W_n
        (SURVEY.md section 7 identity): dense tcgen05 convolutions whose M tile spans images; the demodulation
        d_bo = rsqrt(max(sum_i s_bi^2 sum_nm a_bn a_bm <W_n,W_m>_k [o,i], eps)) is built from small fp32 operators."""
        b, h, w_, i = x.shape
        n, o, _, k, _ = self.weights.shape
        one = torch.ones((1, i), dtype=torch.float32, device=x.device)
        s = ops.add_channels(mod.float().contiguous(), one, b, 1)                    # (b, i) = mod + 1
        xs = ops.scale_channels(x, s, h * w_, b)
        if self.adaptive:
            a = ops.softmax(kernel_mod.float().contiguous())                          # (b, n)
        else:
            a = torch.ones((b, 1), dtype=torch.float32, device=x.device)
        pad = (self.kernel - 1) // 2
        y = None
        for j in range(n):
            wj = self.weights[j]
            if out_pad > o:                 # zero filters up to the padded channel count of image-like tensors
                wj = F.pad(wj, (0, 0, 0, 0, 0, 0, 0, out_pad - o))
            yj = ops.conv2d(xs, wj, pad=pad)
            if self.adaptive:
                yj = ops.scale_channels(yj, a[:, j:j + 1].expand(b, yj.shape[-1]).contiguous(), h * w_, b)
            y = yj if y is None else ops.add(y, yj)
        if not self.demod:
            return y
        s2 = ops.mul(s, s)
        wf = self.weights.reshape(n, o * i, k * k)
        t = None
        for j in range(n):
            for l in range(j, n):
                gram = ops.rowdot(wf[j], wf[l]).reshape(o, i)                         # <W_j, W_l> over the taps
                term = ops.linear(s2, gram)                                           # (b, o)
                if self.adaptive:
                    coef = ops.mul(a[:, j].contiguous(), a[:, l].contiguous())
                    term = ops.scale_rows(term, coef if j == l else ops.axpby(2.0, coef))
                t = term if t is None else ops.add(t, term)
        d = ops.unary(U_RSQRT_EPS8, t)
        return ops.scale_channels(y, d, h * w_, b)

    def forward(self, fmap, mod, kernel_mod=None):
        x = ops.to_nhwc(fmap, fmap.shape[1], compute_dtype())
        y = self.forward_nhwc(x, mod.float(), kernel_mod)
        return ops.to_nchw(y, self.dim_out)


# ----------------------------------------------------------------------------- attention (ref :513-594, :726-760)
class SelfAttention(nn.Module):
    def __init__(self, dim, dim_head=64, heads=8, dot_product=False):
        super().__init__()
        self.heads, self.dim_head = heads, dim_head
        self.scale = dim_head ** -0.5
        dim_inner = dim_head * heads
        self.dot_product = dot_product
        self.norm = ChannelRMSNorm(dim)
        self.to_q = nn.Conv2d(dim, dim_inner, 1, bias=False)
        self.to_k = nn.Conv2d(dim, dim_inner, 1, bias=False) if dot_product else None
        self.to_v = nn.Conv2d(dim, dim_inner, 1, bias=False)
        self.null_kv = nn.Parameter(torch.randn(2, heads, dim_head))
        self.to_out = nn.Conv2d(dim_inner, dim, 1, bias=False)

    def forward_nhwc(self, x, residual=None, fused=None):
        n, hh, ww, _ = x.shape
        seq, heads, d = hh * ww, self.heads, self.dim_head
        fused = _COMPUTE["fused_attention"] if fused is None else fused
        xn = self.norm.forward_nhwc(x, fused)
        q = ops.conv2d(xn, self.to_q.weight)
        v = ops.conv2d(xn, self.to_v.weight)
        k = ops.conv2d(xn, self.to_k.weight) if exists(self.to_k) else q
        if fused:
            qv = q.view(n, seq, heads * d)
            kv = k.view(n, seq, heads * d) if exists(self.to_k) else qv
            o = ops.fused_attention(qv, kv, v.view(n, seq, heads * d), self.null_kv, heads, self.scale,
                                    l2=not self.dot_product)
            o = o.view(n, hh, ww, heads * d)
        else:
            o = self._composed(q, k, v, n, seq, heads, d).reshape(n, hh, ww, heads * d)
        return ops.conv2d(o, self.to_out.weight, res=residual)

    def _composed(self, q, k, v, n, seq, heads, d):
        """Attention from closed-under-differentiation primitives (used inside the gradient penalty).  The key axis
        (null key + tokens) is zero-padded to a multiple of 64 with a -1e30 logit bias so that every matrix has
        16-byte aligned rows for the TMA-fed tcgen05 batched GEMMs; padded columns get probability exactly 0."""
        L = seq + 1
        Lp = (L + 63) // 64 * 64 if (q.dtype == torch.bfloat16 and seq >= 64 and d % 16 == 0) else L
        if (_COMPUTE["attention_node"] and _COMPUTE["attention_augment"] and not self.dot_product and k is q
                and q.dtype == torch.bfloat16 and d == 64 and Lp % 64 == 0):
            # shared-QK L2 form on the benchmarked path: the augmented operands (see the composition below) come from one
            # kernel, and one for each of their derivatives, instead of reductions, casts, fills and concatenations
            qa, ka, va = ops.attn_augment(q.view(n, seq, heads, d), v.view(n, seq, heads, d), self.null_kv, Lp)
            return ops.composed_attention(qa.permute(0, 2, 1, 3), ka.permute(0, 2, 1, 3), va.permute(0, 2, 1, 3), None,
                                          2.0 * self.scale).permute(0, 2, 1, 3)
        # (the broadcast null rows are materialised: a stride-0 input sends the whole cat down ATen's generic gather path)
        parts_k = [self.null_kv[0].to(q.dtype)[None, None].expand(n, 1, heads, d).contiguous(), k.view(n, seq, heads, d)]
        parts_v = [self.null_kv[1].to(q.dtype)[None, None].expand(n, 1, heads, d).contiguous(), v.view(n, seq, heads, d)]
        if Lp > L:
            z = torch.zeros((n, Lp - L, heads, d), dtype=q.dtype, device=q.device)
            parts_k.append(z)
            parts_v.append(z)
        kf = torch.cat(parts_k, dim=1)                                     # (n, Lp, heads, d)
        vf = torch.cat(parts_v, dim=1)
        q4 = q.view(n, seq, heads, d).permute(0, 2, 1, 3)                  # (n, heads, seq, d) view
        kt = kf.permute(0, 2, 3, 1)                                        # (n, heads, d, Lp) view
        mask = None
        if Lp > L:
            mask = torch.zeros((1, Lp), dtype=torch.float32, device=q.device)
            mask[:, L:] = -1e30
        node = _COMPUTE["attention_node"]                                  # one any-order node instead of primitives
        if self.dot_product:
            if node:
                return ops.composed_attention(q4, kf.permute(0, 2, 1, 3), vf.permute(0, 2, 1, 3), mask,
                                              self.scale).permute(0, 2, 1, 3)
            s = ops.bmm(q4, kt, alpha=self.scale)
            p = ops.softmax(s, mask, s.numel() // Lp, 1) if mask is not None else ops.softmax(s)
        else:
            # -|q-k|^2 * scale == (2 q.k - |k|^2) * scale up to a per-row constant (softmax-invariant).  The key term is
            # folded INTO the product: q' = [q, 1, 1, 0..], k' = [k, hi, lo, 0..] with hi + lo = -|k|^2 / 2 split into
            # two storage-precision numbers (exact to ~2^-16 in bf16), padded keys get hi = -1e30.  The logit matrix
            # then needs no bias: on the gradient-penalty path that removes a (tokens x keys)-sized column reduction in
            # the first backward and a zero fill, a broadcast and an accumulation of that size in the second.
            ksq = ops.rowdot(kf, kf)                                        # (n, Lp, heads) fp32
            t = ops.axpby(-0.5, ksq)
            if mask is not None:
                mrow = mask.reshape(Lp, 1).expand(Lp, heads).reshape(1, Lp * heads).contiguous()
                t = ops.add_channels(t.reshape(n, Lp * heads), mrow, n, 1).reshape(n, Lp, heads)
            hi = t.to(q.dtype)
            pad = d % 16 == 0 and q.dtype == torch.bfloat16
            extra = [hi.unsqueeze(-1)]
            if q.dtype == torch.bfloat16:
                extra.append(ops.axpby(1.0, t, -1.0, hi.float()).to(q.dtype).unsqueeze(-1))
            nx = len(extra)
            width = 16 if pad else nx
            if width > nx:
                extra.append(torch.zeros((n, Lp, heads, width - nx), dtype=q.dtype, device=q.device))
            ka = torch.cat([kf] + extra, dim=-1)                            # (n, Lp, heads, d + width)
            qx = [q.view(n, seq, heads, d), torch.ones((n, seq, heads, nx), dtype=q.dtype, device=q.device)]
            if width > nx:
                qx.append(torch.zeros((n, seq, heads, width - nx), dtype=q.dtype, device=q.device))
            qa = torch.cat(qx, dim=-1)
            if node:
                return ops.composed_attention(qa.permute(0, 2, 1, 3), ka.permute(0, 2, 1, 3), vf.permute(0, 2, 1, 3),
                                              None, 2.0 * self.scale).permute(0, 2, 1, 3)
            s = ops.bmm(qa.permute(0, 2, 1, 3), ka.permute(0, 2, 3, 1), alpha=2.0 * self.scale)
            p = ops.softmax(s)
        o = ops.bmm(p, vf.permute(0, 2, 1, 3), out_bmhn=True)             # physical (n, seq, heads, d)
        return o.permute(0, 2, 1, 3)

    def forward(self, fmap):
        x = ops.to_nhwc(fmap, fmap.shape[1], compute_dtype())
        return ops.to_nchw(self.forward_nhwc(x), fmap.shape[1])


def FeedForwardParams(dim, mult=4):
    dim_hidden = int(dim * mult)
    return nn.Sequential(ChannelRMSNorm(dim), nn.Conv2d(dim, dim_hidden, 1), nn.GELU(), nn.Conv2d(dim_hidden, dim, 1))


class SelfAttentionBlock(nn.Module):
    def __init__(self, dim, dim_head=64, heads=8, ff_mult=4, dot_product=False):
        super().__init__()
        self.attn = SelfAttention(dim=dim, dim_head=dim_head, heads=heads, dot_product=dot_product)
        self.ff = FeedForwardParams(dim, ff_mult)

    def forward_nhwc(self, x, fused=None):
        x = self.attn.forward_nhwc(x, residual=x, fused=fused)
        h = self.ff[0].forward_nhwc(x, fused)
        h = ops.conv2d(h, self.ff[1].weight, self.ff[1].bias)
        h = ops.unary(U_GELU, h)
        return ops.conv2d(h, self.ff[3].weight, self.ff[3].bias, res=x)

    def forward(self, fmap):
        x = ops.to_nhwc(fmap, fmap.shape[1], compute_dtype())
        return ops.to_nchw(self.forward_nhwc(x), fmap.shape[1])


# ----------------------------------------------------------------------------- style network (ref :871-921)
class EqualLinear(nn.Module):
    def __init__(self, dim, dim_out, lr_mul=1, bias=True):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(dim_out, dim))
        if bias:
            self.bias = nn.Parameter(torch.zeros(dim_out))
        self.lr_mul = lr_mul


class StyleNetwork(nn.Module):
    def __init__(self, dim, depth, lr_mul=0.1, dim_text_latent=0):
        super().__init__()
        self.dim, self.dim_text_latent = dim, dim_text_latent
        layers = []
        for i in range(depth):
            dim_in = (dim + dim_text_latent) if i == 0 else dim
            layers.extend([EqualLinear(dim_in, dim, lr_mul), nn.LeakyReLU(0.2)])
        self.net = nn.Sequential(*layers)

    def forward(self, x, text_latent=None):
        x = x.float()
        inv = ops.unary(U_INVNORM, ops.rowdot(x, x))
        x = ops.scale_rows(x, inv)
        if self.dim_text_latent > 0:
            assert exists(text_latent)
            x = torch.cat((x, text_latent.float()), dim=-1)
        for m in self.net:
            if isinstance(m, EqualLinear):
                x = ops.leaky_relu(ops.linear(x, m.weight, ops.axpby(m.lr_mul, m.bias), alpha=m.lr_mul))
        return x


class Noise(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.weight = nn.Parameter(torch.zeros(dim, 1, 1))


# ----------------------------------------------------------------------------- generator (ref :947-1250)
class BaseGenerator(nn.Module):
    pass


class Generator(BaseGenerator):
    def __init__(self, *, image_size, dim_capacity=16, dim_max=2048, channels=3,
                 style_network: StyleNetwork | Dict | None = None, style_network_dim=None, text_encoder=None,
                 dim_latent=512, self_attn_resolutions: Tuple[int, ...] = (32, 16), self_attn_dim_head=64,
                 self_attn_heads=8, self_attn_dot_product=True, self_attn_ff_mult=4,
                 cross_attn_resolutions: Tuple[int, ...] = (32, 16), cross_attn_dim_head=64, cross_attn_heads=8,
                 cross_attn_ff_mult=4, num_conv_kernels=2, num_skip_layers_excite=0, unconditional=False,
                 pixel_shuffle_upsample=False):
        super().__init__()
        assert not pixel_shuffle_upsample, "pixel_shuffle_upsample is outside the hot path"
        self.channels = channels
        if isinstance(style_network, dict):
            style_network = StyleNetwork(**style_network)
        self.style_network = style_network
        assert exists(style_network) ^ exists(style_network_dim)
        if not exists(style_network_dim):
            style_network_dim = style_network.dim
        self.style_network_dim = style_network_dim
        if isinstance(text_encoder, dict):
            text_encoder = TextEncoder(**text_encoder)
        self.text_encoder = text_encoder
        self.unconditional = unconditional
        assert not (unconditional and exists(text_encoder))
        assert not (unconditional and exists(style_network) and style_network.dim_text_latent > 0)
        assert unconditional or (exists(text_encoder) and text_encoder.dim == style_network.dim_text_latent), \
            "the `dim_text_latent` on your StyleNetwork must be equal to the `dim` set for the TextEncoder"
        assert is_power_of_two(image_size)
        num_layers = int(math.log2(image_size) - 1)
        self.num_layers = num_layers
        self.image_size = image_size

        is_adaptive = num_conv_kernels > 1
        dim_kernel_mod = num_conv_kernels if is_adaptive else 0
        split = []
        adaptive_conv = partial(AdaptiveConv2DMod, kernel=3, num_conv_kernels=num_conv_kernels)

        self.init_block = nn.Parameter(torch.randn(dim_latent, 4, 4))
        self.init_conv = adaptive_conv(dim_latent, dim_latent)
        split.extend([dim_latent, dim_kernel_mod])

        resolutions = [image_size // (2 ** (num_layers - 1 - i)) for i in range(num_layers)]
        dims = [min((2 ** (i + 1)) * dim_capacity, dim_max) for i in range(num_layers)][::-1]
        dims = [dim_latent] + dims
        dim_pairs = list(zip(dims[:-1], dims[1:]))
        self.num_skip_layers_excite = num_skip_layers_excite
        self.layers = nn.ModuleList([])
        for ind, ((dim_in, dim_out), resolution) in enumerate(zip(dim_pairs, resolutions)):
            is_last, is_first = (ind + 1) == len(dim_pairs), ind == 0
            se = None
            if num_skip_layers_excite > 0 and (ind + num_skip_layers_excite) < len(dim_pairs):
                se = SqueezeExciteParams(dim_in, dim_pairs[ind + num_skip_layers_excite][0])
            resnet_block = nn.ModuleList([adaptive_conv(dim_in, dim_out), Noise(dim_out), nn.LeakyReLU(0.2),
                                          adaptive_conv(dim_out, dim_out), Noise(dim_out), nn.LeakyReLU(0.2)])
            to_rgb = AdaptiveConv2DMod(dim_out, channels, 1, num_conv_kernels=1, demod=False)
            upsample = UpsampleParams(dim_in) if not is_first else None
            rgb_upsample = UpsampleParams(channels) if not is_last else None
            self_attn = cross_attn = None
            if resolution in self_attn_resolutions:
                self_attn = SelfAttentionBlock(dim_out, dim_head=self_attn_dim_head, heads=self_attn_heads,
                                               ff_mult=self_attn_ff_mult, dot_product=self_attn_dot_product)
            if resolution in cross_attn_resolutions and not unconditional:
                cross_attn = CrossAttentionBlock(dim_out, dim_context=text_encoder.dim, dim_head=cross_attn_dim_head,
                                                 heads=cross_attn_heads, ff_mult=cross_attn_ff_mult)
            split.extend([dim_in, dim_kernel_mod, dim_out, dim_kernel_mod, dim_out, 0])
            self.layers.append(nn.ModuleList([se, resnet_block, to_rgb, self_attn, cross_attn, upsample, rgb_upsample]))

        self.style_to_conv_modulations = nn.Linear(style_network_dim, sum(split))
        self.style_embed_split_dims = split
        self.apply(self.init_)
        nn.init.normal_(self.init_block, std=0.02)

    def init_(self, m):
        if type(m) in {nn.Conv2d, nn.Linear}:
            nn.init.kaiming_normal_(m.weight, a=0, mode="fan_in", nonlinearity="leaky_relu")

    @property
    def total_params(self):
        return sum(p.numel() for p in self.parameters() if p.requires_grad)

    @property
    def device(self):
        return next(self.parameters()).device

    def encode_text(self, texts=None, text_encodings=None, global_text_tokens=None, fine_text_tokens=None, text_mask=None):
        """-> (global, fine, mask) or (None, None, None) for the unconditional generator (ref :1156-1170)."""
        if self.unconditional:
            assert not any(map(exists, (texts, text_encodings, global_text_tokens, fine_text_tokens)))
            return None, None, None
        if exists(texts) or exists(text_encodings):
            assert exists(texts) ^ exists(text_encodings)
            return self.text_encoder(texts=texts, text_encodings=text_encodings)
        assert all(map(exists, (global_text_tokens, fine_text_tokens, text_mask))), \
            "raw text or text embeddings were not passed in for conditional training"
        return global_text_tokens, fine_text_tokens, text_mask

    def forward_nhwc(self, styles=None, noise=None, batch_size=1, layer_noises=None, global_text_tokens=None,
                     fine_text_tokens=None, text_mask=None):
        """-> (rgb NHWC, [rgbs NHWC]) in the compute dtype.  Per-layer noise images are drawn with torch.randn in
        the reference's order (ref :938) unless ``layer_noises`` supplies them."""
        dt = compute_dtype()
        if not exists(styles):
            assert exists(self.style_network)
            if not exists(noise):
                noise = torch.randn((batch_size, self.style_network_dim), device=self.device)
            styles = self.style_network(noise, global_text_tokens)
        mods = ops.linear(styles.float(), self.style_to_conv_modulations.weight, self.style_to_conv_modulations.bias)
        mods = iter(mods.split(self.style_embed_split_dims, dim=-1))
        b = styles.shape[0]
        x = self.init_block.permute(1, 2, 0)[None].expand(b, -1, -1, -1).to(dt).contiguous()
        x = self.init_conv.forward_nhwc(x, next(mods), next(mods))
        rgb = None
        excitations = [None] * self.num_skip_layers_excite
        rgbs = []
        ln = list(layer_noises) if layer_noises is not None else None

        def noise_img(t):
            if ln is not None:
                return ln.pop(0)
            return torch.randn(t.shape[0], 1, t.shape[1], t.shape[2], device=t.device)

        for se, (conv1, noise1, _, conv2, noise2, _), to_rgb, self_attn, cross_attn, upsample, upsample_rgb in self.layers:
            if exists(upsample):
                x = ops.upsample2x_blur(x)
            if exists(se):
                excitations.append(squeeze_excite(se, x))
            ex = excitations.pop(0) if excitations else None
            if exists(ex):
                x = apply_excite(x, ex)
            x = conv1.forward_nhwc(x, next(mods), next(mods))
            x = ops.NoiseActFn.apply(x, noise_img(x), noise1.weight)
            x = conv2.forward_nhwc(x, next(mods), next(mods))
            x = ops.NoiseActFn.apply(x, noise_img(x), noise2.weight)
            if exists(self_attn):
                x = self_attn.forward_nhwc(x)
            if exists(cross_attn):
                x = cross_attn.forward_nhwc(x, fine_text_tokens, text_mask)
            layer_rgb = to_rgb.forward_nhwc(x, next(mods), next(mods), out_pad=img_cpad(self.channels))
            rgb = layer_rgb if rgb is None else ops.add(rgb, layer_rgb)
            rgbs.append(rgb)
            if exists(upsample_rgb):
                rgb = ops.upsample2x_blur(rgb)
        assert next(mods, None) is None, "convolutions were incorrectly modulated"
        return rgb, rgbs

    def forward(self, styles=None, noise=None, texts=None, text_encodings=None, global_text_tokens=None,
                fine_text_tokens=None, text_mask=None, batch_size=1, return_all_rgbs=False):
        g, f, tm = self.encode_text(texts, text_encodings, global_text_tokens, fine_text_tokens, text_mask)
        rgb, rgbs = self.forward_nhwc(styles, noise, batch_size, None, g, f, tm)
        rgb = ops.to_nchw(rgb, self.channels)
        if return_all_rgbs:
            return rgb, [ops.to_nchw(t, self.channels) for t in rgbs]
        return rgb


# ----------------------------------------------------------------------------- discriminator (ref :1254-1838)
class SimpleDecoder(nn.Module):
    def __init__(self, dim, *, dims: Tuple[int, ...], patch_dim: int = 1, frac_patches: float = 1.,
                 dropout: float = 0.5):
        super().__init__()
        assert 0 < frac_patches <= 1.
        self.patch_dim, self.frac_patches = patch_dim, frac_patches
        self.static_sel = None           # trainer-owned device buffer int32 (B, nsel) when running under CUDA graphs
        self.dropout = nn.Dropout(dropout)
        dims = [dim, *dims]
        layers = [nn.Conv2d(dim, dim, 3, padding=1)]
        for dim_in, dim_out in zip(dims[:-1], dims[1:]):
            layers.append(nn.Sequential(UpsampleParams(dim_in), nn.Conv2d(dim_in, dim_out, 3, padding=1),
                                        nn.LeakyReLU(0.2)))
        self.net = nn.Sequential(*layers)

    def draw_patch_indices(self, b):
        """int32 (b, nsel): indices p1*patch_dim+p2 of the randomly kept patches; CPU randn argsort like the reference
        (:1310), same host RNG consumption."""
        total = self.patch_dim ** 2
        nsel = max(int(self.frac_patches * total), 1)
        perm = torch.randn((b, total)).sort(dim=-1).indices[:, :nsel]
        return perm.to(torch.int32).contiguous()

    def forward_nhwc(self, fmap, image_nhwc):
        """fmap (B,h,w,C) compute dtype; image_nhwc (B,H,W,3).  RNG draws mirror ref :1295,:1310 (dropout on the
        NCHW-shaped map on device, patch permutation from a CPU randn)."""
        b, h, w, c = fmap.shape
        if self.training and self.dropout.p > 0:
            keep = F.dropout(torch.ones((b, c, h, w), device=fmap.device), self.dropout.p, True)
            fmap = ops.mul(fmap, ops.to_nhwc(keep, c, fmap.dtype))
        if self.frac_patches < 1.:
            pd = self.patch_dim
            total = pd * pd
            nsel = max(int(self.frac_patches * total), 1)
            if self.static_sel is not None:
                sel = self.static_sel
            else:
                sel = self.draw_patch_indices(b).to(fmap.device, non_blocking=True)

            def pick(t):   # '(b p) ...' selection of the kept p1 x p2 sub-blocks: one gather launch
                return ops.patch_select(t, sel, pd)

            fmap, image_nhwc = pick(fmap), pick(image_nhwc)
        x = ops.conv2d(fmap, self.net[0].weight, self.net[0].bias, pad=1)
        for blk in list(self.net)[1:]:
            x = ops.upsample2x_blur(x)
            wgt, bias = blk[1].weight, blk[1].bias
            cpad = image_nhwc.shape[-1]
            if wgt.shape[0] < cpad:          # image-like output: zero filters up to the padded channel count (tcgen05 tiles)
                wgt = F.pad(wgt, (0, 0, 0, 0, 0, 0, 0, cpad - wgt.shape[0]))
                bias = F.pad(bias, (0, cpad - bias.shape[0]))
            x = ops.conv2d(x, wgt, bias, pad=1, act=1)
        d = ops.axpby(1.0, x, -1.0, image_nhwc)                  # padded channels are 0 - 0
        real_numel = d.numel() // d.shape[-1] * list(self.net)[-1][1].weight.shape[0]
        return ops.axpby(1.0 / real_numel, ops.sum_all(ops.mul(d, d)))


class Predictor(nn.Module):
    def __init__(self, dim, depth=4, num_conv_kernels=2, unconditional=False):
        super().__init__()
        self.unconditional = unconditional
        self.residual_fn = nn.Conv2d(dim, dim, 1)
        self.residual_scale = 2 ** -0.5
        self.layers = nn.ModuleList([])
        for _ in range(depth):
            if unconditional:
                c1, c2 = nn.Conv2d(dim, dim, 3, padding=1), nn.Conv2d(dim, dim, 3, padding=1)
                self.layers.append(nn.ModuleList([c1, nn.LeakyReLU(0.2), c2, nn.LeakyReLU(0.2)]))
            else:      # text-conditioned: per-sample modulated filters (ref :1459)
                c1 = AdaptiveConv2DMod(dim, dim, 3, num_conv_kernels=num_conv_kernels)
                l1 = nn.LeakyReLU(0.2)
                c2 = AdaptiveConv2DMod(dim, dim, 3, num_conv_kernels=num_conv_kernels)
                self.layers.append(nn.ModuleList([c1, l1, c2, nn.LeakyReLU(0.2)]))
        self.to_logits = nn.Conv2d(dim, 1, 1)

    def forward_nhwc(self, x, mod=None, kernel_mod=None, fused=None):
        residual = ops.conv2d(x, self.residual_fn.weight, self.residual_fn.bias)
        for conv1, _, conv2, _ in self.layers:
            inner = x
            if self.unconditional:
                x = ops.conv2d(x, conv1.weight, conv1.bias, pad=1, act=1)
                x = ops.conv2d(x, conv2.weight, conv2.bias, pad=1, act=1)
            else:
                x = ops.leaky_relu(conv1.forward_nhwc(x, mod, kernel_mod, fused=fused))
                x = ops.leaky_relu(conv2.forward_nhwc(x, mod, kernel_mod, fused=fused))
            x = ops.axpby(self.residual_scale, x, self.residual_scale, inner)
        x = ops.add(x, residual)
        n, h, w, c = x.shape
        fz = _COMPUTE["fused_attention"] if fused is None else fused
        y = ops.linear_rows(x.reshape(n * h * w, c), self.to_logits.weight.reshape(1, c), self.to_logits.bias, fused=fz,
                            w_param=self.to_logits.weight)
        return y.reshape(n, h, w, 1)


def DownsampleParams(dim):
    return nn.Sequential(nn.Identity(), nn.Conv2d(dim * 4, dim, 1))


class Discriminator(nn.Module):
    def __init__(self, *, dim_capacity=16, image_size, dim_max=2048, channels=3,
                 attn_resolutions: Tuple[int, ...] = (32, 16), attn_dim_head=64, attn_heads=8,
                 self_attn_dot_product=False, ff_mult=4, text_encoder=None, text_dim=None,
                 filter_input_resolutions: bool = True,
                 multiscale_input_resolutions: Tuple[int, ...] = (64, 32, 16, 8),
                 multiscale_output_skip_stages: int = 1, aux_recon_resolutions: Tuple[int, ...] = (8,),
                 aux_recon_patch_dims: Tuple[int, ...] = (2,), aux_recon_frac_patches: Tuple[float, ...] = (0.25,),
                 aux_recon_fmap_dropout: float = 0.5, resize_mode="bilinear", num_conv_kernels=2,
                 num_skip_layers_excite=0, unconditional=False, predictor_depth=2):
        super().__init__()
        assert resize_mode == "bilinear"
        self.unconditional = unconditional
        assert not (unconditional and exists(text_encoder))
        self.channels = channels
        assert is_power_of_two(image_size)
        if filter_input_resolutions:
            multiscale_input_resolutions = [r for r in multiscale_input_resolutions if r < image_size]
        assert len(set(multiscale_input_resolutions)) == len(multiscale_input_resolutions)
        assert all(is_power_of_two(r) and r < image_size for r in multiscale_input_resolutions)
        self.multiscale_input_resolutions = list(multiscale_input_resolutions)
        assert multiscale_output_skip_stages > 0
        ms_out = [r // (2 ** multiscale_output_skip_stages) for r in multiscale_input_resolutions]
        assert all(r >= 4 for r in ms_out)
        self.multiscale_output_resolutions = ms_out
        assert len(aux_recon_resolutions) == len(aux_recon_patch_dims) == len(aux_recon_frac_patches)
        self.aux_recon_resolutions_to_patches = dict(zip(aux_recon_resolutions,
                                                         zip(aux_recon_patch_dims, aux_recon_frac_patches)))
        self.resize_mode = resize_mode
        num_layers = int(math.log2(image_size) - 1)
        self.num_layers, self.image_size = num_layers, image_size
        resolutions = [image_size // (2 ** i) for i in range(num_layers)]
        dims = [min(d, dim_max) for d in [channels] + [(2 ** (i + 1)) * dim_capacity for i in range(num_layers)]]
        dim_last = dims[-1]
        dim_pairs = list(zip(dims[:-1], dims[1:]))
        self.num_skip_layers_excite = num_skip_layers_excite
        self.residual_scale = 2 ** -0.5
        self.layers = nn.ModuleList([])
        upsample_dims = []
        predictor_dims = []
        dim_kernel_attn = num_conv_kernels if num_conv_kernels > 1 else 0
        for ind, ((dim_in, dim_out), resolution) in enumerate(zip(dim_pairs, resolutions)):
            is_first, is_last = ind == 0, (ind + 1) == len(dim_pairs)
            should_downsample = not is_last
            upsample_dims.insert(0, dim_in)
            se = None
            if not is_first and num_skip_layers_excite > 0 and (ind + num_skip_layers_excite) < len(dim_pairs):
                se = SqueezeExciteParams(dim_in, dim_pairs[ind + num_skip_layers_excite][0])
            from_rgb = nn.Conv2d(channels, dim_in, 7, padding=3)
            residual_conv = nn.Conv2d(dim_in, dim_out, 1, stride=(2 if should_downsample else 1))
            resnet_block = nn.Sequential(nn.Conv2d(dim_in, dim_out, 3, padding=1), nn.LeakyReLU(0.2),
                                         nn.Conv2d(dim_out, dim_out, 3, padding=1), nn.LeakyReLU(0.2))
            predictor = None
            if resolution in ms_out:
                predictor = Predictor(dim_out, num_conv_kernels=num_conv_kernels, depth=2, unconditional=unconditional)
                predictor_dims.extend([dim_out, dim_kernel_attn])
            decoder = None
            if resolution in aux_recon_resolutions:
                patch_dim, frac = self.aux_recon_resolutions_to_patches[resolution]
                decoder = SimpleDecoder(dim_out, dims=tuple(upsample_dims), patch_dim=patch_dim, frac_patches=frac,
                                        dropout=aux_recon_fmap_dropout)
            attn = None
            if resolution in attn_resolutions:
                attn = SelfAttentionBlock(dim_out, heads=attn_heads, dim_head=attn_dim_head, ff_mult=ff_mult,
                                          dot_product=self_attn_dot_product)
            self.layers.append(nn.ModuleList([se, from_rgb, resnet_block, residual_conv, attn, predictor, decoder,
                                              DownsampleParams(dim_out) if should_downsample else None]))
        self.to_logits = nn.Sequential(nn.Conv2d(dim_last, dim_last, 3, padding=1), nn.Identity(),
                                       nn.Linear(dim_last * (4 ** 2), 1), nn.Identity())
        assert unconditional or (exists(text_dim) ^ exists(text_encoder))
        if not unconditional:
            if isinstance(text_encoder, dict):
                text_encoder = TextEncoder(**text_encoder)
            self.text_dim = text_dim if exists(text_dim) else text_encoder.dim
            self.predictor_dims = predictor_dims
            self.text_to_conv_conditioning = nn.Linear(self.text_dim, sum(predictor_dims))
        self.text_encoder = text_encoder
        self.apply(self.init_)

    def init_(self, m):
        if type(m) in {nn.Conv2d, nn.Linear}:
            nn.init.kaiming_normal_(m.weight, a=0, mode="fan_in", nonlinearity="leaky_relu")

    @property
    def total_params(self):
        return sum(p.numel() for p in self.parameters())

    @property
    def device(self):
        return next(self.parameters()).device

    # -- reference :1683-1687: bilinear F.interpolate of NCHW images
    def resize_image_to(self, images, resolution):
        x = ops.to_nhwc(images, images.shape[1], torch.float32)
        return ops.to_nchw(ops.resize_bilinear(x, resolution), images.shape[1])

    def real_images_to_rgbs(self, images):
        return [self.resize_image_to(images, r) for r in self.multiscale_input_resolutions]

    def real_images_to_rgbs_nhwc(self, images_nhwc):
        return [ops.resize_bilinear(images_nhwc, r) for r in self.multiscale_input_resolutions]

    def encode_text(self, texts=None, text_encodings=None, text_embeds=None):
        """-> text_embeds (b, text_dim) for the predictors, None when unconditional (ref :1709-1726)."""
        if self.unconditional:
            assert not any(map(exists, (texts, text_embeds)))
            return None
        assert (exists(texts) ^ exists(text_encodings)) ^ exists(text_embeds)
        if exists(texts) or exists(text_encodings):
            assert exists(self.text_encoder)
            text_embeds, *_ = self.text_encoder(texts=texts, text_encodings=text_encodings)
        return text_embeds

    def forward_nhwc(self, images, rgbs: List[torch.Tensor], return_multiscale_outputs=True, calc_aux_loss=True,
                     fused_attention=None, text_embeds=None, aux_batch=None):
        """images (B,S,S,C) NHWC compute dtype; rgbs: NHWC maps.  -> (logits (s,B) fp32, [ms logits NHWC], [aux]).
        ``aux_batch``: the auxiliary reconstruction (ref :1812-1827) only sees the first aux_batch images (the trainer
        passes real and fake images as one batch, real rows first)."""
        x = images
        batch = x.shape[0]
        aux_batch = batch if aux_batch is None else aux_batch
        assert x.shape[1] == x.shape[2] == self.image_size
        by_res = {t.shape[2]: t for t in rgbs} if rgbs is not None else {}
        missing = set(self.multiscale_input_resolutions) - set(by_res.keys())
        assert not missing, f"rgbs of necessary resolution {self.multiscale_input_resolutions} were not passed in"
        ms_outputs, aux_losses = [], []
        conv_mods = None
        if not self.unconditional:
            assert exists(text_embeds), "text embeddings were not passed into the discriminator"
            cm = ops.linear(text_embeds.float(), self.text_to_conv_conditioning.weight, self.text_to_conv_conditioning.bias)
            conv_mods = iter(cm.split(self.predictor_dims, dim=-1))
        excitations = [None] * (self.num_skip_layers_excite + 1)
        for se, from_rgb, block, residual_fn, attn, predictor, decoder, downsample in self.layers:
            resolution = x.shape[2]
            if exists(se):
                excitations.append(squeeze_excite(se, x))
            ex = excitations.pop(0) if excitations else None
            if exists(ex):
                x = apply_excite(x, ex)
            prev = x.shape[0]
            if resolution in self.multiscale_input_resolutions:
                rgb = by_res[resolution]
                f = ops.conv2d(rgb, from_rgb.weight, from_rgb.bias, pad=3)
                if x.shape[0] != f.shape[0]:
                    f = f.repeat(x.shape[0] // f.shape[0], 1, 1, 1)
                x = torch.cat((ops.add(x, f), f), dim=0)
            stride = residual_fn.stride[0]
            residual = ops.conv2d(x, residual_fn.weight, residual_fn.bias, stride=stride)
            x = ops.conv2d(x, block[0].weight, block[0].bias, pad=1, act=1)
            x = ops.conv2d(x, block[2].weight, block[2].bias, pad=1, act=1)
            if exists(attn):
                x = attn.forward_nhwc(x, fused=fused_attention)
            if exists(predictor):
                pk = dict(mod=next(conv_mods), kernel_mod=next(conv_mods)) if exists(conv_mods) else {}
                if return_multiscale_outputs:
                    ms_outputs.append(predictor.forward_nhwc(x[:prev], fused=fused_attention, **pk))
            if exists(downsample):     # pixel-unshuffle + 1x1 == 2x2 stride-2 conv (ref :289-293)
                w = downsample[1].weight
                w2 = w.view(w.shape[0], w.shape[1] // 4, 2, 2)
                x = ops.conv2d(x, w2, downsample[1].bias, stride=2, res=residual, gain=self.residual_scale)
            else:
                x = ops.axpby(self.residual_scale, x, self.residual_scale, residual)
            if exists(decoder) and calc_aux_loss:                 # ref :1812-1827 (post-downsample x, first B rows)
                aux_losses.append(decoder.forward_nhwc(x[:aux_batch], images[:aux_batch]))   # zero pad channels kept
        x = ops.conv2d(x, self.to_logits[0].weight, self.to_logits[0].bias, pad=1)
        lw = self.to_logits[2].weight                                                   # (1, c*h*w) in (c h w) order
        c = x.shape[-1]
        lw = lw.view(1, c, 4, 4).permute(0, 2, 3, 1).reshape(1, 16 * c)                  # -> (h w c) like NHWC rows
        fz = _COMPUTE["fused_attention"] if fused_attention is None else fused_attention
        logits = ops.linear_rows(x.reshape(x.shape[0], 16 * c), lw, self.to_logits[2].bias, fused=fz)
        logits = logits.float().reshape(-1, batch)
        return logits, ms_outputs, aux_losses

    def forward(self, images, rgbs: List[torch.Tensor], texts=None, text_encodings=None, text_embeds=None,
                real_images=None, return_multiscale_outputs=True, calc_aux_loss=True):
        te = self.encode_text(texts, text_encodings, text_embeds)
        dt = compute_dtype()
        x = ops.to_nhwc(images, img_cpad(self.channels), dt)
        r = [ops.to_nhwc(t, img_cpad(self.channels), dt) for t in rgbs]
        logits, ms, aux = self.forward_nhwc(x, r, return_multiscale_outputs, calc_aux_loss, text_embeds=te)
        return logits, [ops.to_nchw(m, 1) for m in ms], aux


# ----------------------------------------------------------------------------- attend.py (ref attend.py:34-110)
class Attend(nn.Module):
    """softmax(q k^T / sqrt(d)) v for (b, h, n, d) tensors — the reference's ``Attend`` (plain or "flash" SDPA path;
    both compute the same function).  Runs the fused attention kernels (tcgen05 for bf16, d = 64, n % 128 == 0)."""

    def __init__(self, dropout=0., flash=False):
        super().__init__()
        assert dropout == 0., "attention dropout is not used on the training hot path (reference default 0.)"
        self.dropout, self.flash = dropout, flash

    def forward(self, q, k, v):
        b, h, n, d = q.shape
        dt = compute_dtype()

        def rows(t):      # (b, h, n, d) -> (b, n, h*d) rows, the layout the kernels read
            return t.to(dt).permute(0, 2, 1, 3).reshape(t.shape[0], t.shape[2], h * d).contiguous()

        o = ops.fused_attention(rows(q), rows(k), rows(v), None, h, d ** -0.5, l2=False)
        return o.view(b, n, h, d).permute(0, 2, 1, 3).to(q.dtype)


# ============================================================================= UnetUpsampler (ref unet_upsampler.py)
def _rmsnorm_vec(x, gamma):
    """unet RMSNorm (ref unet_upsampler.py:224-234): gamma has shape (C,)."""
    return channel_rmsnorm(x, gamma.reshape(-1, 1, 1))


class RMSNorm2d(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.scale = dim ** 0.5
        self.gamma = nn.Parameter(torch.ones(dim))

    def forward_nhwc(self, x):
        return _rmsnorm_vec(x, self.gamma)


class UnetBlock(nn.Module):
    """AdaptiveConv2DMod -> RMSNorm -> SiLU (ref :238-270)"""

    def __init__(self, dim, dim_out, num_conv_kernels=0):
        super().__init__()
        self.proj = AdaptiveConv2DMod(dim, dim_out, kernel=3, num_conv_kernels=num_conv_kernels)
        self.norm = RMSNorm2d(dim_out)
        self.act = nn.SiLU()

    def forward_nhwc(self, x, mods):
        x = self.proj.forward_nhwc(x, next(mods), next(mods))
        return ops.unary(U_SILU, self.norm.forward_nhwc(x))


class UnetResnetBlock(nn.Module):
    def __init__(self, dim, dim_out, *, num_conv_kernels=0, style_dims=None):
        super().__init__()
        mod_dims = [dim, num_conv_kernels, dim_out, num_conv_kernels]
        style_dims.extend(mod_dims)
        self.num_mods = len(mod_dims)
        self.block1 = UnetBlock(dim, dim_out, num_conv_kernels)
        self.block2 = UnetBlock(dim_out, dim_out, num_conv_kernels)
        self.res_conv = nn.Conv2d(dim, dim_out, 1) if dim != dim_out else nn.Identity()

    def forward_nhwc(self, x, mods):
        h = self.block2.forward_nhwc(self.block1.forward_nhwc(x, mods), mods)
        if isinstance(self.res_conv, nn.Identity):
            return ops.add(h, x)
        return ops.conv2d(x, self.res_conv.weight, self.res_conv.bias, res=h)


class UnetLinearAttention(nn.Module):
    """ref :312-349: softmax over d on q, over tokens on k, context = k v^T (d x d per head), out = context^T q."""

    def __init__(self, dim, heads=4, dim_head=32):
        super().__init__()
        self.scale, self.heads, self.dim_head = dim_head ** -0.5, heads, dim_head
        hidden = dim_head * heads
        self.norm = RMSNorm2d(dim)
        self.to_qkv = nn.Conv2d(dim, hidden * 3, 1, bias=False)
        self.to_out = nn.Sequential(nn.Conv2d(hidden, dim, 1), RMSNorm2d(dim))

    def forward_nhwc(self, x):
        b, hh, ww, _ = x.shape
        n, H, d = hh * ww, self.heads, self.dim_head
        qkv = ops.conv2d(self.norm.forward_nhwc(x), self.to_qkv.weight)            # (b,h,w,3*H*d)
        q, k, v = (qkv[..., i * H * d:(i + 1) * H * d] for i in range(3))
        q = ops.axpby(self.scale, ops.softmax(q.reshape(b * n * H, d)))             # softmax over d per (token, head)
        k = ops.softmax_tokens(k.reshape(b, n, H * d))                              # softmax over tokens per channel
        q4 = q.view(b, n, H, d).permute(0, 2, 1, 3)                                 # (b,H,n,d)
        kt = k.view(b, n, H, d).permute(0, 2, 3, 1)                                 # (b,H,d,n)
        v4 = v.reshape(b, n, H, d).permute(0, 2, 1, 3)                              # (b,H,n,e)
        context = ops.bmm(kt, v4)                                                   # (b,H,d,e)
        out = ops.bmm(q4, context, out_bmhn=True).permute(0, 2, 1, 3).reshape(b, hh, ww, H * d)
        out = ops.conv2d(out, self.to_out[0].weight, self.to_out[0].bias)
        return self.to_out[1].forward_nhwc(out)


class UnetAttention(nn.Module):
    """ref :351-380: full attention through Attend (no null key, dot product)."""

    def __init__(self, dim, heads=4, dim_head=32, flash=False):
        super().__init__()
        self.heads, self.dim_head = heads, dim_head
        hidden = dim_head * heads
        self.norm = RMSNorm2d(dim)
        self.attend = Attend(flash=flash)
        self.to_qkv = nn.Conv2d(dim, hidden * 3, 1, bias=False)
        self.to_out = nn.Conv2d(hidden, dim, 1)

    def forward_nhwc(self, x):
        b, hh, ww, _ = x.shape
        n, H, d = hh * ww, self.heads, self.dim_head
        qkv = ops.conv2d(self.norm.forward_nhwc(x), self.to_qkv.weight).view(b, n, 3 * H * d)
        q, k, v = (qkv[..., i * H * d:(i + 1) * H * d] for i in range(3))            # strided row views
        o = ops.fused_attention(q, k, v, None, H, d ** -0.5, l2=False)
        return ops.conv2d(o.view(b, hh, ww, H * d), self.to_out.weight, self.to_out.bias)


def UnetFeedForward(dim, mult=4):
    return nn.Sequential(RMSNorm2d(dim), nn.Conv2d(dim, dim * mult, 1), nn.GELU(), nn.Conv2d(dim * mult, dim, 1))


class UnetTransformer(nn.Module):
    def __init__(self, dim, dim_head=64, heads=8, depth=1, flash_attn=True, ff_mult=4, linear=False):
        super().__init__()
        self.layers = nn.ModuleList([])
        for _ in range(depth):
            attn = (UnetLinearAttention(dim=dim, dim_head=dim_head, heads=heads) if linear
                    else UnetAttention(dim=dim, dim_head=dim_head, heads=heads, flash=flash_attn))
            self.layers.append(nn.ModuleList([attn, UnetFeedForward(dim=dim, mult=ff_mult)]))

    def forward_nhwc(self, x):
        for attn, ff in self.layers:
            x = ops.add(attn.forward_nhwc(x), x)
            h = ops.conv2d(ff[0].forward_nhwc(x), ff[1].weight, ff[1].bias)
            x = ops.conv2d(ops.unary(U_GELU, h), ff[3].weight, ff[3].bias, res=x)
        return x


class UnetDownsample(nn.Module):
    """conv3x3, high-frequency skip x - blur(x), 2x2 max-pool (ref :82-160)."""

    def __init__(self, dim, dim_out=None, skip_downsample=False, has_temporal_layers=False):
        super().__init__()
        assert not has_temporal_layers
        self.skip_downsample = skip_downsample
        self.conv2d = nn.Conv2d(dim, dim_out if dim_out is not None else dim, 3, padding=1)
        self.register_buffer("filter", torch.Tensor([1., 2., 1.]))

    def forward_nhwc(self, x):
        x = ops.conv2d(x, self.conv2d.weight, self.conv2d.bias, pad=1)
        if self.skip_downsample:
            return x, None
        blurred = ops.ResampleFn.apply(x, ops.get_resample_op("blur", x.shape[1], x.shape[2], x.device), False)
        return ops.maxpool2(x), ops.axpby(1.0, x, -1.0, blurred)


class PixelShuffleUpsample(nn.Module):
    """conv1x1 -> SiLU -> pixel shuffle x2 (ref gigagan_pytorch.py:263-287)."""

    def __init__(self, dim, dim_out=None):
        super().__init__()
        dim_out = dim_out if dim_out is not None else dim
        conv = nn.Conv2d(dim, dim_out * 4, 1)
        self.net = nn.Sequential(conv, nn.SiLU(), nn.PixelShuffle(2))
        o, i, h, w = conv.weight.shape
        w0 = torch.empty(o // 4, i, h, w)
        nn.init.kaiming_uniform_(w0)
        conv.weight.data.copy_(w0.repeat_interleave(4, dim=0))
        nn.init.zeros_(conv.bias.data)

    def forward_nhwc(self, x):
        y = ops.unary(U_SILU, ops.conv2d(x, self.net[0].weight, self.net[0].bias))
        b, h, w, c4 = y.shape
        c = c4 // 4                                   # channel index = c*4 + r1*2 + r2
        return y.view(b, h, w, c, 2, 2).permute(0, 1, 4, 2, 5, 3).reshape(b, 2 * h, 2 * w, c)


class UnetUpsampler(BaseGenerator):
    def __init__(self, dim, *, image_size, input_image_size, init_dim=None, out_dim=None, text_encoder=None,
                 style_network: StyleNetwork | Dict | None = None, style_network_dim=None,
                 dim_mults=(1, 2, 4, 8, 16), channels=3, full_attn=(False, False, False, True, True),
                 cross_attn=(False, False, False, True, True), flash_attn=True, self_attn_dim_head=64,
                 self_attn_heads=8, self_attn_dot_product=True, self_attn_ff_mult=4, attn_depths=(1, 1, 1, 1, 1),
                 temporal_attn_depths=(1, 1, 1, 1, 1), cross_attn_dim_head=64, cross_attn_heads=8, cross_ff_mult=4,
                 has_temporal_layers=False, mid_attn_depth=1, num_conv_kernels=2, unconditional=True,
                 skip_connect_scale=None):
        super().__init__()
        assert unconditional and text_encoder is None and not has_temporal_layers, \
            "this build covers the unconditional image upsampler (text / video layers: DESIGN.md section 6)"
        self.can_upsample_video = False
        self.text_encoder = None
        if isinstance(style_network, dict):
            style_network = StyleNetwork(**style_network)
        self.style_network = style_network
        assert exists(style_network) ^ exists(style_network_dim)
        self.unconditional = unconditional
        assert is_power_of_two(image_size) and is_power_of_two(input_image_size) and input_image_size < image_size
        n_no_down = int(math.log2(image_size) - math.log2(input_image_size))
        assert n_no_down <= len(dim_mults)
        self.image_size, self.input_image_size, self.channels = image_size, input_image_size, channels
        split = []
        init_dim = init_dim if init_dim is not None else dim
        self.init_conv = nn.Conv2d(channels, init_dim, 7, padding=3)
        dims = [init_dim, *[dim * m for m in dim_mults]]
        mid_dim = dims[-1]
        in_out = list(zip(dims[:-1], dims[1:]))
        block = partial(UnetResnetBlock, num_conv_kernels=num_conv_kernels, style_dims=split)
        full_attn = full_attn if isinstance(full_attn, tuple) else (full_attn,) * len(dim_mults)
        cross_attn = cross_attn if isinstance(cross_attn, tuple) else (cross_attn,) * len(dim_mults)
        self.skip_connect_scale = skip_connect_scale if skip_connect_scale is not None else 2 ** -0.5

        def attn_klass(full, *a, **k):
            return UnetTransformer(*a, flash_attn=flash_attn, linear=not full, **k)

        self.downs, self.ups = nn.ModuleList([]), nn.ModuleList([])
        skip_dims = []
        for ind, ((dim_in, dim_out), layer_full, layer_depth) in enumerate(zip(in_out, full_attn, attn_depths)):
            no_down = ind < n_no_down
            skip_dims.append(dim_in)
            skip_dims.append(dim_in + (dim_out if not no_down else 0))
            self.downs.append(nn.ModuleList([
                block(dim_in, dim_in), block(dim_in, dim_in), None,
                attn_klass(layer_full, dim_in, dim_head=self_attn_dim_head, heads=self_attn_heads, depth=layer_depth),
                None, None, UnetDownsample(dim_in, dim_out, skip_downsample=no_down)]))
        self.mid_block1 = block(mid_dim, mid_dim)
        self.mid_attn = attn_klass(True, mid_dim, dim_head=self_attn_dim_head, heads=self_attn_heads, depth=mid_attn_depth)
        self.mid_block2 = block(mid_dim, mid_dim)
        self.mid_to_rgb = nn.Conv2d(mid_dim, channels, 1)
        # ref :596 zips reversed(full_attn) under the name layer_cross_attn and reversed(cross_attn) under
        # layer_full_attn: the attention type of the `ups` therefore follows cross_attn reversed.
        for (dim_in, dim_out), layer_full, layer_depth in zip(reversed(in_out), reversed(cross_attn), reversed(attn_depths)):
            self.ups.append(nn.ModuleList([
                PixelShuffleUpsample(dim_out, dim_in), UpsampleParams(), None, None, nn.Conv2d(dim_in, channels, 1),
                block(dim_in + skip_dims.pop(), dim_in), block(dim_in + skip_dims.pop(), dim_in), None,
                attn_klass(layer_full, dim_in, dim_head=cross_attn_dim_head, heads=self_attn_heads, depth=layer_depth),
                None, None]))
        self.out_dim = out_dim if out_dim is not None else channels
        self.final_res_block = block(dim, dim)
        self.final_to_rgb = nn.Conv2d(dim, channels, 1)
        self.style_to_conv_modulations = nn.Linear(style_network.dim, sum(split))
        self.style_embed_split_dims = split

    @property
    def allowable_rgb_resolutions(self):
        a, b = int(math.log2(self.input_image_size)), int(math.log2(self.image_size))
        return [2 ** p for p in range(a, b)]

    @property
    def device(self):
        return next(self.parameters()).device

    @property
    def total_params(self):
        return sum(p.numel() for p in self.parameters())

    def forward_nhwc(self, lowres, styles=None, noise=None):
        """lowres (B,S,S,C) NHWC compute dtype -> (rgb NHWC, [rgbs NHWC incl. the low-res input first])."""
        b = lowres.shape[0]
        if not exists(styles):
            if not exists(noise):
                noise = torch.randn((b, self.style_network.dim), device=self.device)
            styles = self.style_network(noise)
        mods = ops.linear(styles.float(), self.style_to_conv_modulations.weight, self.style_to_conv_modulations.bias)
        mods = iter(mods.split(self.style_embed_split_dims, dim=-1))
        cp = lowres.shape[-1]
        x = ops.conv2d(lowres, self.init_conv.weight, self.init_conv.bias, pad=3)
        hs = []
        for block1, block2, _, attn, _, _, down in self.downs:
            x = block1.forward_nhwc(x, mods)
            hs.append(x)
            x = block2.forward_nhwc(x, mods)
            x = attn.forward_nhwc(x)
            skip = x
            x, hf = down.forward_nhwc(x)
            hs.append(skip if hf is None else torch.cat((skip, hf), dim=-1))
        x = self.mid_block1.forward_nhwc(x, mods)
        x = self.mid_attn.forward_nhwc(x)
        x = self.mid_block2.forward_nhwc(x, mods)

        def to_rgb(conv, t):          # channels -> padded image channel count of the compute dtype
            w = conv.weight
            if cp > w.shape[0]:
                w = F.pad(w, (0, 0, 0, 0, 0, 0, 0, cp - w.shape[0]))
                bias = F.pad(conv.bias, (0, cp - conv.bias.shape[0]))
            else:
                bias = conv.bias
            return ops.conv2d(t, w, bias)

        rgb = to_rgb(self.mid_to_rgb, x)
        rgbs = [rgb]
        for up, _, _, _, rgb_conv, block1, block2, _, attn, _, _ in self.ups:
            x = up.forward_nhwc(x)
            rgb = ops.upsample2x_blur(rgb)
            r1 = ops.axpby(self.skip_connect_scale, hs.pop())
            r2 = ops.axpby(self.skip_connect_scale, hs.pop())
            if x.shape[1:3] != r1.shape[1:3]:
                r1 = ops.resize_bilinear(r1, x.shape[1])
                r2 = ops.resize_bilinear(r2, x.shape[1])
            x = block1.forward_nhwc(torch.cat((x, r1), dim=-1), mods)
            x = block2.forward_nhwc(torch.cat((x, r2), dim=-1), mods)
            x = attn.forward_nhwc(x)
            rgb = ops.add(rgb, to_rgb(rgb_conv, x))
            rgbs.append(rgb)
        x = self.final_res_block.forward_nhwc(x, mods)
        assert next(mods, None) is None
        rgb = ops.add(rgb, to_rgb(self.final_to_rgb, x))
        rgbs = [lowres] + [t for t in rgbs if t.shape[2] > lowres.shape[2]]
        return rgb, rgbs

    def forward(self, lowres_image_or_video=None, styles=None, noise=None, texts=None, global_text_tokens=None,
                fine_text_tokens=None, text_mask=None, return_all_rgbs=False,
                replace_rgb_with_input_lowres_image=True, lowres_image=None):
        # `lowres_image=` is what GigaGAN passes (ref gigagan_pytorch.py:2212); accept both spellings
        x = lowres_image_or_video if lowres_image_or_video is not None else lowres_image
        assert x is not None and x.ndim == 4 and x.shape[-2:] == (self.input_image_size,) * 2
        assert not any(map(exists, (texts, global_text_tokens, fine_text_tokens)))
        xn = ops.to_nhwc(x, img_cpad(self.channels), compute_dtype())
        rgb, rgbs = self.forward_nhwc(xn, styles, noise)
        rgb = ops.to_nchw(rgb, self.channels)
        if not return_all_rgbs:
            return rgb
        return rgb, [ops.to_nchw(t, self.channels) for t in rgbs]


# ============================================================================= text conditioning (ref :234-242, :596-867)
NEG_MASK = -1e30          # additive logit for masked / padded keys (the reference fills -finfo.max; both give p = 0)


class RMSNorm(nn.Module):
    """last-axis RMSNorm for token sequences (ref :234-242); fp32 rows."""

    def __init__(self, dim):
        super().__init__()
        self.scale = dim ** 0.5
        self.gamma = nn.Parameter(torch.ones(dim))

    def forward_rows(self, x2d):
        inv = ops.unary(U_INVNORM, ops.rowdot(x2d, x2d))
        g = ops.axpby(self.scale, self.gamma.reshape(1, -1))
        return ops.scale_channels(ops.scale_rows(x2d, inv), g, x2d.shape[0], 1)


def _masked_attention(q, k, v, bias, heads, scale, rows_per_sample_factor=1):
    """q (b,nq,h*d), k/v (b,L,h*d) same dtype, bias fp32 (b,L) added to every query row of sample b -> (b,nq,h*d).
    Composed from the closed primitives (works for fp32 token sequences and bf16 feature maps alike)."""
    b, nq, hd = q.shape
    L, d = k.shape[1], hd // heads
    q4 = q.reshape(b, nq, heads, d).permute(0, 2, 1, 3)
    kt = k.reshape(b, L, heads, d).permute(0, 2, 3, 1)
    v4 = v.reshape(b, L, heads, d).permute(0, 2, 1, 3)
    s = ops.bmm(q4, kt, alpha=scale)                                   # (b,h,nq,L)
    p = ops.softmax(s, bias, heads * nq, b) if bias is not None else ops.softmax(s)
    o = ops.bmm(p, v4, out_bmhn=True)                                  # physical (b,nq,h,d)
    return o.permute(0, 2, 1, 3).reshape(b, nq, hd)


class TextAttention(nn.Module):
    def __init__(self, dim, dim_head=64, heads=8):
        super().__init__()
        self.heads, self.dim_head, self.scale = heads, dim_head, dim_head ** -0.5
        inner = dim_head * heads
        self.norm = RMSNorm(dim)
        self.to_qkv = nn.Linear(dim, inner * 3, bias=False)
        self.null_kv = nn.Parameter(torch.randn(2, heads, dim_head))
        self.to_out = nn.Linear(inner, dim, bias=False)

    def forward_tokens(self, x, mask=None):
        b, n, dim = x.shape
        inner = self.heads * self.dim_head
        qkv = ops.linear(self.norm.forward_rows(x.reshape(b * n, dim)), self.to_qkv.weight).view(b, n, 3 * inner)
        q, k, v = (qkv[..., i * inner:(i + 1) * inner] for i in range(3))
        nk = self.null_kv[0].reshape(1, 1, inner).expand(b, 1, inner)
        nv = self.null_kv[1].reshape(1, 1, inner).expand(b, 1, inner)
        k, v = torch.cat((nk, k), dim=1), torch.cat((nv, v), dim=1)
        bias = None
        if exists(mask):
            m = F.pad(mask, (1, 0), value=True)
            bias = torch.zeros(m.shape, dtype=torch.float32, device=x.device).masked_fill_(~m, NEG_MASK)
        o = _masked_attention(q.contiguous(), k, v, bias, self.heads, self.scale)
        return ops.linear(o.reshape(b * n, inner), self.to_out.weight).view(b, n, dim)


def TokenFeedForward(dim, mult=4):
    return nn.Sequential(RMSNorm(dim), nn.Linear(dim, int(dim * mult)), nn.GELU(), nn.Linear(int(dim * mult), dim))


class Transformer(nn.Module):
    def __init__(self, dim, depth, dim_head=64, heads=8, ff_mult=4):
        super().__init__()
        self.layers = nn.ModuleList([])
        for _ in range(depth):
            self.layers.append(nn.ModuleList([TextAttention(dim=dim, dim_head=dim_head, heads=heads),
                                              TokenFeedForward(dim=dim, mult=ff_mult)]))
        self.norm = RMSNorm(dim)

    def forward_tokens(self, x, mask=None):
        b, n, dim = x.shape
        for attn, ff in self.layers:
            x = ops.add(attn.forward_tokens(x, mask), x)
            h = ops.linear(ff[0].forward_rows(x.reshape(b * n, dim)), ff[1].weight, ff[1].bias)
            h = ops.linear(ops.unary(U_GELU, h), ff[3].weight, ff[3].bias).view(b, n, dim)
            x = ops.add(h, x)
        return self.norm.forward_rows(x.reshape(b * n, dim)).view(b, n, dim)


class TextEncoder(nn.Module):
    """Learned transformer over CLIP token encodings (ref :808-867).  The frozen OpenCLIP tower is third-party with
    weights that are not available offline: pass pre-encoded ``text_encodings`` (b, n, clip_dim_latent); zeros mark
    padding exactly as in the reference (mask = (encodings != 0).any(-1))."""

    def __init__(self, *, dim, depth, clip=None, dim_head=64, heads=8, clip_dim_latent=512):
        super().__init__()
        self.dim = dim
        self.clip = clip
        dim_latent = clip.dim_latent if exists(clip) else clip_dim_latent
        self.learned_global_token = nn.Parameter(torch.randn(dim))
        self.project_in = nn.Linear(dim_latent, dim) if dim_latent != dim else nn.Identity()
        self.transformer = Transformer(dim=dim, depth=depth, dim_head=dim_head, heads=heads)

    def forward(self, texts=None, text_encodings=None):
        assert exists(texts) ^ exists(text_encodings)
        if not exists(text_encodings):
            assert exists(self.clip), "raw texts need an OpenClipAdapter (not available offline); pass text_encodings"
            with torch.no_grad():
                _, text_encodings = self.clip.embed_texts(texts)
        enc = text_encodings.float()
        mask = (enc != 0.).any(dim=-1)
        b, n, _ = enc.shape
        x = enc
        if not isinstance(self.project_in, nn.Identity):
            x = ops.linear(enc.reshape(b * n, -1), self.project_in.weight, self.project_in.bias).view(b, n, self.dim)
        g = self.learned_global_token.reshape(1, 1, -1).expand(b, 1, self.dim)
        x = torch.cat((g, x), dim=1)
        x = self.transformer.forward_tokens(x, F.pad(mask, (1, 0), value=True))
        return x[:, 0], x[:, 1:], mask


class CrossAttention(nn.Module):
    def __init__(self, dim, dim_context, dim_head=64, heads=8):
        super().__init__()
        self.heads, self.dim_head, self.scale = heads, dim_head, dim_head ** -0.5
        inner = dim_head * heads
        kv_in = dim_context if exists(dim_context) else dim
        self.norm = ChannelRMSNorm(dim)
        self.norm_context = RMSNorm(kv_in)
        self.to_q = nn.Conv2d(dim, inner, 1, bias=False)
        self.to_kv = nn.Linear(kv_in, inner * 2, bias=False)
        self.to_out = nn.Conv2d(inner, dim, 1, bias=False)

    def forward_nhwc(self, x, context, mask=None, residual=None):
        b, hh, ww, _ = x.shape
        inner = self.heads * self.dim_head
        nctx = context.shape[1]
        q = ops.conv2d(self.norm.forward_nhwc(x), self.to_q.weight).view(b, hh * ww, inner)
        ctx = self.norm_context.forward_rows(context.float().reshape(b * nctx, -1))
        kv = ops.linear(ctx, self.to_kv.weight).view(b, nctx, 2 * inner).to(x.dtype)
        lp = (nctx + 15) // 16 * 16                       # key axis padded to 16 tokens (aligned rows for the TMA GEMMs)
        if lp > nctx:
            kv = torch.cat((kv, torch.zeros((b, lp - nctx, 2 * inner), dtype=kv.dtype, device=kv.device)), dim=1)
        k, v = kv[..., :inner].contiguous(), kv[..., inner:].contiguous()
        bias = torch.zeros((b, lp), dtype=torch.float32, device=x.device)
        if exists(mask):
            bias[:, :nctx].masked_fill_(~mask, NEG_MASK)
        bias[:, nctx:] = NEG_MASK
        o = _masked_attention(q, k, v, bias, self.heads, self.scale)
        return ops.conv2d(o.view(b, hh, ww, inner), self.to_out.weight, res=residual)


class CrossAttentionBlock(nn.Module):
    def __init__(self, dim, dim_context, dim_head=64, heads=8, ff_mult=4):
        super().__init__()
        self.attn = CrossAttention(dim=dim, dim_context=dim_context, dim_head=dim_head, heads=heads)
        self.ff = FeedForwardParams(dim, ff_mult)

    def forward_nhwc(self, x, context, mask=None):
        x = self.attn.forward_nhwc(x, context, mask, residual=x)
        h = ops.conv2d(self.ff[0].forward_nhwc(x), self.ff[1].weight, self.ff[1].bias)
        return ops.conv2d(ops.unary(U_GELU, h), self.ff[3].weight, self.ff[3].bias, res=x)
