"""CPU/fp32 ORACLE for the GigaGAN G+D training hot path.  TEST INFRASTRUCTURE ONLY.

This file is a from-scratch *functional* restatement (plain torch ops on a
``state_dict``) of the algorithm that lucidrains/gigagan-pytorch @ 0806433f runs
for the unconditional generator / discriminator step.  It is the checker the
CUDA product path is compared against; it is never imported by the product
package ``gigagan_pytorch_b200`` (only ``tests/``, ``__graft_entry__.smoke()`` and
the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may use it).

Parity status: PINNED.  ``tests/test_oracle_vs_reference.py`` (runs in the build
container where /root/reference exists) checks every function below against the
unmodified reference modules on identical state_dicts/seeds, and
``oracle/make_golden.py`` wrote the known-answer fixtures in ``tests/golden/``
from the reference itself (the reference has no tests/golden vectors of its own,
SURVEY.md section 4).

All "ref:" citations are paths relative to /root/reference/gigagan_pytorch/.
Tensors at this level are NCHW fp32, like the reference's.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]


# --------------------------------------------------------------------------- #
# configuration arithmetic (ref: gigagan_pytorch.py:1026-1040 G, :1565-1578 D)
# --------------------------------------------------------------------------- #

def generator_plan(image_size: int, dim_capacity: int = 16, dim_max: int = 2048, dim_latent: int = 512,
                   num_skip_layers_excite: int = 0, self_attn_resolutions=(32, 16), num_conv_kernels: int = 2,
                   self_attn_heads: int = 8, self_attn_dim_head: int = 64, unconditional: bool = True,
                   cross_attn_resolutions=(32, 16), cross_attn_heads: int = 8, cross_attn_dim_head: int = 64):
    n = int(math.log2(image_size)) - 1
    res = [image_size // (2 ** (n - 1 - i)) for i in range(n)]
    dims = [min((2 ** (i + 1)) * dim_capacity, dim_max) for i in range(n)][::-1]
    dims = [dim_latent] + dims
    pairs = list(zip(dims[:-1], dims[1:]))
    kmod = num_conv_kernels if num_conv_kernels > 1 else 0
    split = [dim_latent, kmod]
    layers = []
    for i, ((ci, co), r) in enumerate(zip(pairs, res)):
        split += [ci, kmod, co, kmod, co, 0]
        layers.append(dict(index=i, dim_in=ci, dim_out=co, resolution=r, upsample=i > 0, upsample_rgb=i + 1 < n,
                           squeeze_excite=num_skip_layers_excite > 0 and i + num_skip_layers_excite < n,
                           self_attn=r in self_attn_resolutions,
                           cross_attn=(r in cross_attn_resolutions) and not unconditional))
    return dict(layers=layers, split=split, num_skip_layers_excite=num_skip_layers_excite, dim_latent=dim_latent,
                heads=self_attn_heads, dim_head=self_attn_dim_head, cross_heads=cross_attn_heads,
                cross_dim_head=cross_attn_dim_head)


def discriminator_plan(image_size: int, dim_capacity: int = 16, dim_max: int = 2048, channels: int = 3,
                       attn_resolutions=(32, 16), multiscale_input_resolutions=(64, 32, 16, 8),
                       multiscale_output_skip_stages: int = 1, aux_recon_resolutions=(8,),
                       aux_recon_patch_dims=(2,), aux_recon_frac_patches=(0.25,), num_skip_layers_excite: int = 0,
                       attn_heads: int = 8, attn_dim_head: int = 64):
    n = int(math.log2(image_size)) - 1
    res = [image_size // (2 ** i) for i in range(n)]
    dims = [min(d, dim_max) for d in [channels] + [(2 ** (i + 1)) * dim_capacity for i in range(n)]]
    pairs = list(zip(dims[:-1], dims[1:]))
    ms_in = [r for r in multiscale_input_resolutions if r < image_size]
    ms_out = [r // (2 ** multiscale_output_skip_stages) for r in ms_in]
    recon = {r: (p, f) for r, p, f in zip(aux_recon_resolutions, aux_recon_patch_dims, aux_recon_frac_patches)}
    layers = []
    for i, ((ci, co), r) in enumerate(zip(pairs, res)):
        layers.append(dict(index=i, dim_in=ci, dim_out=co, resolution=r, downsample=i + 1 < n,
                           squeeze_excite=i > 0 and num_skip_layers_excite > 0 and i + num_skip_layers_excite < n,
                           attn=r in attn_resolutions, predictor=r in ms_out, recon=recon.get(r)))
    return dict(layers=layers, ms_in=ms_in, ms_out=ms_out, num_skip_layers_excite=num_skip_layers_excite,
                image_size=image_size, heads=attn_heads, dim_head=attn_dim_head)


def _sub(sd: SD, prefix: str) -> SD:
    n = len(prefix)
    return {k[n:]: v for k, v in sd.items() if k.startswith(prefix)}


# --------------------------------------------------------------------------- #
# primitive blocks
# --------------------------------------------------------------------------- #

def leaky(x: Tensor) -> Tensor:
    # ref: gigagan_pytorch.py:109-110 (negative slope 0.2 everywhere)
    return torch.where(x >= 0, x, 0.2 * x)


def channel_rmsnorm(x: Tensor, gamma: Tensor) -> Tensor:
    # ref: gigagan_pytorch.py:224-232  F.normalize(dim=1) * sqrt(C) * gamma ; normalize eps 1e-12 on the norm
    c = x.shape[1]
    nrm = x.pow(2).sum(dim=1, keepdim=True).sqrt().clamp(min=1e-12)
    return x / nrm * (c ** 0.5) * gamma.view(1, c, 1, 1)


def blur3(x: Tensor) -> Tensor:
    # ref: gigagan_pytorch.py:246-255 + kornia filter2d(normalized=True): [1,2,1]x[1,2,1]/16, reflect border
    c = x.shape[1]
    k1 = torch.tensor([1.0, 2.0, 1.0], dtype=x.dtype, device=x.device)
    k = (k1[:, None] * k1[None, :]) / 16.0
    xp = F.pad(x, (1, 1, 1, 1), mode="reflect")
    return F.conv2d(xp, k.expand(c, 1, 3, 3).contiguous(), groups=c)


def upsample2x(x: Tensor) -> Tensor:
    # ref: gigagan_pytorch.py:257-261  bilinear x2 (align_corners=False) then blur
    return blur3(F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False))


def squeeze_excite(sd: SD, x: Tensor) -> Tensor:
    # ref: gigagan_pytorch.py:297-307  mean -> Linear -> SiLU -> Linear -> Sigmoid -> (b c 1 1)
    h = x.mean(dim=(2, 3))
    h = F.linear(h, sd["1.weight"], sd["1.bias"])
    h = h * torch.sigmoid(h)
    h = torch.sigmoid(F.linear(h, sd["3.weight"], sd["3.bias"]))
    return h[:, :, None, None]


def adaptive_conv2d_mod(weights: Tensor, fmap: Tensor, mod: Tensor, kernel_mod: Optional[Tensor] = None,
                        demod: bool = True, eps: float = 1e-8) -> Tensor:
    """ref: gigagan_pytorch.py:344-409.  weights (n,o,i,k,k); fmap (b,i,h,w); mod (b',i); kernel_mod (b',n)."""
    b = fmap.shape[0]
    n, o, i, k, _ = weights.shape
    if mod.shape[0] != b:                      # ref :365-366 scale-major repeat '(s b)'
        mod = mod.repeat(b // mod.shape[0], 1)
    if n > 1:
        assert kernel_mod is not None and kernel_mod.numel() > 0
        if kernel_mod.shape[0] != b:           # ref :373-374
            kernel_mod = kernel_mod.repeat(b // kernel_mod.shape[0], 1)
        attn = kernel_mod.softmax(dim=-1)      # ref :387
        w = torch.einsum("bn,noikl->boikl", attn, weights)   # ref :390
    else:
        w = weights.expand(b, o, i, k, k) if weights.shape[0] == 1 else weights
    w = w * (mod[:, None, :, None, None] + 1.0)              # ref :394-396
    if demod:                                                # ref :398-400
        inv = w.pow(2).sum(dim=(2, 3, 4), keepdim=True).clamp(min=eps).rsqrt()
        w = w * inv
    pad = (k - 1) // 2                                       # ref :312-313,406 (stride 1, dilation 1)
    y = F.conv2d(fmap.reshape(1, b * i, *fmap.shape[2:]), w.reshape(b * o, i, k, k), padding=pad, groups=b)
    return y.reshape(b, o, *y.shape[2:])


def style_network(sd: SD, z: Tensor, depth: int, lr_mul: float = 0.1, text_latent: Optional[Tensor] = None) -> Tensor:
    # ref: gigagan_pytorch.py:871-887 EqualLinear, :910-921 StyleNetwork.forward
    x = z / z.pow(2).sum(dim=1, keepdim=True).sqrt().clamp(min=1e-12)
    if text_latent is not None:                                    # ref :917-919
        x = torch.cat((x, text_latent), dim=-1)
    for d in range(depth):
        x = leaky(F.linear(x, sd[f"net.{2 * d}.weight"] * lr_mul, sd[f"net.{2 * d}.bias"] * lr_mul))
    return x


def self_attention(sd: SD, fmap: Tensor, heads: int = 8, dim_head: int = 64, dot_product: bool = False) -> Tensor:
    """ref: gigagan_pytorch.py:538-594."""
    b, _, hh, ww = fmap.shape
    x = channel_rmsnorm(fmap, sd["norm.gamma"])
    q = F.conv2d(x, sd["to_q.weight"])
    v = F.conv2d(x, sd["to_v.weight"])
    k = F.conv2d(x, sd["to_k.weight"]) if "to_k.weight" in sd else q      # ref :560 shared q/k space

    def split(t):   # 'b (h d) x y -> (b h) (x y) d'
        return t.reshape(b, heads, dim_head, hh * ww).permute(0, 1, 3, 2).reshape(b * heads, hh * ww, dim_head)

    q, k, v = split(q), split(k), split(v)
    nk = sd["null_kv"][0][None].expand(b, heads, dim_head).reshape(b * heads, 1, dim_head)   # ref :566
    nv = sd["null_kv"][1][None].expand(b, heads, dim_head).reshape(b * heads, 1, dim_head)
    k = torch.cat((nk, k), dim=1)
    v = torch.cat((nv, v), dim=1)
    if dot_product:
        sim = q @ k.transpose(1, 2)                                          # ref :574
    else:                                                                    # ref :577-580
        sim = -((q * q).sum(-1)[:, :, None] + (k * k).sum(-1)[:, None, :] - 2.0 * (q @ k.transpose(1, 2)))
    attn = (sim * dim_head ** -0.5).softmax(dim=-1)                          # ref :584-588
    out = attn @ v
    out = out.reshape(b, heads, hh * ww, dim_head).permute(0, 1, 3, 2).reshape(b, heads * dim_head, hh, ww)
    return F.conv2d(out, sd["to_out.weight"])


def self_attention_block(sd: SD, x: Tensor, dot_product: bool, heads: int = 8, dim_head: int = 64) -> Tensor:
    # ref: gigagan_pytorch.py:744-760 (+ FeedForward :726-740, exact erf GELU)
    x = self_attention(_sub(sd, "attn."), x, heads, dim_head, dot_product) + x
    h = channel_rmsnorm(x, sd["ff.0.gamma"])
    h = F.conv2d(h, sd["ff.1.weight"], sd["ff.1.bias"])
    h = F.gelu(h)
    h = F.conv2d(h, sd["ff.3.weight"], sd["ff.3.bias"])
    return h + x


# --------------------------------------------------------------------------- #
# generator (ref: gigagan_pytorch.py:1140-1250)
# --------------------------------------------------------------------------- #

def generator_forward(sd: SD, plan: dict, noise: Tensor, style_depth: int = 4,
                      layer_noises: Optional[Sequence[Tensor]] = None, return_all_rgbs: bool = False,
                      self_attn_dot_product: bool = True, global_text_tokens: Optional[Tensor] = None,
                      fine_text_tokens: Optional[Tensor] = None, text_mask: Optional[Tensor] = None):
    """``layer_noises``: the 2*num_layers per-layer noise images (b,1,h,w); when None they are drawn with
    torch.randn in the reference's order (ref :938) so that a shared manual_seed reproduces the reference."""
    b = noise.shape[0]
    styles = style_network(_sub(sd, "style_network."), noise, style_depth, text_latent=global_text_tokens)
    mods = F.linear(styles, sd["style_to_conv_modulations.weight"], sd["style_to_conv_modulations.bias"])
    mods = list(mods.split(plan["split"], dim=-1))                 # ref :1184-1186

    def nxt():
        return mods.pop(0)

    x = sd["init_block"][None].expand(b, -1, -1, -1)               # ref :1192
    x = adaptive_conv2d_mod(sd["init_conv.weights"], x, nxt(), nxt())
    rgb = torch.zeros(b, sd["layers.0.2.weights"].shape[1], 4, 4, dtype=x.dtype, device=x.device)
    excitations: List[Optional[Tensor]] = [None] * plan["num_skip_layers_excite"]
    rgbs = []
    ln = list(layer_noises) if layer_noises is not None else None

    def noise_img(t):
        if ln is not None:
            return ln.pop(0)
        return torch.randn(t.shape[0], 1, t.shape[2], t.shape[3], device=t.device)   # ref :938

    for L in plan["layers"]:
        p = f"layers.{L['index']}."
        if L["upsample"]:
            x = upsample2x(x)                                      # ref :1209-1210
        if L["squeeze_excite"]:
            excitations.append(squeeze_excite(_sub(sd, p + "0."), x))   # ref :1212-1214
        ex = excitations.pop(0) if excitations else None           # ref :1216-1218
        if ex is not None:
            x = x * ex
        x = adaptive_conv2d_mod(sd[p + "1.0.weights"], x, nxt(), nxt())
        x = leaky(x + sd[p + "1.1.weight"][None] * noise_img(x))   # ref :1221-1222, Noise :940
        x = adaptive_conv2d_mod(sd[p + "1.3.weights"], x, nxt(), nxt())
        x = leaky(x + sd[p + "1.4.weight"][None] * noise_img(x))
        if L["self_attn"]:
            x = self_attention_block(_sub(sd, p + "3."), x, self_attn_dot_product, plan["heads"], plan["dim_head"])
        if L.get("cross_attn"):                                    # ref :1231-1232
            x = cross_attention_block(_sub(sd, p + "4."), x, fine_text_tokens, text_mask, plan["cross_heads"],
                                      plan["cross_dim_head"])
        rgb = rgb + adaptive_conv2d_mod(sd[p + "2.weights"], x, nxt(), nxt(), demod=False)   # ref :1234-1236
        rgbs.append(rgb)
        if L["upsample_rgb"]:
            rgb = upsample2x(rgb)
    assert not mods
    return (rgb, rgbs) if return_all_rgbs else rgb


# --------------------------------------------------------------------------- #
# discriminator (ref: gigagan_pytorch.py:1697-1838)
# --------------------------------------------------------------------------- #

def predictor(sd: SD, x: Tensor, depth: int = 2) -> Tensor:
    # ref: gigagan_pytorch.py:1472-1498 (unconditional: plain convs)
    residual = F.conv2d(x, sd["residual_fn.weight"], sd["residual_fn.bias"])
    for d in range(depth):
        inner = x
        x = leaky(F.conv2d(x, sd[f"layers.{d}.0.weight"], sd[f"layers.{d}.0.bias"], padding=1))
        x = leaky(F.conv2d(x, sd[f"layers.{d}.2.weight"], sd[f"layers.{d}.2.bias"], padding=1))
        x = (x + inner) * (2 ** -0.5)
    x = x + residual
    return F.conv2d(x, sd["to_logits.weight"], sd["to_logits.bias"])


def simple_decoder(sd: SD, fmap: Tensor, image: Tensor, patch_dim: int, frac: float,
                   dropout_mask: Optional[Tensor] = None, patch_indices: Optional[Tensor] = None,
                   training: bool = True, p_drop: float = 0.5) -> Tensor:
    """ref: gigagan_pytorch.py:1290-1317.  dropout_mask is the keep-mask already scaled by 1/(1-p);
    patch_indices (b, n_sel) long.  When None they are drawn like the reference (dropout on fmap's device,
    permutation from a CPU randn argsort, ref :1310)."""
    if training:
        if dropout_mask is None:
            fmap = F.dropout(fmap, p_drop, training=True)
        else:
            fmap = fmap * dropout_mask
    if frac < 1.0:
        b = fmap.shape[0]

        def patches(t):   # 'b c (p1 h) (p2 w) -> b (p1 p2) c h w'
            bb, c, hh, ww = t.shape
            h, w = hh // patch_dim, ww // patch_dim
            return t.reshape(bb, c, patch_dim, h, patch_dim, w).permute(0, 2, 4, 1, 3, 5).reshape(
                bb, patch_dim * patch_dim, c, h, w)

        fm, im = patches(fmap), patches(image)
        total = patch_dim ** 2
        nsel = max(int(frac * total), 1)
        if patch_indices is None:
            patch_indices = torch.randn((b, total)).sort(dim=-1).indices[:, :nsel]
        idx = patch_indices.to(fmap.device)
        ar = torch.arange(b, device=fmap.device)[:, None]
        fmap = fm[ar, idx].flatten(0, 1)
        image = im[ar, idx].flatten(0, 1)
    x = F.conv2d(fmap, sd["net.0.weight"], sd["net.0.bias"], padding=1)
    i = 1
    while f"net.{i}.1.weight" in sd:
        x = upsample2x(x)
        x = leaky(F.conv2d(x, sd[f"net.{i}.1.weight"], sd[f"net.{i}.1.bias"], padding=1))
        i += 1
    return F.mse_loss(x, image)


def real_images_to_rgbs(images: Tensor, plan: dict) -> List[Tensor]:
    # ref: gigagan_pytorch.py:1683-1687 (bilinear, align_corners default False, no antialias)
    return [F.interpolate(images, r, mode="bilinear") for r in plan["ms_in"]]


def discriminator_forward(sd: SD, plan: dict, images: Tensor, rgbs: Sequence[Tensor],
                          return_multiscale_outputs: bool = True, calc_aux_loss: bool = True,
                          training: bool = True, recon_dropout_mask=None, recon_patch_indices=None,
                          text_embeds: Optional[Tensor] = None, num_conv_kernels: int = 2):
    """``text_embeds`` (b, text_dim): global text tokens for the text-conditioned predictors (ref :1709-1723)."""
    conv_mods = None
    if text_embeds is not None:
        dims = []
        for L in plan["layers"]:
            if L["predictor"]:
                dims.extend([L["dim_out"], num_conv_kernels if num_conv_kernels > 1 else 0])
        conv_mods = list(F.linear(text_embeds, sd["text_to_conv_conditioning.weight"],
                                  sd["text_to_conv_conditioning.bias"]).split(dims, dim=-1))
    x = images
    batch = x.shape[0]
    by_res = {t.shape[-1]: t for t in rgbs}
    ms_out, aux = [], []
    excitations: List[Optional[Tensor]] = [None] * (plan["num_skip_layers_excite"] + 1)   # ref :1754
    for L in plan["layers"]:
        p = f"layers.{L['index']}."
        res = x.shape[-1]
        if L["squeeze_excite"]:
            excitations.append(squeeze_excite(_sub(sd, p + "0."), x))
        ex = excitations.pop(0) if excitations else None
        if ex is not None:
            x = x * ex.repeat(x.shape[0] // ex.shape[0], 1, 1, 1)        # ref :1766 '(s b)'
        prev = x.shape[0]
        if res in plan["ms_in"]:                                           # ref :1772-1789
            f = F.conv2d(by_res[res], sd[p + "1.weight"], sd[p + "1.bias"], padding=3)
            f = f.repeat(x.shape[0] // f.shape[0], 1, 1, 1)
            x = torch.cat((x + f, f), dim=0)
        residual = F.conv2d(x, sd[p + "3.weight"], sd[p + "3.bias"], stride=2 if L["downsample"] else 1)
        x = leaky(F.conv2d(x, sd[p + "2.0.weight"], sd[p + "2.0.bias"], padding=1))
        x = leaky(F.conv2d(x, sd[p + "2.2.weight"], sd[p + "2.2.bias"], padding=1))
        if L["attn"]:
            x = self_attention_block(_sub(sd, p + "4."), x, False, plan["heads"], plan["dim_head"])
        if L["predictor"]:
            if conv_mods is not None:
                mod, kmod = conv_mods.pop(0), conv_mods.pop(0)             # ref :1799-1800 (consumed even if unused)
                if return_multiscale_outputs:
                    ms_out.append(predictor_conditional(_sub(sd, p + "5."), x[:prev], mod, kmod))
            elif return_multiscale_outputs:
                ms_out.append(predictor(_sub(sd, p + "5."), x[:prev]))     # ref :1803-1804
        if L["downsample"]:                                                # ref :289-293 pixel-unshuffle + 1x1
            x = F.conv2d(F.pixel_unshuffle(x, 2), sd[p + "7.1.weight"], sd[p + "7.1.bias"])
        x = (x + residual) * (2 ** -0.5)                                   # ref :1809-1810
        if L["recon"] is not None and calc_aux_loss:                       # ref :1812-1827 (Q3: post-downsample x)
            pd, fr = L["recon"]
            aux.append(simple_decoder(_sub(sd, p + "6."), x[:batch], images, pd, fr,
                                      recon_dropout_mask, recon_patch_indices, training))
    x = F.conv2d(x, sd["to_logits.0.weight"], sd["to_logits.0.bias"], padding=1)   # ref :1655-1660
    logits = F.linear(x.flatten(1), sd["to_logits.2.weight"], sd["to_logits.2.bias"])[:, 0]
    return logits.reshape(-1, batch), ms_out, aux                          # ref :1836 '(s b) -> s b'


# --------------------------------------------------------------------------- #
# losses and the two step objectives (ref: gigagan_pytorch.py:120-163, :2227-2610)
# --------------------------------------------------------------------------- #

def generator_hinge_loss(fake: Tensor) -> Tensor:
    return fake.mean()                                                      # ref :159-160


def discriminator_hinge_loss(real: Tensor, fake: Tensor) -> Tensor:
    return (F.relu(1 + real) + F.relu(1 - fake)).mean()                     # ref :162-163


def gradient_penalty(images: Tensor, outputs: Sequence[Tensor], grad_output_weights: Sequence[float],
                     weight: float = 10.0) -> Tensor:
    # ref: gigagan_pytorch.py:120-155 (no GradScaler: bf16/fp32)
    g, = torch.autograd.grad(outputs=list(outputs), inputs=images,
                             grad_outputs=[torch.ones_like(o) * w for o, w in zip(outputs, grad_output_weights)],
                             create_graph=True, retain_graph=True)
    return weight * (g.flatten(1).norm(2, dim=1) ** 2).mean()


def discriminator_step_loss(sd_d: SD, plan_d: dict, real: Tensor, fake: Tensor, fake_rgbs: Sequence[Tensor],
                            apply_gradient_penalty: bool, ms_weight: float = 0.1, aux_weight: float = 1.0,
                            calc_multiscale_loss: bool = True, recon_dropout_mask=None, recon_patch_indices=None):
    """Loss of ref train_discriminator_step (:2320-2422), one micro-batch, unconditional, no VD.
    ``real``/``fake``/``fake_rgbs`` must require grad when the penalty is on (ref :2275, :2311-2316)."""
    real_rgbs = real_images_to_rgbs(real, plan_d)
    f_logits, f_ms, _ = discriminator_forward(sd_d, plan_d, fake, fake_rgbs, calc_multiscale_loss, False)
    r_logits, r_ms, aux = discriminator_forward(sd_d, plan_d, real, real_rgbs, calc_multiscale_loss, True,
                                                recon_dropout_mask=recon_dropout_mask,
                                                recon_patch_indices=recon_patch_indices)
    div = discriminator_hinge_loss(r_logits, f_logits)
    ms = sum((discriminator_hinge_loss(r, f) for f, r in zip(f_ms, r_ms)), torch.zeros((), device=real.device))
    gp = torch.zeros((), device=real.device)
    if apply_gradient_penalty:
        w = [1.0] + [ms_weight] * len(r_ms)
        gp = gradient_penalty(real, [r_logits, *r_ms], w) + gradient_penalty(fake, [f_logits, *f_ms], w)
    aux_loss = sum(aux, torch.zeros((), device=real.device))
    total = div + gp + ms * ms_weight + aux_loss * aux_weight
    return total, dict(divergence=div, multiscale=ms, gradient_penalty=gp, aux_reconstruction=aux_loss)


def generator_step_loss(sd_g: SD, plan_g: dict, sd_d: SD, plan_d: dict, noise: Tensor, style_depth: int = 4,
                        layer_noises=None, ms_weight: float = 0.1, calc_multiscale_loss: bool = True):
    """Loss of ref train_generator_step (:2518-2562), one micro-batch, unconditional, no VD."""
    img, rgbs = generator_forward(sd_g, plan_g, noise, style_depth, layer_noises, return_all_rgbs=True)
    logits, ms, _ = discriminator_forward(sd_d, plan_d, img, rgbs, calc_multiscale_loss, False)
    div = generator_hinge_loss(logits)
    msd = sum((generator_hinge_loss(m) for m in ms), torch.zeros((), device=noise.device))
    return div + msd * ms_weight, dict(divergence=div, multiscale=msd), img


def adamw_step(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: int, lr=2e-4, b1=0.5, b2=0.9, eps=1e-8,
               wd: Optional[float] = None):
    """torch.optim.AdamW update as the reference configures it (ref optimizer.py:10-34; SURVEY Q2: wd=1e-2 on
    ndim>=2 parameters, 0 otherwise).  In-place on p, m, v."""
    if wd is None:
        wd = 1e-2 if p.ndim >= 2 else 0.0
    p.mul_(1 - lr * wd)
    m.lerp_(g, 1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-lr / bc1)


# --------------------------------------------------------------------------- #
# a whole optimiser step on the CPU (used as the reference arm / cpu_baseline of bench.py)
# --------------------------------------------------------------------------- #

class OracleTrainer:
    """Minimal unconditional G+D trainer over the functional oracle: one D step (optionally with the gradient
    penalty) and one G step per call, AdamW as the reference configures it.  ref: gigagan_pytorch.py:2227-2610."""

    def __init__(self, sd_g: SD, plan_g: dict, sd_d: SD, plan_d: dict, style_depth: int = 4, lr: float = 2e-4):
        self.sd_g = {k: v.detach().clone().requires_grad_(v.is_floating_point() and not k.endswith(".f"))
                     for k, v in sd_g.items()}
        self.sd_d = {k: v.detach().clone().requires_grad_(v.is_floating_point() and not k.endswith(".f"))
                     for k, v in sd_d.items()}
        self.plan_g, self.plan_d, self.style_depth = plan_g, plan_d, style_depth

        def groups(sd):
            ps = [p for p in sd.values() if p.requires_grad]
            return [dict(params=[p for p in ps if p.ndim >= 2]), dict(params=[p for p in ps if p.ndim < 2],
                                                                       weight_decay=0.0)]
        self.opt_g = torch.optim.AdamW(groups(self.sd_g), lr=lr, betas=(0.5, 0.9), weight_decay=1e-2)
        self.opt_d = torch.optim.AdamW(groups(self.sd_d), lr=lr, betas=(0.5, 0.9), weight_decay=1e-2)

    def step(self, real: Tensor, apply_gradient_penalty: bool):
        b = real.shape[0]
        z = torch.randn(b, self.sd_g["style_network.net.0.weight"].shape[1])
        with torch.no_grad():
            fake, rgbs = generator_forward(self.sd_g, self.plan_g, z, self.style_depth, return_all_rgbs=True)
        real = real.clone().requires_grad_(apply_gradient_penalty)
        fake = fake.detach().requires_grad_(apply_gradient_penalty)
        rgbs = [t.detach().requires_grad_(apply_gradient_penalty) for t in rgbs]
        self.opt_d.zero_grad()
        d_loss, parts = discriminator_step_loss(self.sd_d, self.plan_d, real, fake, rgbs, apply_gradient_penalty)
        d_loss.backward()
        self.opt_d.step()
        self.opt_g.zero_grad()
        z = torch.randn(b, z.shape[1])
        g_loss, _, _ = generator_step_loss(self.sd_g, self.plan_g, self.sd_d, self.plan_d, z, self.style_depth)
        g_loss.backward()
        self.opt_g.step()
        return float(d_loss.detach()), float(g_loss.detach())


# --------------------------------------------------------------------------- #
# UnetUpsampler (ref: unet_upsampler.py:447-898), unconditional image path
# --------------------------------------------------------------------------- #

def unet_plan(dim, image_size, input_image_size, dim_mults=(1, 2, 4, 8, 16), channels=3,
              full_attn=(False, False, False, True, True), cross_attn=None, num_conv_kernels=2, init_dim=None,
              self_attn_dim_head=64, self_attn_heads=8, cross_attn_dim_head=64, attn_depths=(1, 1, 1, 1, 1),
              mid_attn_depth=1, skip_connect_scale=None, self_attn_ff_mult=4):
    """Dimension bookkeeping of the reference constructor (unet_upsampler.py:509-640), incl. the modulation split list
    (ResnetBlock appends [dim, n, dim_out, n] per block in construction order) and its quirk Q7: the `ups` loop zips
    reversed(in_out) with reversed(full_attn) under swapped names, so the ups attention type follows full_attn
    reversed through the *cross_attn* slot order (unet_upsampler.py:596)."""
    init_dim = init_dim or dim
    dims = [init_dim] + [dim * m for m in dim_mults]
    in_out = list(zip(dims[:-1], dims[1:]))
    n_no_down = int(math.log2(image_size) - math.log2(input_image_size))
    split, skip_dims = [], []
    nk = num_conv_kernels

    def res(ci, co):
        split.extend([ci, nk, co, nk])
        return dict(dim_in=ci, dim_out=co)

    downs = []
    for ind, ((ci, co), fa, depth) in enumerate(zip(in_out, full_attn, attn_depths)):
        no_down = ind < n_no_down
        skip_dims.append(ci)
        skip_dims.append(ci + (co if not no_down else 0))
        downs.append(dict(block1=res(ci, ci), block2=res(ci, ci), full_attn=fa, depth=depth, dim=ci, dim_out=co,
                          skip_downsample=no_down))
    mid_dim = dims[-1]
    mid = dict(block1=res(mid_dim, mid_dim), depth=mid_attn_depth, block2=None, dim=mid_dim)
    mid["block2"] = res(mid_dim, mid_dim)
    ups = []
    # ref :596: zip(reversed(in_out), reversed(full_attn) [named layer_cross_attn], reversed(cross_attn) [named
    # layer_full_attn], ...) -> with the default cross_attn == full_attn the attention type is full_attn reversed.
    for (ci, co), fa, depth in zip(reversed(in_out), reversed(cross_attn if cross_attn is not None else full_attn),
                                   reversed(attn_depths)):
        b1 = res(ci + skip_dims.pop(), ci)
        b2 = res(ci + skip_dims.pop(), ci)
        ups.append(dict(dim_in=ci, dim_out=co, block1=b1, block2=b2, full_attn=fa, depth=depth))
    final = res(dim, dim)
    return dict(downs=downs, mid=mid, ups=ups, final=final, split=split, channels=channels,
                heads=self_attn_heads, dim_head=self_attn_dim_head, up_dim_head=cross_attn_dim_head,
                skip_scale=skip_connect_scale if skip_connect_scale is not None else 2 ** -0.5,
                input_image_size=input_image_size, image_size=image_size)


def unet_rmsnorm(x, gamma):
    # ref unet_upsampler.py:224-234
    c = x.shape[1]
    nrm = x.pow(2).sum(dim=1, keepdim=True).sqrt().clamp(min=1e-12)
    return x / nrm * gamma.view(1, c, 1, 1) * (c ** 0.5)


def unet_block(sd, x, mods):
    # ref unet_upsampler.py:238-270  AdaptiveConv -> RMSNorm -> SiLU
    x = adaptive_conv2d_mod(sd["proj.weights"], x, mods.pop(0), mods.pop(0))
    x = unet_rmsnorm(x, sd["norm.gamma"])
    return x * torch.sigmoid(x)


def unet_resnet_block(sd, x, mods):
    # ref unet_upsampler.py:272-310
    h = unet_block(_sub(sd, "block1."), x, mods)
    h = unet_block(_sub(sd, "block2."), h, mods)
    r = F.conv2d(x, sd["res_conv.weight"], sd["res_conv.bias"]) if "res_conv.weight" in sd else x
    return h + r


def unet_linear_attention(sd, x, heads, dim_head):
    # ref unet_upsampler.py:312-349
    b, c, hh, ww = x.shape
    xn = unet_rmsnorm(x, sd["norm.gamma"])
    q, k, v = F.conv2d(xn, sd["to_qkv.weight"]).chunk(3, dim=1)
    q, k, v = (t.reshape(b, heads, dim_head, hh * ww) for t in (q, k, v))
    q = q.softmax(dim=-2) * dim_head ** -0.5
    k = k.softmax(dim=-1)
    context = torch.einsum("bhdn,bhen->bhde", k, v)
    out = torch.einsum("bhde,bhdn->bhen", context, q).reshape(b, heads * dim_head, hh, ww)
    out = F.conv2d(out, sd["to_out.0.weight"], sd["to_out.0.bias"])
    return unet_rmsnorm(out, sd["to_out.1.gamma"])


def unet_attention(sd, x, heads, dim_head):
    # ref unet_upsampler.py:351-380 + attend.py:99-108
    b, c, hh, ww = x.shape
    xn = unet_rmsnorm(x, sd["norm.gamma"])
    q, k, v = F.conv2d(xn, sd["to_qkv.weight"]).chunk(3, dim=1)
    q, k, v = (t.reshape(b, heads, dim_head, hh * ww).transpose(-1, -2) for t in (q, k, v))
    attn = ((q @ k.transpose(-1, -2)) * dim_head ** -0.5).softmax(dim=-1)
    out = (attn @ v).transpose(-1, -2).reshape(b, heads * dim_head, hh, ww)
    return F.conv2d(out, sd["to_out.weight"], sd["to_out.bias"])


def unet_transformer(sd, x, full, depth, heads, dim_head):
    # ref unet_upsampler.py:395-443 (Transformer / LinearTransformer)
    for d in range(depth):
        a = _sub(sd, f"layers.{d}.0.")
        x = (unet_attention if full else unet_linear_attention)(a, x, heads, dim_head) + x
        f = _sub(sd, f"layers.{d}.1.")
        h = unet_rmsnorm(x, f["0.gamma"])
        h = F.conv2d(F.gelu(F.conv2d(h, f["1.weight"], f["1.bias"])), f["3.weight"], f["3.bias"])
        x = h + x
    return x


def unet_downsample(sd, x, skip_downsample):
    # ref unet_upsampler.py:82-160: conv3x3, then (x - blur(x)) as a high-frequency skip and max-pool 2
    x = F.conv2d(x, sd["conv2d.weight"], sd["conv2d.bias"], padding=1)
    if skip_downsample:
        return x, x[:, 0:0]
    return F.max_pool2d(x, 2), x - blur3(x)


def unet_forward(sd: SD, plan: dict, lowres: Tensor, noise: Tensor, style_depth: int = 4,
                 return_all_rgbs: bool = False):
    """ref unet_upsampler.py:657-898 (image input, unconditional)."""
    styles = style_network(_sub(sd, "style_network."), noise, style_depth)
    mods = list(F.linear(styles, sd["style_to_conv_modulations.weight"],
                         sd["style_to_conv_modulations.bias"]).split(plan["split"], dim=-1))
    H, dh = plan["heads"], plan["dim_head"]
    x = F.conv2d(lowres, sd["init_conv.weight"], sd["init_conv.bias"], padding=3)
    hs = []
    for i, L in enumerate(plan["downs"]):
        p = f"downs.{i}."
        x = unet_resnet_block(_sub(sd, p + "0."), x, mods)
        hs.append(x)
        x = unet_resnet_block(_sub(sd, p + "1."), x, mods)
        x = unet_transformer(_sub(sd, p + "3."), x, L["full_attn"], L["depth"], H, dh)
        skip = x
        x, hf = unet_downsample(_sub(sd, p + "6."), x, L["skip_downsample"])
        hs.append(torch.cat((skip, hf), dim=1))
    x = unet_resnet_block(_sub(sd, "mid_block1."), x, mods)
    x = unet_transformer(_sub(sd, "mid_attn."), x, True, plan["mid"]["depth"], H, dh)
    x = unet_resnet_block(_sub(sd, "mid_block2."), x, mods)
    rgb = F.conv2d(x, sd["mid_to_rgb.weight"], sd["mid_to_rgb.bias"])
    rgbs = [rgb]
    for i, L in enumerate(plan["ups"]):
        p = f"ups.{i}."
        x = F.pixel_shuffle(F.silu(F.conv2d(x, sd[p + "0.net.0.weight"], sd[p + "0.net.0.bias"])), 2)   # :267-273
        rgb = upsample2x(rgb)
        r1, r2 = hs.pop() * plan["skip_scale"], hs.pop() * plan["skip_scale"]
        if x.shape[2:] != r1.shape[2:]:
            r1 = F.interpolate(r1, tuple(x.shape[2:]), mode="bilinear")
            r2 = F.interpolate(r2, tuple(x.shape[2:]), mode="bilinear")
        x = unet_resnet_block(_sub(sd, p + "5."), torch.cat((x, r1), dim=1), mods)
        x = unet_resnet_block(_sub(sd, p + "6."), torch.cat((x, r2), dim=1), mods)
        x = unet_transformer(_sub(sd, p + "8."), x, L["full_attn"], L["depth"], plan["heads"], plan["up_dim_head"])
        rgb = rgb + F.conv2d(x, sd[p + "4.weight"], sd[p + "4.bias"])
        rgbs.append(rgb)
    x = unet_resnet_block(_sub(sd, "final_res_block."), x, mods)
    assert not mods
    rgb = rgb + F.conv2d(x, sd["final_to_rgb.weight"], sd["final_to_rgb.bias"])
    if not return_all_rgbs:
        return rgb
    return rgb, [lowres] + [t for t in rgbs if t.shape[-1] > lowres.shape[-1]]


# --------------------------------------------------------------------------- #
# text conditioning (ref: gigagan_pytorch.py:234-242 RMSNorm, :596-722 attention, :780-867 Transformer/TextEncoder)
# on pre-encoded CLIP token sequences (`text_encodings`); the CLIP tower itself is third-party and absent here
# --------------------------------------------------------------------------- #

def rmsnorm_last(x, gamma):
    # ref :234-242
    d = x.shape[-1]
    return x / x.pow(2).sum(dim=-1, keepdim=True).sqrt().clamp(min=1e-12) * (d ** 0.5) * gamma


def _heads(t, h):      # 'b n (h d) -> (b h) n d'
    b, n, hd = t.shape
    return t.reshape(b, n, h, hd // h).permute(0, 2, 1, 3).reshape(b * h, n, hd // h)


def text_attention(sd, x, mask, heads, dim_head):
    # ref :678-722 (dot product, null key/value, key padding mask)
    b, n, _ = x.shape
    xn = rmsnorm_last(x, sd["norm.gamma"])
    q, k, v = F.linear(xn, sd["to_qkv.weight"]).chunk(3, dim=-1)
    q, k, v = _heads(q, heads), _heads(k, heads), _heads(v, heads)
    nk = sd["null_kv"][0][None].expand(b, heads, dim_head).reshape(b * heads, 1, dim_head)
    nv = sd["null_kv"][1][None].expand(b, heads, dim_head).reshape(b * heads, 1, dim_head)
    k, v = torch.cat((nk, k), dim=1), torch.cat((nv, v), dim=1)
    sim = (q @ k.transpose(1, 2)) * dim_head ** -0.5
    if mask is not None:
        m = F.pad(mask, (1, 0), value=True)
        m = m[:, None, None, :].expand(b, heads, 1, n + 1).reshape(b * heads, 1, n + 1)
        sim = sim.masked_fill(~m, -torch.finfo(sim.dtype).max)
    out = sim.softmax(dim=-1) @ v
    out = out.reshape(b, heads, n, dim_head).permute(0, 2, 1, 3).reshape(b, n, heads * dim_head)
    return F.linear(out, sd["to_out.weight"])


def text_transformer(sd, x, mask, depth, heads, dim_head):
    # ref :780-803
    for d in range(depth):
        x = text_attention(_sub(sd, f"layers.{d}.0."), x, mask, heads, dim_head) + x
        f = _sub(sd, f"layers.{d}.1.")
        h = rmsnorm_last(x, f["0.gamma"])
        x = F.linear(F.gelu(F.linear(h, f["1.weight"], f["1.bias"])), f["3.weight"], f["3.bias"]) + x
    return rmsnorm_last(x, sd["norm.gamma"])


def text_encoder(sd, text_encodings, depth, heads=8, dim_head=64):
    # ref :842-867 with text_encodings given -> (global tokens (b,d), fine tokens (b,n,d), mask (b,n))
    mask = (text_encodings != 0.).any(dim=-1)
    x = text_encodings
    if "project_in.weight" in sd:
        x = F.linear(x, sd["project_in.weight"], sd["project_in.bias"])
    b = x.shape[0]
    g = sd["learned_global_token"][None, None].expand(b, 1, -1)
    x = torch.cat((g, x), dim=1)
    x = text_transformer(_sub(sd, "transformer."), x, F.pad(mask, (1, 0), value=True), depth, heads, dim_head)
    return x[:, 0], x[:, 1:], mask


def cross_attention(sd, fmap, context, mask, heads=8, dim_head=64):
    # ref :617-655
    b, _, hh, ww = fmap.shape
    x = channel_rmsnorm(fmap, sd["norm.gamma"])
    ctx = rmsnorm_last(context, sd["norm_context.gamma"])
    q = F.conv2d(x, sd["to_q.weight"])
    k, v = F.linear(ctx, sd["to_kv.weight"]).chunk(2, dim=-1)
    k, v = _heads(k, heads), _heads(v, heads)
    q = q.reshape(b, heads, dim_head, hh * ww).permute(0, 1, 3, 2).reshape(b * heads, hh * ww, dim_head)
    sim = (q @ k.transpose(1, 2)) * dim_head ** -0.5
    if mask is not None:
        n = mask.shape[-1]
        m = mask[:, None, None, :].expand(b, heads, 1, n).reshape(b * heads, 1, n)
        sim = sim.masked_fill(~m, -torch.finfo(sim.dtype).max)
    out = sim.softmax(dim=-1) @ v
    out = out.reshape(b, heads, hh * ww, dim_head).permute(0, 1, 3, 2).reshape(b, heads * dim_head, hh, ww)
    return F.conv2d(out, sd["to_out.weight"])


def cross_attention_block(sd, x, context, mask, heads=8, dim_head=64):
    # ref :762-778
    x = cross_attention(_sub(sd, "attn."), x, context, mask, heads, dim_head) + x
    h = channel_rmsnorm(x, sd["ff.0.gamma"])
    h = F.conv2d(F.gelu(F.conv2d(h, sd["ff.1.weight"], sd["ff.1.bias"])), sd["ff.3.weight"], sd["ff.3.bias"])
    return h + x


def predictor_conditional(sd, x, mod, kernel_mod, depth=2):
    # ref :1472-1498 with AdaptiveConv2DMod layers (text-conditioned discriminator)
    residual = F.conv2d(x, sd["residual_fn.weight"], sd["residual_fn.bias"])
    for d in range(depth):
        inner = x
        x = leaky(adaptive_conv2d_mod(sd[f"layers.{d}.0.weights"], x, mod, kernel_mod))
        x = leaky(adaptive_conv2d_mod(sd[f"layers.{d}.2.weights"], x, mod, kernel_mod))
        x = (x + inner) * (2 ** -0.5)
    x = x + residual
    return F.conv2d(x, sd["to_logits.weight"], sd["to_logits.bias"])
