"""Differentiable operators over the C-ABI kernels (torch.autograd.Function wrappers).

Design rule: the backward of every operator here is itself written with operators from this file (never raw
kernels), so the graph built during a backward pass with ``create_graph=True`` is differentiable again.  That is
what makes the discriminator's gradient penalty (reference gigagan_pytorch.py:120-155, a double backward through
D) work on hand-written kernels.  Operators marked "first-order" (fused fast paths used outside the gradient
penalty) are the only exception.

Conventions: activations are NHWC tensors (N,H,W,C) in the compute dtype (fp32 or bf16); small statistic tensors
(per-row / per-sample-channel) are fp32.  torch is used for memory only (allocation, views, cat/slice, casts).
"""
from __future__ import annotations

import math
import os

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from ._lib import call

U_LRELU, U_RELU, U_GELU, U_SILU, U_SIGMOID, U_INVNORM, U_RSQRT_EPS8 = range(7)
ROWS, SAMPLE_CH = 0, 1
MUL, ADD = 0, 1


def _dt(t):
    if t.dtype == torch.float32:
        return 0
    if t.dtype == torch.bfloat16:
        return 1
    raise TypeError(f"unsupported dtype {t.dtype}")


def _p(t):
    return None if t is None else t.data_ptr()


def _st():
    return torch.cuda.current_stream().cuda_stream


def _c(t):
    if not t.is_cuda:
        raise RuntimeError("gigagan_pytorch_b200 operators need CUDA tensors (there is no CPU path)")
    return t if t.is_contiguous() else t.contiguous()


# ============================================================================= convolution family
class ConvGeom:
    """Static description of one convolution (kernel size, stride, padding, fused epilogue)."""
    __slots__ = ("kh", "kw", "stride", "pad", "per_sample", "act", "gain")

    def __init__(self, kh, kw, stride=1, pad=0, per_sample=False, act=0, gain=1.0):
        self.kh, self.kw, self.stride, self.pad = kh, kw, stride, pad
        self.per_sample, self.act, self.gain = per_sample, act, gain

    def out_hw(self, h, w):
        return (h + 2 * self.pad - self.kh) // self.stride + 1, (w + 2 * self.pad - self.kw) // self.stride + 1

    def plain(self):
        return ConvGeom(self.kh, self.kw, self.stride, self.pad, self.per_sample)


_WCACHE = {}


def clear_weight_cache():
    """Prepared (kernel-layout) weights are cached between the passes of one optimiser step; the trainer clears the
    cache whenever parameters change (and at the start of every captured step)."""
    _WCACHE.clear()


class WeightBank:
    """Both kernel layouts of every 4-D conv weight (and every filter of the 5-D AdaptiveConv banks) of a module
    whose parameters live in one flat fp32 buffer, refreshed by ONE kernel launch per optimiser step."""

    def __init__(self, flat, params, dtype, cin_pad):
        dev = flat.device
        self.flat, self.dtype = flat, dtype
        entries, chunks, self.views = [], [], {}
        fo = bo = 0
        base = flat.data_ptr()

        def add(ptr, shape):
            nonlocal fo, bo
            O, I, KH, KW = shape
            ipad = cin_pad(I)
            n = O * KH * KW * ipad
            eid = len(entries)
            entries.append([(ptr - base) // 4, O, I, KH * KW, ipad, fo, bo, 0])
            ti = max(1, min(32, 380 // (KH * KW)))      # (32 o) x (ti i) x taps tile per block, <= 380 floats per row
            for o0 in range(0, O, 32):
                for i0 in range(0, ipad, ti):
                    chunks.append([eid, o0, i0, ti])
            self.views[(ptr, (O, I, KH, KW), ipad)] = (fo, (O, KH, KW, ipad), bo, (ipad, KH, KW, O))
            fo += (n + 7) // 8 * 8
            bo += (n + 7) // 8 * 8

        for p in params:
            if p.ndim == 4:
                add(p.data_ptr(), tuple(p.shape))
            elif p.ndim == 5:
                per = p[0].numel() * 4
                for j in range(p.shape[0]):
                    add(p.data_ptr() + j * per, tuple(p.shape[1:]))
        self.fwd = torch.empty(max(fo, 8), dtype=dtype, device=dev)
        self.bwd = torch.empty(max(bo, 8), dtype=dtype, device=dev)
        self.entries = torch.tensor(entries, dtype=torch.int64, device=dev)
        self.chunks = torch.tensor(chunks, dtype=torch.int32, device=dev)
        self.dirty = True                  # layouts not built yet / parameters changed since the last refresh

    def refresh(self):
        call("gg_weight_prep_multi", _p(self.flat), _p(self.entries), _p(self.chunks), self.chunks.shape[0],
             _p(self.fwd), _p(self.bwd), _dt(self.fwd), _st())

    def lookup(self, weight, cin, dtype):
        if dtype != self.dtype:
            return None
        v = self.views.get((weight.data_ptr(), tuple(weight.shape), cin))
        if v is None:
            return None
        fo, fs, bo, bs = v
        n = fs[0] * fs[1] * fs[2] * fs[3]
        return self.fwd[fo:fo + n].view(fs), self.bwd[bo:bo + n].view(bs)


_BANKS = []


def register_weight_bank(bank):
    _BANKS.append(bank)


def unregister_weight_bank(bank):
    if bank in _BANKS:
        _BANKS.remove(bank)


def _bank_lookup(weight, cin, dtype):
    for b in _BANKS:
        r = b.lookup(weight, cin, dtype)
        if r is not None:
            return r
    return None


def prep_weight(weight, cin, dtype):
    """fp32 master (O,I,KH,KW) -> kernel layout (O,KH,KW,cin) in the compute dtype, zero-padding I up to cin."""
    base = weight._base if weight._base is not None else weight
    ver = -1 if base.is_inference() else base._version     # inference-mode temporaries carry no version: never cached
    key = (weight.data_ptr(), tuple(weight.shape), cin, dtype, "f")
    hit = _WCACHE.get(key)
    if hit is not None and hit[1] is base and hit[2] == ver and ver >= 0:  # same storage owner, not modified since
        return hit[0]
    r = _bank_lookup(weight, cin, dtype)
    if r is not None:
        _WCACHE[key] = (r[0], base, ver)
        _WCACHE[(r[0].data_ptr(), tuple(r[0].shape), r[0].dtype, "t")] = (r[0], r[1])
        return r[0]
    w = weight.detach().permute(0, 2, 3, 1)
    if w.shape[-1] != cin:
        w = torch.nn.functional.pad(w, (0, cin - w.shape[-1]))
    w = w.to(dtype).contiguous()
    if ver >= 0:
        _WCACHE[key] = (w, base, ver)
    return w


def flipped_weight(wk):
    """kernel-layout ([N,]Cout,KH,KW,Cin) -> ([N,]Cin,KH,KW,Cout) with the taps reversed (dgrad as a convolution)."""
    key = (wk.data_ptr(), tuple(wk.shape), wk.dtype, "t")
    hit = _WCACHE.get(key)
    if hit is not None and hit[0] is wk:
        return hit[1]
    wt = wk.flip((-3, -2)).transpose(-4, -1).contiguous()
    _WCACHE[key] = (wk, wt)
    return wt


def unprep_weight_grad(gk, like):
    """kernel-layout fp32 gradient (O,KH,KW,cin) -> master layout (O,I,KH,KW)."""
    return gk[..., : like.shape[1]].permute(0, 3, 1, 2).contiguous().to(like.dtype)


_PROFILER = [None]


class ConvProfiler:
    """Context manager: CUDA-event timing of every convolution launch (bench.py's roofline leg)."""

    def __init__(self):
        self.records = []

    def __enter__(self):
        _PROFILER[0] = self
        return self

    def __exit__(self, *a):
        _PROFILER[0] = None

    def summary(self, dtype=torch.bfloat16, kinds=("fprop", "dgrad_s2")):
        torch.cuda.synchronize()
        ms = fl = 0.0
        n = 0
        for kind, flops, e0, e1, dt, *_ in self.records:
            if dt == dtype and kind in kinds:
                ms += e0.elapsed_time(e1)
                fl += flops
                n += 1
        return dict(ms=ms, tflops=(fl / ms / 1e9) if ms else 0.0, launches=n)


class _Timed:
    """records one launch group into the active ConvProfiler (no-op otherwise)"""

    def __init__(self, kind, flops, dtype, desc):
        self.prof = _PROFILER[0]
        self.args = (kind, flops, dtype, desc)

    def __enter__(self):
        if self.prof is not None:
            self.e0, self.e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *a):
        if self.prof is not None:
            self.e1.record()
            kind, flops, dtype, desc = self.args
            self.prof.records.append((kind, flops, self.e0, self.e1, dtype, desc))


def _conv_fprop_raw(x, wk, bias, res, g, out_c):
    n, h, w, cin = x.shape
    oh, ow = g.out_hw(h, w)
    y = torch.empty((n, oh, ow, out_c), dtype=x.dtype, device=x.device)
    prof = _PROFILER[0]
    if prof is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    call("gg_conv2d_fprop", _p(x), _p(wk), _p(bias), _p(res), _p(y), n, h, w, cin, oh, ow, out_c, g.kh, g.kw, g.stride,
         g.pad, int(g.per_sample), g.act, float(g.gain), _dt(x), _st())
    if prof is not None:
        e1.record()
        prof.records.append(("fprop", 2.0 * n * oh * ow * out_c * cin * g.kh * g.kw, e0, e1, x.dtype,
                             f"n{n} {h}x{w} {cin}->{out_c} k{g.kh} s{g.stride} ps{int(g.per_sample)} act{g.act} res{int(res is not None)}"))
    return y


def _conv_dgrad_raw(gy, wk, g, in_shape):
    n, h, w, cin = in_shape
    _, oh, ow, cout = gy.shape
    if g.stride == 1:
        # data gradient of a stride-1 convolution == convolution of dy with the spatially flipped, in/out-swapped
        # filter and padding k-1-pad: reuse the forward kernel (tcgen05 when eligible).
        wt = flipped_weight(wk)                                        # ([N,]Cin,KH,KW,Cout)
        gt = ConvGeom(g.kh, g.kw, 1, g.kh - 1 - g.pad, g.per_sample)
        assert g.kh == g.kw
        return _conv_fprop_raw(gy, wt, None, None, gt, cin)
    if g.stride == 2 and g.pad == 0 and g.kh == g.kw and g.kh <= 2 and not g.per_sample and h == 2 * oh and w == 2 * ow:
        # non-overlapping stride-2 taps: one 1x1 GEMM per tap, scattered into the interleaved gradient tensor
        dx = (torch.empty if g.kh == 2 else torch.zeros)(in_shape, dtype=gy.dtype, device=gy.device)
        with _Timed("dgrad_s2", 2.0 * n * oh * ow * cout * cin * g.kh * g.kw, gy.dtype,
                    f"dgrad n{n} {h}x{w} {cin}<-{cout} k{g.kh} s2"):
            if g.kh == 2:
                # per filter row ky ONE 1x1 GEMM with 2*Cin output channels (kx, ci): for a fixed input row 2*oy+ky the
                # pixels 2*ox and 2*ox+1 are adjacent, so each output pixel's result is one contiguous 2*Cin run
                for ky in range(2):
                    wt = wk[:, ky].reshape(cout, 2 * cin).t().contiguous()       # ((kx, ci), Cout) = 1x1 kernel layout
                    call("gg_conv2d_fprop_strided", _p(gy), _p(wt), None, _p(dx), n, oh, ow, cout, oh, ow, 2 * cin, 1, 1, 1,
                         0, 0, 0, 1.0, ky * w * cin, h * w * cin, 2 * w * cin, 2 * cin, _dt(gy), _st())
            else:
                wt = wk[:, 0, 0, :].t().contiguous()                             # (Cin, Cout) = 1x1 kernel layout
                call("gg_conv2d_fprop_strided", _p(gy), _p(wt), None, _p(dx), n, oh, ow, cout, oh, ow, cin, 1, 1, 1,
                     0, 0, 0, 1.0, 0, h * w * cin, 2 * w * cin, 2 * cin, _dt(gy), _st())
        return dx
    dx = torch.empty(in_shape, dtype=gy.dtype, device=gy.device)
    call("gg_conv2d_dgrad", _p(gy), _p(wk), _p(dx), n, h, w, cin, oh, ow, cout, g.kh, g.kw, g.stride, g.pad,
         int(g.per_sample), _dt(gy), _st())
    return dx


def _conv_wgrad_raw(x, gy, g):
    n, h, w, cin = x.shape
    _, oh, ow, cout = gy.shape
    shape = (n, cout, g.kh, g.kw, cin) if g.per_sample else (cout, g.kh, g.kw, cin)
    dw = torch.empty(shape, dtype=torch.float32, device=x.device)
    with _Timed("wgrad", 2.0 * n * oh * ow * cout * cin * g.kh * g.kw, x.dtype,
                f"wgrad n{n} {h}x{w} {cin}->{cout} k{g.kh} s{g.stride} ps{int(g.per_sample)}"):
        call("gg_conv2d_wgrad", _p(x), _p(gy), _p(dw), n, h, w, cin, oh, ow, cout, g.kh, g.kw, g.stride, g.pad,
             int(g.per_sample), _dt(x), _st())
    return dw


class Conv2dFn(Function):
    """y = (act(conv(x, W) + bias) + res) * gain.  ``weight`` is either the fp32 master (O,I,KH,KW) [master=True]
    or an already prepared kernel-layout tensor ([N,]O,KH,KW,I) in x's dtype (per-sample AdaptiveConv weights)."""

    @staticmethod
    def forward(ctx, x, weight, bias, res, geom, master):
        x = _c(x)
        wk = prep_weight(weight, x.shape[-1], x.dtype) if master else _c(weight)
        out_c = wk.shape[-4]
        assert not (geom.act and (res is not None or geom.gain != 1.0)), "act and res/gain epilogues are exclusive"
        y = _conv_fprop_raw(x, wk, bias, None if res is None else _c(res), geom, out_c)
        ctx.geom, ctx.master, ctx.has_bias, ctx.has_res = geom, master, bias is not None, res is not None
        ctx.bias_ref = bias if (bias is not None and getattr(bias, "_gg_sink1", False)) else None
        ctx.save_for_backward(x, weight, y if geom.act else None)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, weight, y = ctx.saved_tensors
        g = ctx.geom
        if g.gain != 1.0:
            gy = axpby(g.gain, gy)
        gres = gy if ctx.has_res else None
        gx = gw = gb = None
        want_gb = ctx.has_bias and ctx.needs_input_grad[2] and not _skip_param_grads()
        bias_sink = None
        if want_gb and not torch.is_grad_enabled():
            bias_sink = _bias_sink(ctx.bias_ref)               # terminal pass: add straight into the flat gradient buffer
        if g.act:
            if want_gb and not torch.is_grad_enabled() and _lrelu_bias_fusable(y):
                gy, gb = _lrelu_bwd_bias(y, gy, bias_sink)  # one pass: activation gradient + bias gradient
                if bias_sink is not None:
                    want_gb = False
            else:
                gy = Unary1Fn.apply(U_LRELU, y, gy)
        pg = g.plain()
        if ctx.needs_input_grad[1] and not _skip_param_grads():      # first: it forks to the side stream when terminal
            if not (ctx.master and _wgrad_to_sink(x, gy, pg, weight)):
                gw = ConvWgradFn.apply(x, gy, pg, weight if ctx.master else None)
        if ctx.needs_input_grad[0]:
            gx = ConvDgradFn.apply(gy, weight, pg, tuple(x.shape), ctx.master)
        if want_gb and gb is None:
            if bias_sink is not None:
                gyc = _c(gy)
                C = gyc.shape[-1]
                call("gg_red_dot_sc_acc", _p(gyc), None, _p(bias_sink), gyc.numel() // C, C, gyc.numel() // C, 1, _dt(gyc), _st())
            else:
                gb = dot_sc(gy, None, gy.numel() // gy.shape[-1], 1).reshape(-1)
        return gx, gw, gb, gres, None, None


class ConvDgradFn(Function):
    @staticmethod
    def forward(ctx, gy, weight, geom, in_shape, master):
        gy = _c(gy)
        wk = prep_weight(weight, in_shape[-1], gy.dtype) if master else _c(weight)
        ctx.geom, ctx.master = geom, master
        ctx.save_for_backward(gy, weight)
        return _conv_dgrad_raw(gy, wk, geom, in_shape)

    @staticmethod
    def backward(ctx, ggx):
        gy, weight = ctx.saved_tensors
        g = ctx.geom
        d_gy = d_w = None
        if ctx.needs_input_grad[1] and not _skip_param_grads():
            if not (ctx.master and _wgrad_to_sink(ggx, gy, g, weight)):
                d_w = ConvWgradFn.apply(ggx, gy, g, weight if ctx.master else None)
        if ctx.needs_input_grad[0]:
            d_gy = Conv2dFn.apply(ggx, weight, None, None, g, ctx.master)
        return d_gy, d_w, None, None, None


def _lrelu_bias_fusable(y):
    nvec = y.shape[-1] // (8 if y.dtype == torch.bfloat16 else 4)
    return y.shape[-1] % (8 if y.dtype == torch.bfloat16 else 4) == 0 and 0 < nvec <= 256 and (nvec & (nvec - 1)) == 0


def _bias_sink(bias):
    """the fp32 .grad slice of a 1-D parameter that lives in an optimiser's flat gradient buffer (FlatAdamW marks those
    with ``_gg_sink1``), or None: bias gradients are then ADDED there by the reduction kernel itself instead of going
    through a temporary + autograd's AccumulateGrad (one launch instead of memset + kernel + add per bias)"""
    if bias is None:
        return None
    dst = bias.grad
    if dst is None or dst.dtype != torch.float32 or not dst.is_contiguous():
        return None
    return dst


def _lrelu_bwd_bias(y, gy, sink=None):
    """terminal backward of a bias + LeakyReLU conv epilogue: (gy * lrelu'(y), column sums of it).  With ``sink`` the
    column sums are added to it and None is returned for them."""
    y, gy = _c(y), _c(gy)
    C = y.shape[-1]
    out = torch.empty_like(gy)
    gb = sink if sink is not None else torch.empty(C, dtype=torch.float32, device=y.device)
    call("gg_lrelu_bwd_bias", _p(y), _p(gy), _p(out), _p(gb), y.numel() // C, C, int(sink is not None), _dt(y), _st())
    return out, (None if sink is not None else gb)


# ---- side stream for weight gradients.  dW is off the critical path of a backward pass (only the optimiser needs it),
# while the data gradient chain is a sequence of persistent one-CTA-per-SM kernels whose tails, under-filled grids and
# small pointwise launches leave SMs idle: the terminal weight-gradient launches therefore go to a second stream (also
# inside CUDA-graph captures: fork/join through events) and fill those holes.  The join is queued as an autograd-engine
# callback, i.e. it runs at the end of the backward pass that forked.
SIDE_STREAM = {"on": os.environ.get("GG_SIDE_STREAM", "1") != "0", "stream": {}, "pending": [], "armed": False}


def _side_stream(device):
    if not SIDE_STREAM["on"]:
        return None
    st = SIDE_STREAM["stream"].get(device)
    if st is None:
        st = SIDE_STREAM["stream"][device] = torch.cuda.Stream(device=device)
    return st


def join_side_stream():
    """the current stream waits for everything forked to the side stream; operands kept alive for it are released"""
    SIDE_STREAM["armed"] = False
    pend, SIDE_STREAM["pending"] = SIDE_STREAM["pending"], []
    if pend:
        dev = pend[0][0].device
        torch.cuda.current_stream(dev).wait_stream(SIDE_STREAM["stream"][dev])
    del pend


def _wgrad_to_sink(x, gy, geom, weight):
    """Terminal (no higher-order graph) weight gradient of a parameter whose .grad is a view of an optimiser's flat
    gradient buffer (FlatAdamW marks those with ``_gg_sink``): run the wgrad kernel and accumulate its kernel-layout
    result straight into that buffer.  Returns False when the ordinary autograd route must be used."""
    if torch.is_grad_enabled() or not getattr(weight, "_gg_sink", False):
        return False
    dst = weight.grad
    if dst is None or dst.dtype != torch.float32 or not dst.is_contiguous():
        return False
    x, gy = _c(x), _c(gy)
    side = _side_stream(x.device) if _PROFILER[0] is None else None
    if side is None:
        dw = _conv_wgrad_raw(x, gy, geom)                           # (O, KH, KW, Ipad) fp32
        O, KH, KW, ipad = dw.shape
        call("gg_wgrad_sink", _p(dw), _p(dst), O, weight.shape[1], KH * KW, ipad, _st())
        return True
    side.wait_stream(torch.cuda.current_stream(x.device))           # x, gy (and the zeroed gradient buffer) are ready
    with torch.cuda.stream(side):
        dw = _conv_wgrad_raw(x, gy, geom)
        O, KH, KW, ipad = dw.shape
        call("gg_wgrad_sink", _p(dw), _p(dst), O, weight.shape[1], KH * KW, ipad, _st())
    SIDE_STREAM["pending"].append((x, gy, dw))      # alive until the join: their memory must not be reused before it
    if not SIDE_STREAM["armed"]:
        SIDE_STREAM["armed"] = True
        torch.autograd.Variable._execution_engine.queue_callback(join_side_stream)
    return True


class ConvWgradFn(Function):
    """dW from (x, gy).  Returns master layout fp32 when ``like`` (the master weight) is given, else kernel layout
    in x's dtype (per-sample prepared weights)."""

    @staticmethod
    def forward(ctx, x, gy, geom, like):
        x, gy = _c(x), _c(gy)
        ctx.geom, ctx.master = geom, like is not None
        ctx.save_for_backward(x, gy)
        ctx.like_shape = None if like is None else tuple(like.shape)
        dw = _conv_wgrad_raw(x, gy, geom)
        return unprep_weight_grad(dw, like) if like is not None else dw.to(x.dtype)

    @staticmethod
    def backward(ctx, ggw):
        x, gy = ctx.saved_tensors
        g = ctx.geom
        dx = dgy = None
        if ctx.needs_input_grad[0]:
            dx = ConvDgradFn.apply(gy, ggw, g, tuple(x.shape), ctx.master)
        if ctx.needs_input_grad[1]:
            dgy = Conv2dFn.apply(x, ggw, None, None, g, ctx.master)
        return dx, dgy, None, None


_SKIP_PARAM_GRADS = [False]


def _skip_param_grads():
    return _SKIP_PARAM_GRADS[0]


class skip_param_grads:
    """Context: convolutions do not compute weight/bias gradients (used while taking d(outputs)/d(images) for the
    gradient penalty, where autograd.grad only asks for the image gradient; ref gigagan_pytorch.py:138-145)."""

    def __enter__(self):
        self.prev = _SKIP_PARAM_GRADS[0]
        _SKIP_PARAM_GRADS[0] = True

    def __exit__(self, *a):
        _SKIP_PARAM_GRADS[0] = self.prev


def conv2d(x, weight, bias=None, *, stride=1, pad=0, act=0, res=None, gain=1.0):
    """NHWC convolution with the fp32 master weight (O,I,KH,KW)."""
    g = ConvGeom(weight.shape[2], weight.shape[3], stride, pad, False, act, gain)
    return Conv2dFn.apply(x, weight, bias, res, g, True)


def conv2d_prepared(x, wk, *, pad=0, per_sample=False):
    g = ConvGeom(wk.shape[-3], wk.shape[-2], 1, pad, per_sample)
    return Conv2dFn.apply(x, wk, None, None, g, False)


# ============================================================================= batched GEMM
def _strides4(t):
    return (torch.tensor if False else list)(t.stride())


def _bmm_raw(a, b, bias=None, alpha=1.0, out_bmhn=False):
    """The batched-GEMM kernel call itself (no autograd): C = alpha * A @ B (+ bias).  ``out_bmhn``: a tensor whose
    strides the result takes, True for a physical (b1, M, b2, N) layout, False for a contiguous result."""
    import ctypes
    b1, b2, m, k = a.shape
    n = b.shape[-1]
    assert b.shape[:3] == (b1, b2, k) and a.dtype == b.dtype
    if isinstance(out_bmhn, torch.Tensor):      # lay the result out like this tensor (same shape, last dim dense)
        c = torch.empty_strided((b1, b2, m, n), out_bmhn.stride(), dtype=a.dtype, device=a.device)
    elif out_bmhn:
        c = torch.empty((b1, m, b2, n), dtype=a.dtype, device=a.device).permute(0, 2, 1, 3)
    else:
        c = torch.empty((b1, b2, m, n), dtype=a.dtype, device=a.device)
    sa = (ctypes.c_int64 * 4)(*a.stride())
    sb = (ctypes.c_int64 * 4)(*b.stride())
    sc = (ctypes.c_int64 * 3)(*c.stride()[:3])
    assert c.stride(3) == 1
    call("gg_bmm", _p(a), _p(b), _p(bias), _p(c), b1, b2, m, n, k, ctypes.cast(sa, ctypes.c_void_p),
         ctypes.cast(sb, ctypes.c_void_p), ctypes.cast(sc, ctypes.c_void_p), float(alpha), _dt(a), _st())
    return c


class BmmFn(Function):
    """C[b1,b2] = alpha * A[b1,b2] @ B[b1,b2] (+ bias over the last axis).  A (b1,b2,M,K), B (b1,b2,K,N): any
    strides.  ``out_bmhn``: lay C out physically as (b1, M, b2, N) (returned view is still (b1,b2,M,N))."""

    @staticmethod
    def forward(ctx, a, b, bias, alpha, out_bmhn):
        c = _bmm_raw(a, b, bias, alpha, out_bmhn)
        ctx.alpha, ctx.has_bias = alpha, bias is not None
        ctx.save_for_backward(a, b)
        return c

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        ga = gb = gbias = None
        # gradients are produced directly in the memory layout of their operand (views of (batch, tokens, heads, d)
        # activations), so that autograd's view-backward / accumulation never needs a strided copy
        if ctx.needs_input_grad[0]:
            if _dense_like(a):
                ga = BmmFn.apply(g, b.transpose(-1, -2), None, ctx.alpha, a)
            elif _dense_like(a.transpose(-1, -2)):
                ga = BmmFn.apply(b, g.transpose(-1, -2), None, ctx.alpha, a.transpose(-1, -2)).transpose(-1, -2)
            else:
                ga = BmmFn.apply(g, b.transpose(-1, -2), None, ctx.alpha, False)
        if ctx.needs_input_grad[1]:
            if _dense_like(b):
                gb = BmmFn.apply(a.transpose(-1, -2), g, None, ctx.alpha, b)
            elif _dense_like(b.transpose(-1, -2)):
                gb = BmmFn.apply(g.transpose(-1, -2), a, None, ctx.alpha, b.transpose(-1, -2)).transpose(-1, -2)
            else:
                gb = BmmFn.apply(a.transpose(-1, -2), g, None, ctx.alpha, False)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gc = _c(g)
            gbias = dot_sc(gc, None, gc.numel() // gc.shape[-1], 1).reshape(-1)
        return ga, gb, gbias, None, None


def _dense_like(t):
    """True when t (b1,b2,M,N) has a unit last stride and is a non-overlapping permutation of a dense tensor, so a
    result can be written with exactly its strides."""
    if t.stride(-1) != 1 or 0 in t.stride():
        return False
    dims = sorted(range(4), key=lambda i: t.stride(i))
    expect = 1
    for i in dims:
        if t.shape[i] == 1:
            continue
        if t.stride(i) != expect:
            return False
        expect *= t.shape[i]
    return True


def bmm(a, b, alpha=1.0, out_bmhn=False):
    return BmmFn.apply(a, b, None, alpha, out_bmhn)


class RowLinearFn(Function):
    """y[r] = x[r,:] . w + b for ONE output channel (the logit heads), fp32 output (R,1).  First-order only.  The weight /
    bias gradients are added straight into the flat gradient buffer when ``w_param`` / ``bias`` are sink parameters."""

    @staticmethod
    def forward(ctx, x2d, weight, bias, w_param):
        x = _c(x2d)
        R, K = x.shape
        w = _c(weight.detach().reshape(-1).float())
        b = None if bias is None else _c(bias.detach().reshape(-1).float())
        y = torch.empty((R, 1), dtype=torch.float32, device=x.device)
        call("gg_row_linear_fwd", _p(x), _p(w), _p(b), _p(y), R, K, _dt(x), _st())
        ctx.save_for_backward(x, w)
        ctx.wshape, ctx.w_param = tuple(weight.shape), w_param
        ctx.bias_ref = bias if (bias is not None and getattr(bias, "_gg_sink1", False)) else None
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        R, K = x.shape
        gy = _c(gy.float().reshape(-1))
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dw = db = dw_ret = db_ret = None
        if ctx.needs_input_grad[1] and not _skip_param_grads():
            wp = ctx.w_param
            sink = wp.grad if (wp is not None and getattr(wp, "_gg_sink", False) and wp.grad is not None
                               and wp.grad.dtype == torch.float32 and wp.grad.is_contiguous() and wp.numel() == K) else None
            dw = sink if sink is not None else torch.zeros(K, dtype=torch.float32, device=x.device)
            dw_ret = None if sink is not None else dw.reshape(ctx.wshape)
        if ctx.has_bias and ctx.needs_input_grad[2] and not _skip_param_grads():
            sink = _bias_sink(ctx.bias_ref)
            db = sink if sink is not None else torch.zeros(1, dtype=torch.float32, device=x.device)
            db_ret = None if sink is not None else db
        if dx is not None or dw is not None or db is not None:
            call("gg_row_linear_bwd", _p(x), _p(w), _p(gy), _p(dx), _p(dw), _p(db), R, K, _dt(x), _st())
        return dx, dw_ret, db_ret, None


def _row_linear_ok(x2d, weight):
    K = x2d.shape[-1]
    v = 8 if x2d.dtype == torch.bfloat16 else 4
    nv = K // v
    return weight.shape[0] == 1 and K % v == 0 and 0 < nv <= 256 and (nv & (nv - 1)) == 0


def linear_rows(x2d, weight, bias=None, fused=False, w_param=None):
    """(R,K) activations (compute dtype) @ fp32 master weight (O,K)^T + bias -> (R,O).  ``fused`` (first-order passes,
    one output channel): a bandwidth kernel pair instead of a padded batched GEMM.  Otherwise in bf16 the output width is
    padded to 16 so the product runs on the tcgen05 batched GEMM."""
    O, K = weight.shape
    if fused and _row_linear_ok(x2d, weight):
        return RowLinearFn.apply(x2d, weight, bias, w_param)
    if x2d.dtype == torch.float32:
        return linear(x2d, weight, bias)
    opad = (O + 15) // 16 * 16
    w = torch.nn.functional.pad(weight, (0, 0, 0, opad - O)).to(x2d.dtype)
    b = None if bias is None else torch.nn.functional.pad(bias.float(), (0, opad - O))
    y = BmmFn.apply(x2d[None, None], w.t()[None, None], b, 1.0, False)[0, 0]
    return y[:, :O]


def linear(x, weight, bias=None, alpha=1.0):
    """x (R,K) @ weight(O,K)^T * alpha + bias  (fp32 small-tensor path; bias not scaled by alpha)."""
    y = BmmFn.apply(x[None, None], weight.t()[None, None], bias, alpha, False)
    return y[0, 0]


# ============================================================================= pointwise with derivative levels
class UnaryFn(Function):
    @staticmethod
    def forward(ctx, kind, x):
        x = _c(x)
        out = torch.empty_like(x)
        call("gg_pw_unary", kind, 0, _p(x), None, None, _p(out), x.numel(), _dt(x), _st())
        ctx.kind = kind
        ctx.save_for_backward(x)
        return out

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return None, Unary1Fn.apply(ctx.kind, x, g)


class Unary1Fn(Function):
    """g * f'(x)"""

    @staticmethod
    def forward(ctx, kind, x, g):
        x, g = _c(x), _c(g)
        out = torch.empty_like(x)
        call("gg_pw_unary", kind, 1, _p(x), _p(g), None, _p(out), x.numel(), _dt(x), _st())
        ctx.kind = kind
        ctx.save_for_backward(x, g)
        return out

    @staticmethod
    def backward(ctx, gg):
        x, g = ctx.saved_tensors
        dx = dg = None
        if ctx.needs_input_grad[1] and ctx.kind not in (U_LRELU, U_RELU):
            dx = Unary2Fn.apply(ctx.kind, x, g, gg)
        if ctx.needs_input_grad[2]:
            dg = Unary1Fn.apply(ctx.kind, x, gg)
        return None, dx, dg


class Unary2Fn(Function):
    """a * b * f''(x)  (terminal: third derivatives are never needed)"""

    @staticmethod
    def forward(ctx, kind, x, a, b):
        x, a, b = _c(x), _c(a), _c(b)
        out = torch.empty_like(x)
        call("gg_pw_unary", kind, 2, _p(x), _p(a), _p(b), _p(out), x.numel(), _dt(x), _st())
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        raise RuntimeError("third-order derivatives are not implemented")


def unary(kind, x):
    return UnaryFn.apply(kind, x)


def leaky_relu(x):
    return UnaryFn.apply(U_LRELU, x)


class MulFn(Function):
    @staticmethod
    def forward(ctx, a, b):
        a, b = _c(a), _c(b)
        out = torch.empty_like(a)
        call("gg_pw_mul", _p(a), _p(b), _p(out), a.numel(), _dt(a), _st())
        ctx.save_for_backward(a, b)
        return out

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        return (MulFn.apply(g, b) if ctx.needs_input_grad[0] else None,
                MulFn.apply(g, a) if ctx.needs_input_grad[1] else None)


def mul(a, b):
    return MulFn.apply(a, b)


_AXPBY_PASS = os.environ.get("GG_AXPBY_PASS", "1") != "0"      # diagnostics: 0 restores the copying backward


class AxpbyFn(Function):
    @staticmethod
    def forward(ctx, alpha, x, beta, y):
        x = _c(x)
        y = None if y is None else _c(y)
        out = torch.empty_like(x)
        call("gg_pw_axpby", float(alpha), _p(x), float(beta), _p(y), _p(out), x.numel(), _dt(x), _st())
        ctx.alpha, ctx.beta = alpha, beta
        ctx.yshape = None if y is None else y.shape
        return out

    @staticmethod
    def backward(ctx, g):
        # a unit coefficient passes the gradient through as it is (x + y is the common case: two full copies of g otherwise)
        gx = gy = None
        if ctx.needs_input_grad[1]:
            gx = g if (ctx.alpha == 1.0 and _AXPBY_PASS) else AxpbyFn.apply(ctx.alpha, g, 0.0, None)
        if ctx.needs_input_grad[3]:
            gy = g if (ctx.beta == 1.0 and _AXPBY_PASS) else AxpbyFn.apply(ctx.beta, g, 0.0, None)
            if gy.shape != ctx.yshape:
                gy = gy.reshape(ctx.yshape)
        return None, gx, None, gy


def axpby(alpha, x, beta=0.0, y=None):
    return AxpbyFn.apply(alpha, x, beta, y)


def add(x, y):
    return AxpbyFn.apply(1.0, x, 1.0, y)


# ============================================================================= broadcasts and their reductions
class BcastFn(Function):
    """x viewed [R, C] (C = last dim) combined with an fp32 statistic tensor s.
    mode ROWS: s has R entries.  mode SAMPLE_CH: s is (Ns, C); row r belongs to sample (r // P) % Ns."""

    @staticmethod
    def forward(ctx, x, s, P, Ns, mode, op):
        x, s = _c(x), _c(s)
        assert s.dtype == torch.float32
        C = x.shape[-1]
        R = x.numel() // C
        assert s.numel() == (R if mode == ROWS else Ns * C), (s.shape, R, Ns, C)
        out = torch.empty_like(x)
        call("gg_pw_bcast", _p(x), _p(s), _p(out), R, C, P, Ns, mode, op, _dt(x), _st())
        ctx.cfg = (P, Ns, mode, op)
        ctx.save_for_backward(x, s)
        return out

    @staticmethod
    def backward(ctx, g):
        x, s = ctx.saved_tensors
        P, Ns, mode, op = ctx.cfg
        gx = gs = None
        if ctx.needs_input_grad[0]:
            gx = BcastFn.apply(g, s, P, Ns, mode, MUL) if op == MUL else g
        if ctx.needs_input_grad[1]:
            other = x if op == MUL else None
            if mode == ROWS:
                gs = RowDotFn.apply(g, other if other is not None else _ones_like(g)).reshape(s.shape)
            else:
                gs = DotSCFn.apply(g, other, P, Ns).reshape(s.shape)
        return gx, gs, None, None, None, None


def _ones_like(t):
    return torch.ones_like(t)


def scale_rows(x, s):
    return BcastFn.apply(x, s, 1, 1, ROWS, MUL)


def scale_channels(x, s, rows_per_sample, num_samples):
    """x * s[(sample % num_samples), c]  (s fp32 (num_samples, C))"""
    return BcastFn.apply(x, s, rows_per_sample, num_samples, SAMPLE_CH, MUL)


def add_channels(x, s, rows_per_sample, num_samples):
    return BcastFn.apply(x, s, rows_per_sample, num_samples, SAMPLE_CH, ADD)


class RowDotFn(Function):
    """out[r] = sum_c a[r,c] b[r,c]   (fp32)"""

    @staticmethod
    def forward(ctx, a, b):
        a, b = _c(a), _c(b)
        C = a.shape[-1]
        R = a.numel() // C
        out = torch.empty(a.shape[:-1], dtype=torch.float32, device=a.device)
        call("gg_red_rowdot", _p(a), _p(b), _p(out), R, C, _dt(a), _st())
        ctx.save_for_backward(a, b)
        return out

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        g = g.float()
        return (scale_rows(b, g) if ctx.needs_input_grad[0] else None,
                scale_rows(a, g) if ctx.needs_input_grad[1] else None)


def rowdot(a, b):
    return RowDotFn.apply(a, b)


class DotSCFn(Function):
    """out[n', c] = sum over samples n == n' (mod Ns) and their P rows of a*b (b optional) -> fp32 (Ns, C)"""

    @staticmethod
    def forward(ctx, a, b, P, Ns):
        a = _c(a)
        b = None if b is None else _c(b)
        C = a.shape[-1]
        R = a.numel() // C
        assert R % (P * Ns) == 0
        out = torch.empty((Ns, C), dtype=torch.float32, device=a.device)
        call("gg_red_dot_sc", _p(a), _p(b), _p(out), R, C, P, Ns, _dt(a), _st())
        ctx.cfg = (P, Ns, b is None)
        ctx.save_for_backward(a, b)
        return out

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        P, Ns, no_b = ctx.cfg
        g = g.float()
        if no_b:
            ga = BcastFn.apply(torch.zeros_like(a), g, P, Ns, SAMPLE_CH, ADD) if ctx.needs_input_grad[0] else None
            return ga, None, None, None
        return (BcastFn.apply(b, g, P, Ns, SAMPLE_CH, MUL) if ctx.needs_input_grad[0] else None,
                BcastFn.apply(a, g, P, Ns, SAMPLE_CH, MUL) if ctx.needs_input_grad[1] else None, None, None)


def dot_sc(a, b, rows_per_sample, num_samples):
    return DotSCFn.apply(a, b, rows_per_sample, num_samples)


def mean_hw(x):
    """(N,H,W,C) -> fp32 (N,C) spatial mean."""
    n, h, w, c = x.shape
    return axpby(1.0 / (h * w), dot_sc(x, None, h * w, n))


def sum_all(x):
    """scalar fp32 sum of any tensor (large tensors: 64 column sums with vector loads first, then their sum)"""
    n = x.numel()
    if n >= 4096 and n % 64 == 0:
        x = dot_sc(x.reshape(-1, 64), None, n // 64, 1)
        n = 64
    return dot_sc(x.reshape(-1, 1), None, n, 1).reshape(())


class SoftmaxFn(Function):
    """softmax over the last axis of (s + bias); bias fp32 (Ns, C) indexed by (row // P) % Ns (optional)."""

    @staticmethod
    def forward(ctx, s, bias, P, Ns):
        s = _c(s)
        C = s.shape[-1]
        p = torch.empty_like(s)
        if bias is not None:
            bias = _c(bias)
            assert bias.dtype == torch.float32 and bias.numel() == Ns * C
        call("gg_softmax_rows", _p(s), _p(bias), _p(p), s.numel() // C, C, P, Ns, _dt(s), _st())
        ctx.save_for_backward(p)
        ctx.cfg = (P, Ns, bias is not None, None if bias is None else tuple(bias.shape))
        return p

    @staticmethod
    def backward(ctx, gp):
        (p,) = ctx.saved_tensors
        P, Ns, has_bias, bshape = ctx.cfg
        ds = SoftmaxBwdFn.apply(p, gp)
        gb = None
        if has_bias and ctx.needs_input_grad[1]:
            gb = dot_sc(ds, None, P, Ns).reshape(bshape)
        return ds, gb, None, None


class SoftmaxBwdFn(Function):
    """dS = P * (gP - rowdot(P, gP)) in one pass; its own backward is composed from closed primitives."""

    @staticmethod
    def forward(ctx, p, gp):
        p, gp = _c(p), _c(gp)
        C = p.shape[-1]
        ds = torch.empty_like(p)
        call("gg_softmax_bwd_rows", _p(p), _p(gp), _p(ds), p.numel() // C, C, _dt(p), _st())
        ctx.save_for_backward(p, gp)
        return ds

    @staticmethod
    def backward(ctx, G):
        p, gp = ctx.saved_tensors
        d_p = d_gp = None
        C = p.shape[-1]
        V = 4 if p.dtype == torch.float32 else 8
        if (not torch.is_grad_enabled()) and C % V == 0 and C // V <= 160 and all(ctx.needs_input_grad):
            G = _c(G)
            d_p, d_gp = torch.empty_like(p), torch.empty_like(p)       # terminal (third order is never needed)
            call("gg_softmax_bwd2_rows", _p(p), _p(gp), _p(G), _p(d_p), _p(d_gp), p.numel() // C, C, _dt(p), _st())
            return d_p, d_gp
        if ctx.needs_input_grad[1]:
            d_gp = SoftmaxBwdFn.apply(p, G)
        if ctx.needs_input_grad[0]:
            # d/dp_k = G_k (gp_k - r) - gp_k * rowdot(G, p),  r = rowdot(p, gp)
            r = rowdot(p, gp)
            t = axpby(1.0, mul(G, gp), -1.0, scale_rows(G, r))
            d_p = axpby(1.0, t, -1.0, scale_rows(gp, rowdot(G, p)))
        return d_p, d_gp


def softmax(s, bias=None, rows_per_sample=1, num_samples=1):
    return SoftmaxFn.apply(s, bias, rows_per_sample, num_samples)


# ============================================================================= any-order attention node
# Attention for the gradient-penalty pass as ONE autograd node with a hand-written first and second derivative.  The same
# kernels as the composed form (batched GEMM, row softmax and its first / second order backward), but the
# (tokens x keys)-sized traffic of the double backward drops from ~40 to ~30 passes per layer:
#   * the two products that make up d/d(dS) (u_q k^T + q u_k^T) are ONE GEMM over concatenated K axes;
#   * the probabilities' three gradient contributions (second-order softmax term, dO u_v^T from dV = P^T dO, and u_o v^T
#     from O = P V) are never summed in memory: the two rank-d terms are one GEMM over concatenated K axes, the softmax
#     term is added on the fly inside the softmax-backward kernel (gg_softmax_bwd_rows_add);
#   * no autograd accumulation (ATen add) on any (tokens x keys) tensor.
# Raw kernel calls are module-level callables so that tests/ can check the derivative plumbing against torch autograd
# with stand-ins on a machine without a GPU; the product path has no such stand-ins.
def _k_softmax(s, bias, P, Ns):
    s = _c(s)
    C = s.shape[-1]
    p = torch.empty_like(s)
    call("gg_softmax_rows", _p(s), _p(bias), _p(p), s.numel() // C, C, P, Ns, _dt(s), _st())
    return p


def _k_softmax_bwd(p, gp, gp2=None):
    p, gp = _c(p), _c(gp)
    C = p.shape[-1]
    ds = torch.empty_like(p)
    if gp2 is None:
        call("gg_softmax_bwd_rows", _p(p), _p(gp), _p(ds), p.numel() // C, C, _dt(p), _st())
    else:
        call("gg_softmax_bwd_rows_add", _p(p), _p(gp), _p(_c(gp2)), _p(ds), p.numel() // C, C, _dt(p), _st())
    return ds


def _k_softmax_bwd2(p, gp, G):
    """(d_p, d_gp) of dS = p * (gp - rowdot(p, gp)) for the upstream G; one pass when the row fits the kernel."""
    C = p.shape[-1]
    V = 4 if p.dtype == torch.float32 else 8
    if C % V == 0 and C // V <= 160:
        p, gp, G = _c(p), _c(gp), _c(G)
        d_p, d_gp = torch.empty_like(p), torch.empty_like(p)
        call("gg_softmax_bwd2_rows", _p(p), _p(gp), _p(G), _p(d_p), _p(d_gp), p.numel() // C, C, _dt(p), _st())
        return d_p, d_gp
    with torch.no_grad():
        d_gp = _k_softmax_bwd(p, G)
        r = rowdot(p, gp)
        t = axpby(1.0, mul(G, gp), -1.0, scale_rows(G, r))
        d_p = axpby(1.0, t, -1.0, scale_rows(gp, rowdot(G, p)))
    return d_p, d_gp


def _k_bmm(a, b, alpha=1.0, out=False):
    return _bmm_raw(a, b, None, alpha, out)


def _t(x):
    return x.transpose(-1, -2)


def _cat_last(a, b):
    """cat along the last axis of two (b, h, n, d) tensors.  Attention operands are views of (b, n, h, d) storage: the copy
    then runs in that layout (contiguous vector copies; the generic strided gather is ~6x slower) and the result is the
    same kind of view."""
    ap, bp = a.permute(0, 2, 1, 3), b.permute(0, 2, 1, 3)
    if ap.is_contiguous() and bp.is_contiguous():
        return torch.cat((ap, bp), dim=-1).permute(0, 2, 1, 3)
    return torch.cat((a, b), dim=-1)


def _attn_first_order(qa, ka, v, p, go, alpha, gp_extra=None, lowrank=None):
    """dqa, dka, dv of o = softmax(alpha qa ka^T) v for the cotangent go; gp_extra / lowrank: further gradients of the
    probabilities ((tokens x keys) tensor added inside the softmax backward; (a, b) meaning a b^T folded into the dP GEMM)."""
    if lowrank is None:
        dP = _k_bmm(go, _t(v))
    else:
        a2, b2 = lowrank
        dP = _k_bmm(_cat_last(go, a2), _t(_cat_last(v, b2)))
    dv = _k_bmm(_t(p), go, 1.0, v if _dense_like(v) else False)
    dS = _k_softmax_bwd(p, dP, gp_extra)
    dqa = _k_bmm(dS, ka, alpha, qa if _dense_like(qa) else False)
    dka = _k_bmm(_t(dS), qa, alpha, ka if _dense_like(ka) else False)
    return dqa, dka, dv, dP, dS


class ComposedAttnFn(Function):
    """(o, p) = attention(qa, ka, v): p = softmax(alpha * qa @ ka^T + mask), o = p @ v.  qa (b, h, n, D), ka (b, h, m, D),
    v (b, h, m, d), any strides; mask fp32 (1, m) or None (constant).  o is laid out physically as (b, n, h, d).  p is
    returned so that it is part of the graph (the second-order node differentiates through it); callers ignore it."""

    @staticmethod
    def forward(ctx, qa, ka, v, mask, alpha, holder):
        s = _k_bmm(qa, _t(ka), alpha)
        p = _k_softmax(s, mask, s.numel() // s.shape[-1], 1) if mask is not None else _k_softmax(s, None, 1, 1)
        del s
        o = _k_bmm(p, v, 1.0, True)
        ctx.save_for_backward(qa, ka, v, p)
        ctx.alpha, ctx.holder = alpha, holder
        ctx.set_materialize_grads(False)
        return o, p

    @staticmethod
    def backward(ctx, go, gp):
        qa, ka, v, p = ctx.saved_tensors
        if go is None:
            go = torch.zeros((qa.shape[0], qa.shape[2], qa.shape[1], v.shape[-1]), dtype=v.dtype,
                             device=v.device).permute(0, 2, 1, 3)
        if torch.is_grad_enabled():          # create_graph=True: the first derivative as a differentiable node
            assert gp is None
            dqa, dka, dv = ComposedAttnBwdFn.apply(qa, ka, v, p, go, ctx.alpha, ctx.holder)
        else:
            lowrank = ctx.holder.pop("lowrank", None)
            dqa, dka, dv, _, _ = _attn_first_order(qa, ka, v, p, go, ctx.alpha, gp, lowrank)
        return dqa, dka, dv, None, None, None


class ComposedAttnBwdFn(Function):
    """The first derivative of ComposedAttnFn as a node: (dqa, dka, dv) from (qa, ka, v, p, go); its backward is the
    second derivative (terminal: a third order is never needed)."""

    @staticmethod
    def forward(ctx, qa, ka, v, p, go, alpha, holder):
        dqa, dka, dv, dP, dS = _attn_first_order(qa, ka, v, p, go, alpha)
        ctx.save_for_backward(qa, ka, v, p, go, dP, dS)
        ctx.alpha, ctx.holder = alpha, holder
        ctx.set_materialize_grads(False)
        return dqa, dka, dv

    @staticmethod
    @once_differentiable
    def backward(ctx, uq, uk, uv):
        qa, ka, v, p, go, dP, dS = ctx.saved_tensors
        alpha = ctx.alpha
        g_qa = g_ka = g_v = g_p = g_go = None
        # d/d(dS) = alpha (uq ka^T + qa uk^T): one GEMM over the concatenated K axes
        if uq is not None and uk is not None:
            G = _k_bmm(_cat_last(uq, qa), _t(_cat_last(ka, uk)), alpha)
        elif uq is not None:
            G = _k_bmm(uq, _t(ka), alpha)
        elif uk is not None:
            G = _k_bmm(qa, _t(uk), alpha)
        else:
            G = None
        if G is not None:
            g_p, g_dP = _k_softmax_bwd2(p, dP, G)
            del G
            g_v = _k_bmm(_t(g_dP), go, 1.0, v if _dense_like(v) else False)          # dP = go v^T
            g_go = _k_bmm(g_dP, v, 1.0, True)      # physical (b, n, h, d): what the out-projection's convolutions read
            del g_dP
            if uk is not None:
                g_qa = _k_bmm(dS, uk, alpha, qa if _dense_like(qa) else False)       # dka = alpha dS^T qa
            if uq is not None:
                g_ka = _k_bmm(_t(dS), uq, alpha, ka if _dense_like(ka) else False)   # dqa = alpha dS ka
        if uv is not None:                                                             # dv = p^T go
            t = _k_bmm(p, uv, 1.0, True)
            if g_go is not None:                   # summed in the physical layout (contiguous there: no staging copies)
                t = axpby(1.0, g_go.permute(0, 2, 1, 3), 1.0, t.permute(0, 2, 1, 3)).permute(0, 2, 1, 3)
            g_go = t
            if g_p is None:
                g_p = _k_bmm(go, _t(uv))
            else:
                # go uv^T is a rank-d gradient of the probabilities: handed to the forward node, which folds it into
                # its dP GEMM (K axes concatenated) instead of a (tokens x keys)-sized accumulation here
                ctx.holder["lowrank"] = (go, uv)
        return g_qa, g_ka, g_v, g_p, g_go, None, None


# ---- operand builder of that node for the shared-QK L2 form (bf16, dim_head 64): one kernel each for the forward, the
# first and the second derivative instead of row reductions, casts, zero fills and three concatenations (and, backwards,
# strided slice copies and gradient accumulations of the key-sized tensors).
def _k_aug_fwd(q4, v4, null_kv, Lp):
    n, seq, h, d = q4.shape
    qa = torch.empty((n, seq, h, d + 16), dtype=q4.dtype, device=q4.device)
    ka = torch.empty((n, Lp, h, d + 16), dtype=q4.dtype, device=q4.device)
    vf = torch.empty((n, Lp, h, d), dtype=q4.dtype, device=q4.device)
    call("gg_attn_augment_fwd", _p(q4), _p(v4), _p(null_kv), _p(qa), _p(ka), _p(vf), n, seq, Lp, h, d, _st())
    return qa, ka, vf


def _k_aug_bwd(dqa, dka, dvf, q4, null_kv):
    n, seq, h, d = q4.shape
    dq, dv = torch.empty_like(q4), torch.empty_like(q4)
    dnull = torch.empty((2, h, d), dtype=torch.float32, device=q4.device)
    call("gg_attn_augment_bwd", _p(dqa), _p(dka), _p(dvf), _p(q4), _p(null_kv), _p(dq), _p(dv), _p(dnull), n, seq, dka.shape[1],
         h, d, _st())
    return dq, dv, dnull


def _k_aug_bwd2(wq, wv, wnull, q4, null_kv, dka):
    n, seq, h, d = q4.shape
    Lp = dka.shape[1]
    g_dqa = torch.empty((n, seq, h, d + 16), dtype=q4.dtype, device=q4.device)
    g_dka, g_dvf = torch.empty_like(dka), torch.empty((n, Lp, h, d), dtype=q4.dtype, device=q4.device)
    g_q = torch.empty_like(q4)
    g_null = torch.empty((2, h, d), dtype=torch.float32, device=q4.device)
    call("gg_attn_augment_bwd2", _p(wq), _p(wv), _p(wnull), _p(q4), _p(null_kv), _p(dka), _p(g_dqa), _p(g_dka), _p(g_dvf), _p(g_q),
         _p(g_null), n, seq, Lp, h, d, _st())
    return g_dqa, g_dka, g_dvf, g_q, g_null


class AttnAugmentFn(Function):
    """(qa, ka, vf) of the shared-QK L2 attention from q (= k) and v, both (n, seq, h, 64) contiguous, and the null
    key/value parameter (2, h, 64) fp32; see csrc/attn_augment.cu for the layouts."""

    @staticmethod
    def forward(ctx, q4, v4, null_kv, Lp):
        q4, v4, null_kv = _c(q4), _c(v4), _c(null_kv)
        ctx.save_for_backward(q4, null_kv)
        return _k_aug_fwd(q4, v4, null_kv, Lp)

    @staticmethod
    def backward(ctx, dqa, dka, dvf):
        q4, null_kv = ctx.saved_tensors
        if torch.is_grad_enabled():
            dq, dv, dnull = AttnAugmentBwdFn.apply(dqa, dka, dvf, q4, null_kv)
        else:
            dq, dv, dnull = _k_aug_bwd(_c(dqa), _c(dka), _c(dvf), q4, null_kv)
        return dq, dv, dnull, None


class AttnAugmentBwdFn(Function):
    """first derivative of AttnAugmentFn as a node (linear in the gradients, bilinear in (dka[..., 64], q)); its backward
    is terminal"""

    @staticmethod
    def forward(ctx, dqa, dka, dvf, q4, null_kv):
        dqa, dka, dvf = _c(dqa), _c(dka), _c(dvf)
        ctx.save_for_backward(q4, null_kv, dka)
        return _k_aug_bwd(dqa, dka, dvf, q4, null_kv)

    @staticmethod
    @once_differentiable
    def backward(ctx, wq, wv, wnull):
        q4, null_kv, dka = ctx.saved_tensors
        g_dqa, g_dka, g_dvf, g_q, g_null = _k_aug_bwd2(_c(wq), _c(wv), _c(wnull), q4, null_kv, dka)
        return g_dqa, g_dka, g_dvf, g_q, g_null


def attn_augment(q4, v4, null_kv, Lp):
    return AttnAugmentFn.apply(q4, v4, null_kv, Lp)


def composed_attention(qa, ka, v, mask, alpha):
    o, _ = ComposedAttnFn.apply(qa, ka, v, mask, alpha, {})
    return o


# ============================================================================= resampling (separable, sparse)
class ResampleOp:
    """A separable linear map on the (H, W) axes given by dense 1-D matrices Ay (OH,H), Ax (OW,W); holds device
    tap tables for the map and its transpose."""

    def __init__(self, ay, ax, device):
        self.shape_in = (ay.shape[1], ax.shape[1])
        self.shape_out = (ay.shape[0], ax.shape[0])
        self.fwd = (self._taps(ay, device), self._taps(ax, device))
        self.bwd = (self._taps(ay.t(), device), self._taps(ax.t(), device))

    @staticmethod
    def _taps(a, device):
        a = a.double()
        nnz = int((a != 0).sum(dim=1).max().item())
        idx = torch.zeros((a.shape[0], nnz), dtype=torch.int32)
        wgt = torch.zeros((a.shape[0], nnz), dtype=torch.float32)
        for r in range(a.shape[0]):
            cols = torch.nonzero(a[r]).flatten()
            idx[r, : len(cols)] = cols.to(torch.int32)
            wgt[r, : len(cols)] = a[r, cols].float()
        return idx.to(device).contiguous(), wgt.to(device).contiguous(), nnz


def bilinear_matrix(n_in, n_out):
    """1-D F.interpolate(mode='bilinear', align_corners=False) as an (n_out, n_in) matrix (torch semantics:
    src = (dst + 0.5) * n_in / n_out - 0.5, clamped below at 0, right neighbour clamped to n_in - 1)."""
    a = torch.zeros((n_out, n_in), dtype=torch.float64)
    scale = n_in / n_out
    for j in range(n_out):
        src = max((j + 0.5) * scale - 0.5, 0.0)
        i0 = min(int(math.floor(src)), n_in - 1)
        i1 = min(i0 + 1, n_in - 1)
        f = src - i0
        a[j, i0] += 1.0 - f
        a[j, i1] += f
    return a


def blur_matrix(n):
    """1-D [1,2,1]/4 with reflect border (kornia filter2d default), (n, n)."""
    a = torch.zeros((n, n), dtype=torch.float64)
    for j in range(n):
        for off, wt in ((-1, 0.25), (0, 0.5), (1, 0.25)):
            i = j + off
            if i < 0:
                i = -i
            if i >= n:
                i = 2 * (n - 1) - i
            a[j, i] += wt
    return a


_RESAMPLE_CACHE = {}


def get_resample_op(kind, h, w, device, out=None):
    key = (kind, h, w, str(device), out)
    op = _RESAMPLE_CACHE.get(key)
    if op is None:
        if kind == "up2_blur":      # ref gigagan_pytorch.py:257-261
            ay, ax = blur_matrix(2 * h) @ bilinear_matrix(h, 2 * h), blur_matrix(2 * w) @ bilinear_matrix(w, 2 * w)
        elif kind == "bilinear":    # ref gigagan_pytorch.py:1683-1684
            ay, ax = bilinear_matrix(h, out[0]), bilinear_matrix(w, out[1])
        elif kind == "blur":
            ay, ax = blur_matrix(h), blur_matrix(w)
        else:
            raise ValueError(kind)
        op = _RESAMPLE_CACHE[key] = ResampleOp(ay, ax, device)
    return op


class ResampleFn(Function):
    @staticmethod
    def forward(ctx, x, op, transposed):
        x = _c(x)
        n, h, w, c = x.shape
        (ty, tx), (oh, ow) = (op.bwd, op.shape_in) if transposed else (op.fwd, op.shape_out)
        assert (h, w) == (op.shape_out if transposed else op.shape_in)
        y = torch.empty((n, oh, ow, c), dtype=x.dtype, device=x.device)
        call("gg_resample2d", _p(x), _p(y), n, h, w, c, oh, ow, _p(ty[0]), _p(ty[1]), ty[2], _p(tx[0]), _p(tx[1]),
             tx[2], _dt(x), _st())
        ctx.op, ctx.transposed = op, transposed
        return y

    @staticmethod
    def backward(ctx, g):
        return ResampleFn.apply(g, ctx.op, not ctx.transposed), None, None


def upsample2x_blur(x):
    return ResampleFn.apply(x, get_resample_op("up2_blur", x.shape[1], x.shape[2], x.device), False)


def resize_bilinear(x, size):
    return ResampleFn.apply(x, get_resample_op("bilinear", x.shape[1], x.shape[2], x.device, (size, size)), False)


# ============================================================================= layout at the API edge
class ToNHWCFn(Function):
    """NCHW fp32 (reference layout) -> NHWC compute dtype, channels zero-padded to cpad."""

    @staticmethod
    def forward(ctx, x, cpad, dtype):
        x = _c(x.float())
        n, c, h, w = x.shape
        y = torch.empty((n, h, w, cpad), dtype=dtype, device=x.device)
        call("gg_nchw_to_nhwc", _p(x), _p(y), n, c, h * w, cpad, _dt(y), _st())
        ctx.c = c
        return y

    @staticmethod
    def backward(ctx, g):
        return ToNCHWFn.apply(g, ctx.c), None, None


class ToNCHWFn(Function):
    """NHWC (cpad channels) -> NCHW fp32 keeping the first c channels."""

    @staticmethod
    def forward(ctx, x, c):
        x = _c(x)
        n, h, w, cpad = x.shape
        y = torch.empty((n, c, h, w), dtype=torch.float32, device=x.device)
        call("gg_nhwc_to_nchw", _p(x), _p(y), n, c, h * w, cpad, _dt(x), _st())
        ctx.cpad, ctx.dtype = cpad, x.dtype
        return y

    @staticmethod
    def backward(ctx, g):
        return ToNHWCFn.apply(g, ctx.cpad, ctx.dtype), None


def to_nhwc(x, cpad, dtype):
    return ToNHWCFn.apply(x, cpad, dtype)


def to_nchw(x, c):
    return ToNCHWFn.apply(x, c)


# ============================================================================= generator-only fused ops (first-order)
def _rows(t):
    """fp32 2-D statistic (B, C) as the kernels read it: unit column stride, any row stride (column slices of the style
    projection are used in place, no copy)"""
    if t.dtype != torch.float32:
        t = t.float()
    if t.dim() != 2 or t.stride(1) != 1 or t.stride(0) < t.shape[1]:
        t = t.contiguous()
    return t


class AdaConvWeightsFn(Function):
    """Per-sample modulated / demodulated filter from the bank (ref gigagan_pytorch.py:378-400).
    bank (n,O,I,k,k) fp32; mod (B,I) fp32; kmod (B,n) fp32 or None -> (B,O,k,k,I) kernel layout, compute dtype."""

    @staticmethod
    def forward(ctx, bank, mod, kmod, demod, eps, dtype, opad=0):
        bank, mod = _c(bank), _rows(mod)
        n, O, I, k, _ = bank.shape
        B = mod.shape[0]
        opad = max(opad, O)
        if n > 1:
            assert kmod is not None and kmod.numel() > 0
            kmod = _rows(kmod)
        else:
            kmod = None
        w = (torch.empty if opad == O else torch.zeros)((B, opad, k, k, I), dtype=dtype, device=bank.device)
        attn = torch.empty((B, n), dtype=torch.float32, device=bank.device)
        dinv = torch.empty((B, O), dtype=torch.float32, device=bank.device)
        call("gg_adaconv_weights_fwd", _p(bank), _p(mod), _p(kmod), _p(w), _p(attn), _p(dinv), B, n, O, I, k * k,
             int(demod), float(eps), opad, mod.stride(0), 0 if kmod is None else kmod.stride(0), _dt(w), _st())
        ctx.cfg = (demod, eps, kmod is not None, opad)
        ctx.save_for_backward(bank, mod, attn, dinv)
        return w

    @staticmethod
    @once_differentiable
    def backward(ctx, gw):
        bank, mod, attn, dinv = ctx.saved_tensors
        demod, eps, has_kmod, opad = ctx.cfg
        n, O, I, k, _ = bank.shape
        B = mod.shape[0]
        gw = _c(gw.float())
        dbank = torch.empty_like(bank)
        dmod = torch.empty((B, I), dtype=torch.float32, device=bank.device)
        dkmod = torch.empty((B, n), dtype=torch.float32, device=bank.device) if has_kmod else None
        ws = torch.empty((B * n + B * O,), dtype=torch.float32, device=bank.device)
        call("gg_adaconv_weights_bwd", _p(bank), _p(mod), _p(attn), _p(dinv), _p(gw), _p(dbank), _p(dmod), _p(dkmod),
             _p(ws), B, n, O, I, k * k, int(demod), float(eps), opad, mod.stride(0), _st())
        return dbank, dmod, dkmod, None, None, None, None


class NoiseActFn(Function):
    """lrelu(x + weight[c] * noise[n,h,w])  (ref gigagan_pytorch.py:925-940 + :1222)"""

    @staticmethod
    def forward(ctx, x, noise, weight):
        x, noise, weight = _c(x), _c(noise.float()), _c(weight.float())
        C = x.shape[-1]
        y = torch.empty_like(x)
        call("gg_noise_act_fwd", _p(x), _p(noise), _p(weight), _p(y), x.numel() // C, C, _dt(x), _st())
        ctx.save_for_backward(y, noise)
        ctx.wshape = weight.shape
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        y, noise = ctx.saved_tensors
        gy = _c(gy)
        C = y.shape[-1]
        dx = torch.empty_like(y)
        dw = torch.empty(C, dtype=torch.float32, device=y.device)
        call("gg_noise_act_bwd", _p(y), _p(gy), _p(noise), _p(dx), _p(dw), y.numel() // C, C, _dt(y), _st())
        return dx, None, dw.reshape(ctx.wshape)


class FusedAttnFn(Function):
    """softmax(logits) V with online softmax; q,k,v (B, n, heads*d) rows (views into NHWC conv outputs allowed:
    last dim contiguous).  First-order only (the gradient-penalty path uses the composed attention)."""

    @staticmethod
    def forward(ctx, q, k, v, null_kv, heads, scale, l2, shared_qk):
        B, nq, hd = q.shape
        nk, d = k.shape[1], hd // heads
        assert q.stride(-1) == 1 and k.stride(-1) == 1 and v.stride(-1) == 1
        assert q.stride(0) == nq * q.stride(1) and k.stride(0) == nk * k.stride(1) and v.stride(0) == nk * v.stride(1)
        nkv = None if null_kv is None else _c(null_kv.float())
        o = torch.empty((B, nq, hd), dtype=q.dtype, device=q.device)
        lse = torch.empty((B * heads, nq), dtype=torch.float32, device=q.device)
        ws = torch.empty((B * heads * nk,), dtype=torch.float32, device=q.device) if l2 else None
        call("gg_attn_fwd", _p(q), _p(k), _p(v), _p(nkv), _p(o), _p(lse), _p(ws), B, heads, nq, nk, d, q.stride(1),
             k.stride(1), v.stride(1), o.stride(1), float(scale), int(l2), _dt(q), _st())
        ctx.cfg = (heads, scale, l2, shared_qk)
        ctx.save_for_backward(q, k, v, nkv, o, lse)
        return o

    @staticmethod
    @once_differentiable
    def backward(ctx, go):
        q, k, v, nkv, o, lse = ctx.saved_tensors
        heads, scale, l2, shared = ctx.cfg
        B, nq, hd = q.shape
        nk, d = k.shape[1], hd // heads
        go = _c(go)
        dq = torch.empty((B, nq, hd), dtype=q.dtype, device=q.device)
        dk = torch.empty((B, nk, hd), dtype=q.dtype, device=q.device)
        dv = torch.empty((B, nk, hd), dtype=q.dtype, device=q.device)
        dnull = None if nkv is None else torch.empty_like(nkv)
        delta = torch.empty((3 * lse.numel(),), dtype=torch.float32, device=q.device)   # delta | ds_null | p_null
        # the backward kernel indexes dq/dk/dv with the strides of q/k/v: give it dense copies' strides
        assert go.stride(1) == o.stride(1)
        qc, kc, vc = _c(q), _c(k), _c(v)
        ws = torch.empty((B * heads * nk,), dtype=torch.float32, device=q.device) if l2 else None
        call("gg_attn_bwd", _p(qc), _p(kc), _p(vc), _p(nkv), _p(o), _p(go), _p(lse), _p(dq), _p(dk), _p(dv), _p(dnull),
             _p(delta), _p(ws), B, heads, nq, nk, d, qc.stride(1), kc.stride(1), vc.stride(1), o.stride(1), float(scale),
             int(l2), _dt(q), _st())
        if shared:
            dq = add(dq, dk)
            dk = None
        return dq, dk, dv, dnull, None, None, None, None


def fused_attention(q, k, v, null_kv, heads, scale, l2=False):
    shared = k is q
    return FusedAttnFn.apply(q, k, v, null_kv, heads, scale, l2, shared)


class SharedBankConvFn(Function):
    """AdaptiveConv2DMod (ref gigagan_pytorch.py:378-409) of the 4x4 / 8x8 layers in shared-bank form, ONE autograd node:
        y_b = dinv_b (.) sum_n attn_bn conv(x_b * (mod_b + 1), W_n)
    forward: prep (attn, dinv, xs) -> one dense tcgen05 convolution over the concatenated bank [W_0; ..; W_{n-1}] whose
    128-row tiles span images -> combine;  backward: combine^T, n data-gradient convolutions accumulated through the
    residual epilogue, n weight-gradient launches into one [n][O][KK][I] buffer, the demodulation chain (which also
    folds the weight gradients into the bank's layout), and the input scaling.  First-order only (the generator is never
    inside the gradient penalty); ``AdaptiveConv2DMod._forward_shared_bank`` is the any-order composed form.
    x (B,H,W,I) compute dtype; bank (n,O,I,k,k) fp32 master; mod (B,I), kmod (B,n) fp32 (column slices allowed)."""

    @staticmethod
    def forward(ctx, x, bank, mod, kmod, eps):
        x, bank, mod = _c(x), _c(bank), _rows(mod)
        B, H, W, I = x.shape
        n, O, _, k, _ = bank.shape
        kmod = _rows(kmod) if n > 1 else None
        dev, dt = x.device, x.dtype
        wks = [prep_weight(bank[j], I, dt) for j in range(n)]            # kernel layout (O,k,k,I) each
        per = wks[0].numel() * wks[0].element_size()
        if all(wks[j].data_ptr() == wks[0].data_ptr() + j * per for j in range(n)) and n > 1 and wks[0]._base is not None:
            base = wks[0]._base                                          # views of the weight bank: already back to back
            off = (wks[0].data_ptr() - base.data_ptr()) // base.element_size()
            wcat = base[off:off + n * wks[0].numel()].view(n * O, k, k, I)
        else:
            wcat = wks[0] if n == 1 else torch.cat(wks, dim=0)
        xs = torch.empty_like(x)
        attn = torch.empty((B, n), dtype=torch.float32, device=dev)
        dinv = torch.empty((B, O), dtype=torch.float32, device=dev)
        call("gg_sbank_prep", _p(bank), _p(mod), _p(kmod), _p(x), _p(xs), _p(attn), _p(dinv), B, n, O, I, k * k, H * W, 1,
             float(eps), mod.stride(0), 0 if kmod is None else kmod.stride(0), _dt(x), _st())
        pad = (k - 1) // 2
        ycat = _conv_fprop_raw(xs, wcat, None, None, ConvGeom(k, k, 1, pad), n * O)
        y = torch.empty((B, H, W, O), dtype=dt, device=dev)
        call("gg_sbank_combine_fwd", _p(ycat), _p(attn), _p(dinv), _p(y), B, H * W, n, O, _dt(x), _st())
        ctx.cfg = (float(eps), kmod is not None)
        ctx.wks = wks
        ctx.save_for_backward(x, xs, ycat, attn, dinv, bank, mod)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        x, xs, ycat, attn, dinv, bank, mod = ctx.saved_tensors
        eps, has_kmod = ctx.cfg
        B, H, W, I = x.shape
        n, O, _, k, _ = bank.shape
        dev, dt = x.device, x.dtype
        gy = _c(gy)
        gyn = torch.empty((n, B, H, W, O), dtype=dt, device=dev)
        ws = torch.empty((B * n + B * O,), dtype=torch.float32, device=dev)      # gattn | (unused q slot)
        gdinv = torch.empty((B, O), dtype=torch.float32, device=dev)
        call("gg_sbank_combine_bwd", _p(gy), _p(ycat), _p(attn), _p(dinv), _p(gyn), _p(gdinv), _p(ws), B, H * W, n, O,
             _dt(x), _st())
        pad = (k - 1) // 2
        gt = ConvGeom(k, k, 1, k - 1 - pad)
        gxs = None
        for j in range(n):                      # d xs = sum_j conv^T(gy_j, W_j): accumulated by the residual epilogue
            gxs = _conv_fprop_raw(gyn[j], flipped_weight(ctx.wks[j]), None, gxs, gt, I)
        dwk = torch.empty((n, O, k, k, I), dtype=torch.float32, device=dev)
        with _Timed("wgrad", 2.0 * n * B * H * W * O * I * k * k, dt, f"wgrad sbank n{B} {H}x{W} {I}->{n}x{O} k{k}"):
            for j in range(n):
                call("gg_conv2d_wgrad", _p(xs), _p(gyn[j]), dwk[j].data_ptr(), B, H, W, I, H, W, O, k, k, 1, pad, 0,
                     _dt(x), _st())
        dbank = torch.empty_like(bank)
        dmod = torch.empty((B, I), dtype=torch.float32, device=dev)
        dkmod = torch.empty((B, n), dtype=torch.float32, device=dev) if has_kmod else None
        call("gg_sbank_bwd_stats", _p(bank), _p(mod), _p(attn), _p(dinv), _p(gdinv), _p(dwk), _p(dbank), _p(dmod),
             _p(dkmod), _p(ws), B, n, O, I, k * k, eps, mod.stride(0), _st())
        gx = torch.empty_like(x)
        call("gg_sbank_bwd_x", _p(gxs), _p(x), _p(mod), _p(gx), _p(dmod), B, H * W, I, mod.stride(0), _dt(x), _st())
        return gx, dbank, dmod, dkmod, None


def shared_bank_conv(x, bank, mod, kmod, eps):
    return SharedBankConvFn.apply(x, bank, mod, kmod, eps)


class PatchSelectFn(Function):
    """rows (b, s) of the output = patch sel[b, s] of image b (the auxiliary decoder's random patch subset, ref
    gigagan_pytorch.py:1300-1312).  t (B, pd*hh, pd*ww, C); sel int32 (B, nsel) on the device.  Linear: the backward is
    the adjoint scatter (also differentiable: it is this same operator with the roles swapped)."""

    @staticmethod
    def forward(ctx, t, sel, pd, transposed):
        t = _c(t)
        B, nsel = sel.shape
        assert sel.dtype == torch.int32 and sel.is_contiguous()
        if not transposed:
            _, H, W, C = t.shape
            hh, ww = H // pd, W // pd
            out = torch.empty((B * nsel, hh, ww, C), dtype=t.dtype, device=t.device)
        else:
            _, hh, ww, C = t.shape
            out = torch.empty((B, pd * hh, pd * ww, C), dtype=t.dtype, device=t.device)
        call("gg_patch_select", _p(t), _p(out), _p(sel), B, nsel, pd, hh, ww, C, int(transposed), _dt(t), _st())
        ctx.cfg = (pd, transposed)
        ctx.save_for_backward(sel)
        return out

    @staticmethod
    def backward(ctx, g):
        (sel,) = ctx.saved_tensors
        pd, transposed = ctx.cfg
        return PatchSelectFn.apply(g, sel, pd, not transposed), None, None, None


def patch_select(t, sel, pd):
    return PatchSelectFn.apply(t, sel, pd, False)


# ============================================================================= GAN objective (first-order)
class GanLossFn(Function):
    """Hinge objective over all logit tensors of one discriminator pass in ONE launch (ref gigagan_pytorch.py:159-163
    and the weighted sums at :2327-2347 / :2538-2551).  tensors[0] = main logits, the rest = multiscale logits; each is
    viewed as rows of ``rows[j]`` elements whose first ``splits[j]`` are real logits and the rest fake (mode 0), or
    averaged as a whole (mode 1, generator).  -> (total = L_0 + w_ms * sum_j L_j, L_0, sum_{j>=1} L_j); only ``total``
    is differentiable (the parts are logged).  First-order only: the loss is outside the gradient penalty's graph."""

    @staticmethod
    def forward(ctx, mode, w_ms, rows, splits, *tensors):
        import ctypes
        k = len(tensors)
        xs = [_c(t) for t in tensors]
        meta = (ctypes.c_int64 * (4 * k))()
        for j, t in enumerate(xs):
            meta[4 * j:4 * j + 4] = [t.numel(), rows[j], splits[j], _dt(t)]
        ptrs = (ctypes.c_void_p * k)(*[t.data_ptr() for t in xs])
        out = torch.empty(3, dtype=torch.float32, device=xs[0].device)
        call("gg_gan_loss_fwd", ctypes.cast(ptrs, ctypes.c_void_p), ctypes.cast(meta, ctypes.c_void_p), k, mode,
             float(w_ms), _p(out), _st())
        ctx.cfg = (mode, float(w_ms), k, meta)
        ctx.save_for_backward(*xs)
        total, l0, lms = out[2], out[0], out[1]
        ctx.mark_non_differentiable(l0, lms)
        return total, l0, lms

    @staticmethod
    @once_differentiable
    def backward(ctx, g, _g0, _g1):
        import ctypes
        mode, w_ms, k, meta = ctx.cfg
        xs = ctx.saved_tensors
        dxs = [torch.empty_like(t) for t in xs]
        ptrs = (ctypes.c_void_p * k)(*[t.data_ptr() for t in xs])
        dptrs = (ctypes.c_void_p * k)(*[t.data_ptr() for t in dxs])
        g = _c(g.float().reshape(1))
        call("gg_gan_loss_bwd", ctypes.cast(ptrs, ctypes.c_void_p), ctypes.cast(dptrs, ctypes.c_void_p),
             ctypes.cast(meta, ctypes.c_void_p), k, mode, w_ms, _p(g), _st())
        return (None, None, None, None, *dxs)


def gan_loss(mode, w_ms, tensors, rows, splits):
    return GanLossFn.apply(mode, w_ms, tuple(rows), tuple(splits), *tensors)


# ============================================================================= UnetUpsampler extras
class RMSNormFn(Function):
    """Fused ChannelRMSNorm over the last axis (first-order only; the composed rowdot/invnorm/broadcast chain in
    modules.channel_rmsnorm is the any-order differentiable form used on gradient-penalty steps)."""

    @staticmethod
    def forward(ctx, x, gamma, scale):
        x = _c(x)
        C = x.shape[-1]
        R = x.numel() // C
        g = _c(gamma.detach().reshape(-1).float())
        y = torch.empty_like(x)
        inv = torch.empty(R, dtype=torch.float32, device=x.device)
        call("gg_rmsnorm_fwd", _p(x), _p(g), _p(y), _p(inv), R, C, float(scale), _dt(x), _st())
        ctx.save_for_backward(x, g, inv)
        ctx.scale, ctx.gshape = float(scale), tuple(gamma.shape)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        x, g, inv = ctx.saved_tensors
        gy = _c(gy)
        C = x.shape[-1]
        gx = torch.empty_like(x)
        dg = torch.zeros(C, dtype=torch.float32, device=x.device)
        call("gg_rmsnorm_bwd", _p(x), _p(g), _p(inv), _p(gy), _p(gx), _p(dg), x.numel() // C, C, ctx.scale, _dt(x), _st())
        return gx, dg.reshape(ctx.gshape), None


def rmsnorm_fused_ok(x):
    c = x.shape[-1]
    return c % 8 == 0 and c <= (1024 if x.dtype == torch.bfloat16 else 512)


def rmsnorm_fused(x, gamma, scale):
    return RMSNormFn.apply(x, gamma, scale)


class MaxPool2Fn(Function):
    """2x2 max-pool of an NHWC map (first-order; the upsampler is never inside the gradient penalty)."""

    @staticmethod
    def forward(ctx, x):
        x = _c(x)
        n, h, w, c = x.shape
        y = torch.empty((n, h // 2, w // 2, c), dtype=x.dtype, device=x.device)
        call("gg_maxpool2_fwd", _p(x), _p(y), n, h, w, c, _dt(x), _st())
        ctx.save_for_backward(x)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        (x,) = ctx.saved_tensors
        gy = _c(gy)
        n, h, w, c = x.shape
        gx = torch.empty_like(x)
        call("gg_maxpool2_bwd", _p(x), _p(gy), _p(gx), n, h, w, c, _dt(x), _st())
        return gx


def maxpool2(x):
    return MaxPool2Fn.apply(x)


class SoftmaxTokensFn(Function):
    """softmax over the token axis of (B, n, C)."""

    @staticmethod
    def forward(ctx, x):
        x = _c(x)
        b, n, c = x.shape
        y = torch.empty_like(x)
        call("gg_softmax_tokens", _p(x), _p(y), b, n, c, _dt(x), _st())
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, g):
        (p,) = ctx.saved_tensors
        b, n, c = p.shape
        r = dot_sc(p, g, n, b)                                # (b, c) = sum_tokens p * g
        return axpby(1.0, mul(p, g), -1.0, scale_channels(p, r, n, b))


def softmax_tokens(x):
    return SoftmaxTokensFn.apply(x)
