"""Stand-in for kornia.filters.filter2d / filter3d (normalized, reflect border)."""
import torch
import torch.nn.functional as F


def _normalize(kernel):
    return kernel / kernel.abs().sum(dim=tuple(range(1, kernel.ndim)), keepdim=True)


def filter2d(x, kernel, border_type="reflect", normalized=False, padding="same"):
    if normalized:
        kernel = _normalize(kernel)
    b, c, h, w = x.shape
    kh, kw = kernel.shape[-2:]
    weight = kernel.to(x)[:, None].expand(c, 1, kh, kw) if kernel.shape[0] == 1 else kernel.to(x)[:, None]
    xp = F.pad(x, (kw // 2, kw // 2, kh // 2, kh // 2), mode=border_type)
    return F.conv2d(xp, weight.contiguous(), groups=c)


def filter3d(x, kernel, border_type="replicate", normalized=False):
    if normalized:
        kernel = _normalize(kernel)
    b, c, d, h, w = x.shape
    kd, kh, kw = kernel.shape[-3:]
    weight = kernel.to(x)[:, None].expand(c, 1, kd, kh, kw)
    xp = F.pad(x, (kw // 2, kw // 2, kh // 2, kh // 2, kd // 2, kd // 2), mode=border_type)
    return F.conv3d(xp, weight.contiguous(), groups=c)
