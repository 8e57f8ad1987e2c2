"""Stand-in for the un-installable `resize-right==0.0.2` (reference call sites: minimagen/helpers.py:159,
minimagen/training.py:168).  Only the inter-stage resize of the full cascade (a "next" row) needs it;
the per-step hot path never calls it.  The package source is NOT in this container, so this is a restatement of the
published algorithm (separable Keys cubic a=-0.5, support 4, pixel-centre aligned grid, weights renormalised,
antialiasing only when down-scaling) and is PARITY-UNPINNED.  TEST INFRASTRUCTURE ONLY."""
import math
import torch


def _cubic(x):
    ax = x.abs()
    ax2, ax3 = ax * ax, ax * ax * ax
    return ((1.5 * ax3 - 2.5 * ax2 + 1.) * (ax <= 1.).to(x.dtype) +
            (-0.5 * ax3 + 2.5 * ax2 - 4. * ax + 2.) * ((ax > 1.) & (ax <= 2.)).to(x.dtype))


def _reflect(idx, n):
    # numpy/torch 'reflect' (no edge repeat), valid for |overshoot| < n
    idx = torch.where(idx < 0, -idx, idx)
    idx = torch.where(idx >= n, 2 * (n - 1) - idx, idx)
    return idx


def _resize_axis(x, dim, scale, pad_mode):
    n_in = x.shape[dim]
    n_out = int(math.ceil(n_in * scale))
    support = 4.
    aa = scale < 1.
    stretch = (1. / scale) if aa else 1.
    cur_support = support * stretch
    out_coords = torch.arange(n_out, dtype=x.dtype, device=x.device)
    proj = out_coords / scale + (n_in - 1) / 2 - (n_out - 1) / (2 * scale)
    eps = 1e-7
    left = torch.ceil(proj - cur_support / 2 - eps).long()
    n_taps = int(math.ceil(cur_support - eps))
    taps = left[:, None] + torch.arange(n_taps, device=x.device)[None, :]
    w = _cubic((proj[:, None] - taps.to(x.dtype)) / stretch)
    w = w / w.sum(dim=1, keepdim=True)
    if pad_mode == 'reflect':
        idx = _reflect(taps, n_in)
    elif pad_mode in ('edge', 'replicate'):
        idx = taps.clamp(0, n_in - 1)
    else:  # 'constant'
        valid = ((taps >= 0) & (taps < n_in)).to(x.dtype)
        w = w * valid
        idx = taps.clamp(0, n_in - 1)
    xm = x.movedim(dim, -1)
    gathered = xm[..., idx]                       # (..., n_out, n_taps)
    out = (gathered * w).sum(-1)
    return out.movedim(-1, dim)


def resize(input, scale_factors=None, out_shape=None, interp_method=None, support_sz=None,
           antialiasing=True, by_convs=False, scale_tolerance=None, max_numerator=10, pad_mode='constant'):
    assert scale_factors is not None
    if not isinstance(scale_factors, (tuple, list)):
        scale_factors = (scale_factors, scale_factors)
    out = input
    for dim, s in zip((-2, -1), scale_factors):
        if s != 1:
            out = _resize_axis(out, dim % input.ndim, float(s), pad_mode)
    return out
