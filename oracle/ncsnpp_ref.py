"""ORACLE (test infrastructure, not product code): CPU restatement of the NCSN++ score
backbone used by DiffSep, written functionally over a flat parameter dict.

Each function cites the reference lines it restates.  Works in fp32 (the parity
target) or fp64 (to measure the *true* error of both the reference and the CUDA path).
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu/reference legs may
import this module; the product path never does.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from .weights import ATTN_RESOLUTIONS, CH_MULT, NUM_RES_BLOCKS

FIR_TAPS = (1.0, 3.0, 3.0, 1.0)  # reference models/ncsnpp.py:56


def silu(x):
    return x * torch.sigmoid(x)  # nn.SiLU, reference layers.py:38-39


def group_norm(x, w, b):
    # nn.GroupNorm(num_groups=min(C//4, 32), eps=1e-6), reference layerspp.py:264-266
    c = x.shape[1]
    return F.group_norm(x, min(c // 4, 32), w, b, eps=1e-6)


def fir_down2(x):
    """downsample_2d(x, [1,3,3,1], factor=2): reference up_or_down_sampling.py:242-273.

    Separable closed form: per axis, zeros outside,
    ``y[i] = (x[2i-1] + 3 x[2i] + 3 x[2i+1] + x[2i+2]) / 8``.
    """
    k = torch.tensor(FIR_TAPS, dtype=x.dtype) / 8.0
    c = x.shape[1]
    xp = F.pad(x, (1, 1, 1, 1))
    kw = k.view(1, 1, 1, 4).expand(c, 1, 1, 4)
    kh = k.view(1, 1, 4, 1).expand(c, 1, 4, 1)
    y = F.conv2d(xp, kw, stride=(1, 2), groups=c)
    y = F.conv2d(y, kh, stride=(2, 1), groups=c)
    return y


def fir_up2(x):
    """upsample_2d(x, [1,3,3,1], factor=2): reference up_or_down_sampling.py:206-239.

    Separable closed form: per axis (gain 2 per axis, 4 total), zeros outside,
    ``y[2i] = (x[i-1] + 3 x[i]) / 4``, ``y[2i+1] = (3 x[i] + x[i+1]) / 4``.
    """
    def up_axis(v, dim):
        n = v.shape[dim]
        pad = [0, 0, 0, 0]
        idx = 0 if dim == 3 else 2
        pad[idx] = 1
        pad[idx + 1] = 1
        vp = F.pad(v, pad)
        prev = vp.narrow(dim, 0, n)
        cur = vp.narrow(dim, 1, n)
        nxt = vp.narrow(dim, 2, n)
        even = (prev + 3.0 * cur) / 4.0
        odd = (3.0 * cur + nxt) / 4.0
        out = torch.stack((even, odd), dim=dim + 1)
        shape = list(v.shape)
        shape[dim] = 2 * n
        return out.reshape(shape)

    return up_axis(up_axis(x, 3), 2)


def nin(x, W, b):
    # NIN.forward: per-pixel x @ W + b with W stored (in, out); reference layers.py:678-689
    y = torch.einsum("bchw,cd->bdhw", x, W)
    return y + b.view(1, -1, 1, 1)


class _P:
    """Tiny helper: parameters of module ``all_modules.{i}``."""

    def __init__(self, params, i):
        self.p = params
        self.prefix = f"all_modules.{i}."

    def __getitem__(self, k):
        return self.p[self.prefix + k]

    def has(self, k):
        return (self.prefix + k) in self.p


def resblock(p, x, temb_act, up=False, down=False):
    """ResnetBlockBigGANpp.forward, reference layerspp.py:291-323 (dropout p=0 in eval)."""
    h = silu(group_norm(x, p["GroupNorm_0.weight"], p["GroupNorm_0.bias"]))
    if up:
        h, x = fir_up2(h), fir_up2(x)
    elif down:
        h, x = fir_down2(h), fir_down2(x)
    h = F.conv2d(h, p["Conv_0.weight"], p["Conv_0.bias"], padding=1)
    h = h + F.linear(temb_act, p["Dense_0.weight"], p["Dense_0.bias"])[:, :, None, None]
    h = silu(group_norm(h, p["GroupNorm_1.weight"], p["GroupNorm_1.bias"]))
    h = F.conv2d(h, p["Conv_1.weight"], p["Conv_1.bias"], padding=1)
    if p.has("Conv_2.weight"):
        x = F.conv2d(x, p["Conv_2.weight"], p["Conv_2.bias"])
    return (x + h) / math.sqrt(2.0)


def attnblock(p, x):
    """AttnBlockpp.forward, reference layerspp.py:76-92 (skip_rescale=True)."""
    B, C, H, W = x.shape
    h = group_norm(x, p["GroupNorm_0.weight"], p["GroupNorm_0.bias"])
    q = nin(h, p["NIN_0.W"], p["NIN_0.b"])
    k = nin(h, p["NIN_1.W"], p["NIN_1.b"])
    v = nin(h, p["NIN_2.W"], p["NIN_2.b"])
    w = torch.einsum("bchw,bcij->bhwij", q, k) * (int(C) ** (-0.5))
    w = F.softmax(w.reshape(B, H, W, H * W), dim=-1).reshape(B, H, W, H, W)
    h = torch.einsum("bhwij,bcij->bchw", w, v)
    h = nin(h, p["NIN_3.W"], p["NIN_3.b"])
    return (x + h) / math.sqrt(2.0)


def time_embedding(params, t):
    """Gaussian-Fourier embedding + 2 Linear; reference ncsnpp.py:324-343, layerspp.py:39-41.

    The multiplication order ``log(t)[:,None] * W[None,:] * 2 * np.pi`` is kept (fp32
    rounding of a ~1e3 rad phase depends on it).
    """
    W = params["all_modules.0.W"]
    x_proj = torch.log(t)[:, None] * W[None, :] * 2 * math.pi
    emb = torch.cat([torch.sin(x_proj), torch.cos(x_proj)], dim=-1)
    temb = F.linear(emb, params["all_modules.1.weight"], params["all_modules.1.bias"])
    temb = F.linear(silu(temb), params["all_modules.2.weight"], params["all_modules.2.bias"])
    return temb


def ncsnpp_forward(params, x, t, taps=None):
    """NCSNpp.forward, reference models/ncsnpp.py:319-478.

    ``x``: [B, 6, 256, W] (W multiple of 64), ``t``: [B].  ``taps`` (optional dict) receives
    named intermediate activations for layer-by-layer parity tests.
    """
    def tap(name, v):
        if taps is not None:
            taps[name] = v

    nres = len(CH_MULT)
    m = 0
    temb = time_embedding(params, t); m += 3
    temb_act = silu(temb)  # every block applies act(temb) before Dense_0, layerspp.py:313
    tap("temb", temb)

    x = 2 * x - 1.0  # centered=False, ncsnpp.py:347-349
    input_pyramid = x
    h = F.conv2d(x, params[f"all_modules.{m}.weight"], params[f"all_modules.{m}.bias"], padding=1)
    m += 1
    hs = [h]
    tap("conv_in", h)

    for lvl in range(nres):
        for _ in range(NUM_RES_BLOCKS):
            h = resblock(_P(params, m), hs[-1], temb_act); m += 1
            if h.shape[-2] in ATTN_RESOLUTIONS:  # executed by H, ncsnpp.py:367-371
                h = attnblock(_P(params, m), h); m += 1
            hs.append(h)
        if lvl != nres - 1:
            h = resblock(_P(params, m), hs[-1], temb_act, down=True); m += 1
            input_pyramid = fir_down2(input_pyramid)          # ncsnpp.py:384
            p = _P(params, m); m += 1                            # Combine(sum), layerspp.py:52-57
            h = F.conv2d(input_pyramid, p["Conv_0.weight"], p["Conv_0.bias"]) + h
            hs.append(h)
        tap(f"down{lvl}", h)

    h = hs[-1]
    h = resblock(_P(params, m), h, temb_act); m += 1
    h = attnblock(_P(params, m), h); m += 1
    h = resblock(_P(params, m), h, temb_act); m += 1
    tap("mid", h)

    pyramid = None
    for lvl in reversed(range(nres)):
        for _ in range(NUM_RES_BLOCKS + 1):
            h = resblock(_P(params, m), torch.cat([h, hs.pop()], dim=1), temb_act); m += 1
        if h.shape[-2] in ATTN_RESOLUTIONS:
            h = attnblock(_P(params, m), h); m += 1
        # progressive == "output_skip", ncsnpp.py:419-440
        gp = _P(params, m); m += 1
        cp = _P(params, m); m += 1
        ph = silu(group_norm(h, gp["weight"], gp["bias"]))
        ph = F.conv2d(ph, cp["weight"], cp["bias"], padding=1)
        pyramid = ph if pyramid is None else fir_up2(pyramid) + ph
        if lvl != 0:
            h = resblock(_P(params, m), h, temb_act, up=True); m += 1
        tap(f"up{lvl}", h)
        tap(f"pyr{lvl}", pyramid)
    assert not hs

    h = pyramid / t.reshape(-1, 1, 1, 1)   # scale_by_sigma, ncsnpp.py:472-474
    return F.conv2d(h, params["output_layer.weight"], params["output_layer.bias"])


def count_flops(nf=128, W=256, ch_in=6, ch_out=4):
    """Algorithmic FLOPs (2*MAC of conv/NIN/attention/linear) per sample per evaluation.

    Restates the hook-based count of SURVEY.md §8d (532.891 GFLOP at nf=128, 256x256).
    """
    from .weights import backbone_param_shapes
    total = {"conv3x3": 0, "conv1x1": 0, "nin": 0, "attn": 0, "linear": 0}
    temb = 4 * nf
    total["linear"] += 2 * (2 * nf * temb + temb * temb)
    nres = len(CH_MULT)

    def rb(cin, cout, H, Wd, short):
        total["conv3x3"] += 2 * 9 * (cin * cout + cout * cout) * H * Wd
        total["linear"] += 2 * temb * cout
        if short:
            total["conv1x1"] += 2 * cin * cout * H * Wd

    def at(c, H, Wd):
        S = H * Wd
        total["nin"] += 4 * 2 * c * c * S
        total["attn"] += 2 * 2 * S * S * c

    H, Wd = 256, W
    total["conv3x3"] += 2 * 9 * ch_in * nf * H * Wd
    hs_c = [nf]
    in_ch = nf
    for lvl in range(nres):
        for _ in range(NUM_RES_BLOCKS):
            out_ch = nf * CH_MULT[lvl]
            rb(in_ch, out_ch, H, Wd, in_ch != out_ch)
            in_ch = out_ch
            if H in ATTN_RESOLUTIONS:
                at(in_ch, H, Wd)
            hs_c.append(in_ch)
        if lvl != nres - 1:
            H, Wd = H // 2, Wd // 2
            rb(in_ch, in_ch, H, Wd, True)
            total["conv1x1"] += 2 * ch_in * in_ch * H * Wd
            hs_c.append(in_ch)
    rb(in_ch, in_ch, H, Wd, False)
    at(in_ch, H, Wd)
    rb(in_ch, in_ch, H, Wd, False)
    for lvl in reversed(range(nres)):
        for _ in range(NUM_RES_BLOCKS + 1):
            out_ch = nf * CH_MULT[lvl]
            rb(in_ch + hs_c.pop(), out_ch, H, Wd, True)
            in_ch = out_ch
        if H in ATTN_RESOLUTIONS:
            at(in_ch, H, Wd)
        total["conv3x3"] += 2 * 9 * in_ch * ch_in * H * Wd
        if lvl != 0:
            H, Wd = H * 2, Wd * 2
            rb(in_ch, in_ch, H, Wd, True)
    total["conv1x1"] += 2 * ch_in * ch_out * H * Wd
    total["total"] = sum(total.values())
    return total
