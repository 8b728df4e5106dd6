"""Seeded synthetic parameters for the DiffSep score model, keyed exactly like the reference's
``state_dict`` (``backbone.all_modules.{i}.…``): lets ``bench.py`` and ``separate.py --model
synthetic`` run the real architecture when no trained checkpoint is reachable (no network).

Why not fresh-init weights: the reference initialises every ``Conv_1`` / ``NIN_3`` / pyramid conv
with ``init_scale=0 -> 1e-10`` (``models/ncsnpp_utils/layers.py:99-102``, ``layerspp.py:73,282``,
``ncsnpp.py:259,283``), so a fresh model outputs a constant.  Values here are a pure function of
``(seed, parameter name, shape)``; the test oracle has its own copy of this generator and
``tests/test_host_cpu.py`` checks the two agree bit for bit.

The module enumeration restates ``NCSNpp.__init__`` (reference ``models/ncsnpp.py:104-308``) for
the only configuration the hot path uses.
"""
from __future__ import annotations

import hashlib
import math
from collections import OrderedDict

import torch

CH_MULT = (1, 1, 2, 2, 2, 2, 2)      # reference models/ncsnpp.py:50
NUM_RES_BLOCKS = 2                   # :51
ATTN_RESOLUTIONS = (16,)             # :52
IMAGE_SIZE = 256                     # :64


def _resblock(shapes, prefix, cin, cout, temb_dim, up=False, down=False):
    # reference layerspp.py:246-289 (parameter registration order)
    shapes[f"{prefix}.GroupNorm_0.weight"] = (cin,)
    shapes[f"{prefix}.GroupNorm_0.bias"] = (cin,)
    shapes[f"{prefix}.Conv_0.weight"] = (cout, cin, 3, 3)
    shapes[f"{prefix}.Conv_0.bias"] = (cout,)
    shapes[f"{prefix}.Dense_0.weight"] = (cout, temb_dim)
    shapes[f"{prefix}.Dense_0.bias"] = (cout,)
    shapes[f"{prefix}.GroupNorm_1.weight"] = (cout,)
    shapes[f"{prefix}.GroupNorm_1.bias"] = (cout,)
    shapes[f"{prefix}.Conv_1.weight"] = (cout, cout, 3, 3)
    shapes[f"{prefix}.Conv_1.bias"] = (cout,)
    if cin != cout or up or down:
        shapes[f"{prefix}.Conv_2.weight"] = (cout, cin, 1, 1)
        shapes[f"{prefix}.Conv_2.bias"] = (cout,)


def _attn(shapes, prefix, c):
    # reference layerspp.py:65-74
    shapes[f"{prefix}.GroupNorm_0.weight"] = (c,)
    shapes[f"{prefix}.GroupNorm_0.bias"] = (c,)
    for i in range(4):
        shapes[f"{prefix}.NIN_{i}.W"] = (c, c)
        shapes[f"{prefix}.NIN_{i}.b"] = (c,)


def backbone_param_shapes(nf=128, ch_in=6, ch_out=4):
    """Ordered ``name -> shape`` for the NCSN++ backbone, in ``parameters()`` order.

    Names are relative to the backbone (``output_layer.*`` first, then ``all_modules.i.*``),
    as registered in reference ``models/ncsnpp.py:105,308``.
    """
    shapes = OrderedDict()
    temb_dim = 4 * nf
    shapes["output_layer.weight"] = (ch_out, ch_in, 1, 1)
    shapes["output_layer.bias"] = (ch_out,)
    m = 0

    def name(i):
        return f"all_modules.{i}"

    shapes[f"{name(m)}.W"] = (nf,); m += 1                       # GaussianFourierProjection
    shapes[f"{name(m)}.weight"] = (temb_dim, 2 * nf)
    shapes[f"{name(m)}.bias"] = (temb_dim,); m += 1
    shapes[f"{name(m)}.weight"] = (temb_dim, temb_dim)
    shapes[f"{name(m)}.bias"] = (temb_dim,); m += 1
    shapes[f"{name(m)}.weight"] = (nf, ch_in, 3, 3)               # input conv3x3
    shapes[f"{name(m)}.bias"] = (nf,); m += 1

    hs_c = [nf]
    in_ch = nf
    nres = len(CH_MULT)
    for lvl in range(nres):
        res = IMAGE_SIZE // (2 ** lvl)
        for _ in range(NUM_RES_BLOCKS):
            out_ch = nf * CH_MULT[lvl]
            _resblock(shapes, name(m), in_ch, out_ch, temb_dim); m += 1
            in_ch = out_ch
            if res in ATTN_RESOLUTIONS:
                _attn(shapes, name(m), in_ch); m += 1
            hs_c.append(in_ch)
        if lvl != nres - 1:
            _resblock(shapes, name(m), in_ch, in_ch, temb_dim, down=True); m += 1
            shapes[f"{name(m)}.Conv_0.weight"] = (in_ch, ch_in, 1, 1)   # Combine
            shapes[f"{name(m)}.Conv_0.bias"] = (in_ch,); m += 1
            hs_c.append(in_ch)

    in_ch = hs_c[-1]
    _resblock(shapes, name(m), in_ch, in_ch, temb_dim); m += 1
    _attn(shapes, name(m), in_ch); m += 1
    _resblock(shapes, name(m), in_ch, in_ch, temb_dim); m += 1

    for lvl in reversed(range(nres)):
        res = IMAGE_SIZE // (2 ** lvl)
        for _ in range(NUM_RES_BLOCKS + 1):
            out_ch = nf * CH_MULT[lvl]
            _resblock(shapes, name(m), in_ch + hs_c.pop(), out_ch, temb_dim); m += 1
            in_ch = out_ch
        if res in ATTN_RESOLUTIONS:
            _attn(shapes, name(m), in_ch); m += 1
        shapes[f"{name(m)}.weight"] = (in_ch,)                   # pyramid GroupNorm
        shapes[f"{name(m)}.bias"] = (in_ch,); m += 1
        shapes[f"{name(m)}.weight"] = (ch_in, in_ch, 3, 3)       # pyramid conv3x3 -> ch_in
        shapes[f"{name(m)}.bias"] = (ch_in,); m += 1
        if lvl != 0:
            _resblock(shapes, name(m), in_ch, in_ch, temb_dim, up=True); m += 1
    assert not hs_c
    return shapes


def _gen(seed, name):
    h = hashlib.sha256(f"{seed}:{name}".encode()).digest()
    g = torch.Generator(device="cpu")
    g.manual_seed(int.from_bytes(h[:7], "little"))
    return g


def make_backbone_params(nf=128, ch_in=6, ch_out=4, seed=0, dtype=torch.float32):
    """Seeded non-degenerate parameters (fp32 values; optionally widened to fp64)."""
    shapes = backbone_param_shapes(nf, ch_in, ch_out)
    params = OrderedDict()
    for name, shape in shapes.items():
        g = _gen(seed, name)
        leaf = name.rsplit(".", 1)[-1]
        if leaf == "W" and len(shape) == 1:              # Fourier frequencies ~ N(0, 16^2)
            v = torch.randn(shape, generator=g) * 16.0
        elif "GroupNorm" in name or (leaf in ("weight", "bias") and len(shape) == 1
                                     and _is_norm(name, shapes)):
            if leaf == "weight":
                v = 1.0 + 0.1 * torch.randn(shape, generator=g)
            else:
                v = 0.1 * torch.randn(shape, generator=g)
        elif leaf in ("bias", "b"):
            v = 0.05 * torch.randn(shape, generator=g)
        else:
            # variance-scaling fan_avg uniform (reference layers.py:63-102), scale 1 for
            # conv/linear, 0.1 for NIN (layers.py:679)
            if leaf == "W":                               # NIN: (in, out)
                fan_in, fan_out, scale = shape[0], shape[1], 0.1
            else:
                rf = 1
                for s in shape[2:]:
                    rf *= s
                fan_in, fan_out, scale = shape[1] * rf, shape[0] * rf, 1.0
            var = scale / ((fan_in + fan_out) / 2.0)
            v = (torch.rand(shape, generator=g) * 2.0 - 1.0) * math.sqrt(3.0 * var)
        params[name] = v.to(torch.float32).to(dtype)
    return params


def _is_norm(name, shapes):
    # the pyramid GroupNorm is registered as bare ``all_modules.i.{weight,bias}`` of rank 1
    # with no sibling of rank > 1 ... whereas a conv/linear bias has a rank>1 ``weight``.
    base = name.rsplit(".", 1)[0]
    w = shapes.get(base + ".weight")
    return w is not None and len(w) == 1


def make_score_model_state_dict(nf=128, num_sources=2, seed=0, n_fft=510):
    """Full ``score_model`` state-dict layout of a checkpoint:
    ``backbone.*`` + ``stft.window`` + ``stft_inv.window``."""
    sd = OrderedDict()
    for k, v in make_backbone_params(nf, 2 * num_sources + 2, 2 * num_sources, seed).items():
        sd["backbone." + k] = v
    sd["stft.window"] = torch.hann_window(n_fft)
    sd["stft_inv.window"] = torch.hann_window(n_fft)
    return sd
