"""Seeded inputs shared by ``make_golden.py`` (which runs the real reference, only in the
build container) and by the tests (which re-create the same inputs anywhere)."""
from __future__ import annotations

import torch


def gen(seed):
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return g


def synthetic_mix(u, T, scale=0.1):
    """SURVEY.md §8d synthetic utterance ``u``: two white sources, summed. -> [1, T]"""
    g = gen(1234 + u)
    s1 = scale * torch.randn(T, generator=g)
    s2 = scale * torch.randn(T, generator=g)
    return (s1 + s2)[None]


def batch_mix(B, T, first=0):
    return torch.stack([synthetic_mix(first + u, T) for u in range(B)])  # [B,1,T]


def score_inputs(B, T, seed=7):
    """(xt [B,2,T], t [B], mix [B,1,T]) for a score-model evaluation."""
    g = gen(seed)
    mix = torch.randn(B, 1, T, generator=g)
    xt = 0.5 * mix + 0.3 * torch.randn(B, 2, T, generator=g)
    t = 0.03 + 0.97 * torch.rand(B, generator=g)
    return xt, t, mix


def sampler_noises(B, T, N, cs, seed=999):
    g = gen(seed)
    n = 1 + N * (cs + 1)
    return [torch.randn(B, 2, T, generator=g) for _ in range(n)]


def analytic_score(x, t, mix):
    """Cheap closed-form stand-in for the network."""
    return -0.7 * x + 0.1 * mix


# op-level cases: (name, channels, H, W)
RESBLOCK_CASES = [
    ("plain", dict(cin=16, cout=16, up=False, down=False), (2, 16, 8, 12)),
    ("widen", dict(cin=16, cout=32, up=False, down=False), (2, 16, 8, 12)),
    ("cat", dict(cin=48, cout=32, up=False, down=False), (1, 48, 8, 8)),
    ("down", dict(cin=16, cout=16, up=False, down=True), (2, 16, 8, 12)),
    ("up", dict(cin=16, cout=16, up=True, down=False), (2, 16, 4, 6)),
]


# 3-source / true_mean sampler cases of make_golden_ndim.py: (name, sde, ndim, true_mean channels, corrector steps)
NDIM_CASES = [("prior3", "priormix", 3, 0, 1), ("prior3_cs2", "priormix", 3, 0, 2),
              ("prior3_true_mean", "priormix", 3, 3, 1), ("prior2_true_mean", "priormix", 2, 2, 1),
              ("mix2_true_mean", "mix", 2, 2, 1)]
NDIM_N, NDIM_B, NDIM_T = 6, 2, 1024


def ndim_noises(ndim, cs):
    g = gen(4242 + ndim)
    return [torch.randn(NDIM_B, ndim, NDIM_T, generator=g) for _ in range(1 + NDIM_N * (cs + 1))]


def ndim_true_mean(ch):
    g = gen(77 + ch)
    return 0.4 * torch.randn(NDIM_B, ch, NDIM_T, generator=g)
