"""Forward SDEs of DiffSep: ``MixSDE`` (separation) and ``PriorMixSDE`` (enhancement).

Same constructor arguments, registry names and attributes as the reference
(``sdes/sdes.py:180-349, 352-590``).  The arithmetic the reference spreads over einsums and ~10
element-wise launches per update lives in three fused kernels (``csrc/sde.cu``): every matrix it
builds is ``a A + b Pn`` (A = channel averaging, Pn = I - A), so ``L v = a vbar + b (v - vbar)``.
The reference hard-codes 2 sources in ``prior_sampling`` (:344); so does this path.
"""
from __future__ import annotations

import math
import warnings

import torch

from .. import ops
from ..utils.registry import Registry
from . import noise as _noise

SDERegistry = Registry("SDE")


class SDE:
    """Common surface used by predictors / correctors / the sampler."""

    prior = False

    def __init__(self, ndim, d_lambda, sigma_min, sigma_max, N=1000):
        if ndim != 2:
            raise NotImplementedError("the DiffSep hot path separates 2 sources (reference sdes.py:344)")
        self.ndim, self.d_lambda = ndim, d_lambda
        self.sigma_min, self.sigma_max = sigma_min, sigma_max
        self.ratiosig = sigma_max / sigma_min
        self.logsig = math.log(self.ratiosig)
        self.N = N
        self._sig_cache = None

    @property
    def T(self):
        return 1.0

    def _params(self):
        return ops.sde_params(self.d_lambda, self.sigma_min, self.sigma_max, self.T)

    def _sigma_mix(self, y):
        return None

    def _check(self, x, y):
        if x.dim() != 3 or x.shape[1] != 2:
            raise ValueError(f"expected x of shape [B, 2, T], got {tuple(x.shape)}")
        if y.dim() != 3 or y.shape[1] != 1 or y.shape[0] != x.shape[0] or y.shape[2] != x.shape[2]:
            raise ValueError(f"expected mix of shape [B, 1, T] matching x, got {tuple(y.shape)}")

    def prior_sampling(self, shape, y):
        """x_T = 0.5 y (both channels) + L(T) z   (sdes.py:334-346 / 564-587)."""
        if tuple(shape) != tuple(y.shape):
            warnings.warn(f"Target shape {shape} does not match shape of y {y.shape}! Ignoring target shape.")
        y = y.contiguous().float()
        B, _, T = y.shape
        x = torch.empty(B, 2, T, device=y.device, dtype=torch.float32)
        z, seed, off = _noise.SOURCE.next((B, 2, T), y.device)
        ops.sde_prior(self._params(), y, self._sigma_mix(y), z, seed, off, B, T, x)
        return x

    def corrector_update(self, x, score, t, y, snr):
        """ald2 step (correctors.py:116-126): -> (x', x_mean)."""
        self._check(x, y)
        B, _, T = x.shape
        x_out, x_mean = torch.empty_like(x), torch.empty_like(x)
        z, seed, off = _noise.SOURCE.next(x.shape, x.device)
        ops.sde_corrector(self._params(), x, score, t, self._sigma_mix(y), z, seed, off, float(snr), B, T,
                          x_out, x_mean)
        return x_out, x_mean

    def predictor_update(self, x, score, t, y, dt, probability_flow=False):
        """reverse-diffusion / Euler-Maruyama step (predictors.py:39-66, sdes.py:93-107,163-171); with
        ``probability_flow`` the score term is halved and the noise dropped (sdes.py:143-152,167-170) —
        a noise tensor is still drawn, as the reference's ``randn_like`` is."""
        self._check(x, y)
        B, _, T = x.shape
        x_out, x_mean = torch.empty_like(x), torch.empty_like(x)
        z, seed, off = _noise.SOURCE.next(x.shape, x.device)
        ops.sde_predictor(self._params(), x, score, t, self._sigma_mix(y), z, seed, off, float(dt), B, T,
                          x_out, x_mean, probability_flow=probability_flow)
        return x_out, x_mean

    def ald_update(self, x, score, t, snr):
        """original annealed Langevin step (correctors.py:58-91), MixSDE only."""
        B, _, T = x.shape
        x_out, x_mean = torch.empty_like(x), torch.empty_like(x)
        z, seed, off = _noise.SOURCE.next(x.shape, x.device)
        ops.sde_corrector_ald(self._params(), x, score, t, z, seed, off, float(snr), B, T, x_out, x_mean)
        return x_out, x_mean

    def reverse(self, score_fn, probability_flow=False):
        return RSDE(self, score_fn, probability_flow)


class RSDE:
    """Reverse-time SDE handle (reference sdes.py:109-173), reduced to what predictors use."""

    def __init__(self, sde, score_fn, probability_flow=False):
        self.sde, self.score_fn, self.probability_flow = sde, score_fn, probability_flow
        self.N = sde.N

    @property
    def T(self):
        return self.sde.T

    def step(self, x, t, *args, dt=None):
        score = self.score_fn(x, t, *args)
        return self.sde.predictor_update(x, score, t, args[0], 1.0 / self.sde.N if dt is None else dt,
                                         probability_flow=self.probability_flow)


@SDERegistry.register("mix")
class MixSDE(SDE):
    def __init__(self, ndim, d_lambda, sigma_min, sigma_max, N=1000):
        super().__init__(ndim, d_lambda, sigma_min, sigma_max, N)

    def copy(self):
        return MixSDE(self.ndim, self.d_lambda, self.sigma_min, self.sigma_max, N=self.N)


@SDERegistry.register("priormix")
class PriorMixSDE(SDE):
    prior = True

    def __init__(self, ndim, d_lambda, sigma_min, sigma_max, N=1000, avg_len=510):
        super().__init__(ndim, d_lambda, sigma_min, sigma_max, N)
        self.avg_len = avg_len

    def copy(self):
        return PriorMixSDE(self.ndim, self.d_lambda, self.sigma_min, self.sigma_max, N=self.N,
                           avg_len=self.avg_len)

    def _std_sigma_mix(self, mix):
        """0.5 sqrt(clamp(avgpool_k(mix^2), 1e-4))  (sdes.py:477-489) -> [B, T]."""
        mix = mix.contiguous().float()
        B, _, T = mix.shape
        out = torch.empty(B, T, device=mix.device, dtype=torch.float32)
        return ops.sigma_mix(mix, B, T, self.avg_len, out)

    def _sigma_mix(self, y):
        # constant over a sampling run: computed once per mixture tensor
        key = (y.data_ptr(), tuple(y.shape))
        if self._sig_cache is None or self._sig_cache[0] != key:
            self._sig_cache = (key, self._std_sigma_mix(y))
        return self._sig_cache[1]

    def reset_cache(self):
        self._sig_cache = None
