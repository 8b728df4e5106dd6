"""Forward SDEs of DiffSep: ``MixSDE`` (separation) and ``PriorMixSDE`` (enhancement).

Same constructor arguments, registry names and attributes as the reference
(``sdes/sdes.py:180-349, 352-590``).  The arithmetic the reference spreads over einsums and ~10
element-wise launches per update lives in three fused kernels (``csrc/sde.cu``): every matrix it
builds is ``a A + b Pn`` (A = channel averaging, Pn = I - A), so ``L v = a vbar + b (v - vbar)`` — for
``ndim`` = 2 sources and for the 3-speaker models (``ndim`` = 3) alike.  ``MixSDE.prior_sampling`` hard-codes 2
sources in the reference (:344), so only ``PriorMixSDE`` can sample with ``ndim`` = 3 there, and here.
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
        if ndim not in (2, 3):
            raise NotImplementedError(f"ndim={ndim}: the DiffSep models separate 2 or 3 sources")
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
        return ops.sde_params(self.d_lambda, self.sigma_min, self.sigma_max, self.T, self.ndim)

    def _sigma_mix(self, y):
        return None

    def _check(self, x, y):
        if x.dim() != 3 or x.shape[1] != self.ndim:
            raise ValueError(f"expected x of shape [B, {self.ndim}, T], got {tuple(x.shape)}")
        if y.dim() != 3 or y.shape[1] != 1 or y.shape[0] != x.shape[0] or y.shape[2] != x.shape[2]:
            raise ValueError(f"expected mix of shape [B, 1, T] matching x, got {tuple(y.shape)}")

    def _prior_mean(self, y):
        """-> (mix channels, factor on the mean) of prior_sampling for an input with y.shape[1] channels."""
        raise NotImplementedError

    def prior_sampling(self, shape, y):
        """x_T = mean(y) + L(T) z   (sdes.py:334-346 / 564-587); y is the mixture [B,1,T] or, through the sampler's
        ``true_mean``, a [B,ndim,T] tensor."""
        if tuple(shape) != tuple(y.shape):
            warnings.warn(f"Target shape {shape} does not match shape of y {y.shape}! Ignoring target shape.")
        if y.dim() != 3:
            raise ValueError(f"expected a [B, C, T] input, got {tuple(y.shape)}")
        ch, mean_scale = self._prior_mean(y)
        sig = self._sigma_mix(y)              # keyed on the caller's tensor, not on a contiguous temporary
        y = y.contiguous().float()
        B, _, T = y.shape
        x = torch.empty(B, self.ndim, T, device=y.device, dtype=torch.float32)
        z, seed, off = _noise.SOURCE.next((B, self.ndim, T), y.device)
        ops.sde_prior(self._params(), y, sig, z, seed, off, B, T, x, mix_channels=ch, mean_scale=mean_scale,
                      sigma_channels=ch if sig is not None else 1)
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
        if self.ndim != 2:
            raise NotImplementedError("the ald corrector is defined for MixSDE with 2 sources (correctors.py:64-67)")
        B, _, T = x.shape
        x_out, x_mean = torch.empty_like(x), torch.empty_like(x)
        z, seed, off = _noise.SOURCE.next(x.shape, x.device)
        ops.sde_corrector_ald(self._params(), x, score, t, z, seed, off, float(snr), B, T, x_out, x_mean)
        return x_out, x_mean

    def reverse(self, score_fn, probability_flow=False):
        return RSDE(self, score_fn, probability_flow)

    # ---- training-side forward pieces (forward only, no autograd)
    def marginal_sample(self, x0, t, y):
        """x_t ~ p_t(x | x0) = N(mean, L L^T) of ``marginal_prob`` (sdes.py:322-324 / 560-562), i.e. the
        ``x_t = mean + mult_std(L, z)`` of ``DiffSepModel.sample_prior`` (pl_model.py:247), in one kernel: -> (x_t, z)."""
        self._check(x0, y)
        x0, t = x0.contiguous().float(), t.contiguous().float()
        B, _, T = x0.shape
        if t.shape != (B,):
            raise ValueError(f"expected one time per batch entry, got {tuple(t.shape)}")
        x_t, z_out = torch.empty_like(x0), torch.empty_like(x0)
        z, seed, off = _noise.SOURCE.next(x0.shape, x0.device)
        ops.sde_perturb(self._params(), x0, t, self._sigma_mix(y), z, seed, off, B, T, x_t, z_out)
        return x_t, z_out

    def score_loss(self, score, z, t, y):
        """per-sample ``mean((mult_std(L, score) + z)^2)`` over (channel, time): the MSE of ``compute_score_loss``
        (pl_model.py:418-424) with ``reduction="none"`` semantics; float32 [B]."""
        self._check(score, y)
        B, _, T = score.shape
        loss = torch.empty(B, dtype=torch.float64, device=score.device)
        ops.score_loss(self._params(), score.contiguous().float(), z.contiguous().float(), t.contiguous().float(),
                       self._sigma_mix(y), B, T, loss)
        return loss.float()


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

    def _prior_mean(self, y):
        # mean = broadcast_to(0.5 * y, (B, 2, T)) (sdes.py:344): two sources hard-coded, and the 0.5 stays even
        # when a [B,2,T] true_mean comes in; with ndim = 3 the reference fails here (shape error in std @ noise)
        if self.ndim != 2:
            raise RuntimeError("MixSDE.prior_sampling supports 2 sources only (the reference hard-codes the shape "
                               "(B, 2, T), sdes.py:344); 3-speaker models use PriorMixSDE")
        if y.shape[1] not in (1, 2):
            raise RuntimeError(f"cannot broadcast a {y.shape[1]}-channel input to (B, 2, T)")
        return y.shape[1], 0.5


@SDERegistry.register("priormix")
class PriorMixSDE(SDE):
    prior = True

    def __init__(self, ndim, d_lambda, sigma_min, sigma_max, N=1000, avg_len=510):
        super().__init__(ndim, d_lambda, sigma_min, sigma_max, N)
        self.avg_len = avg_len

    def copy(self):
        return PriorMixSDE(self.ndim, self.d_lambda, self.sigma_min, self.sigma_max, N=self.N,
                           avg_len=self.avg_len)

    def _prior_mean(self, y):
        # sdes.py:571-583
        if y.shape[1] == self.ndim:
            return self.ndim, 1.0
        if y.shape[1] == 1:
            return 1, 0.5
        raise ValueError("The input provided to prior_sampling should have 1 channel, or the same as the number of "
                         f"speakers. Found {y.shape[1]} channels instead.")

    def _std_sigma_mix(self, mix):
        """0.5 sqrt(clamp(avgpool_k(mix^2), 1e-4)) per channel  (sdes.py:477-489) -> [B, T] ([B, C, T] for C > 1)."""
        mix = mix.contiguous().float()
        B, Cc, T = mix.shape
        out = torch.empty(B * Cc, T, device=mix.device, dtype=torch.float32)
        ops.sigma_mix(mix, B * Cc, T, self.avg_len, out)
        return out if Cc == 1 else out.view(B, Cc, T)

    def _sigma_mix(self, y):
        # constant over a sampling run: computed once per mixture tensor.  The cache holds the tensor itself (so
        # its storage cannot be handed to another tensor while cached) and its version counter (in-place edits)
        c = self._sig_cache
        if c is None or c[0] is not y or c[1] != y._version:
            self._sig_cache = (y, y._version, self._std_sigma_mix(y))
        return self._sig_cache[2]

    def reset_cache(self):
        self._sig_cache = None
