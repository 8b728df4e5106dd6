"""Corrector plugins (reference ``sdes/correctors.py:11-141``): ``ald2`` — the one the CLIs hard-code
(``separate.py:87-92``) —, ``ald`` and ``langevin`` are one fused kernel per step (``langevin`` plus a
norm reduction); ``none`` keeps the reference's behaviour."""
from __future__ import annotations

import abc

import torch

from .. import ops
from ..utils.registry import Registry
from . import noise as _noise
from . import sdes

CorrectorRegistry = Registry("Corrector")


class Corrector(abc.ABC):
    """The abstract class for a corrector algorithm."""

    def __init__(self, sde, score_fn, snr, n_steps):
        super().__init__()
        self.rsde = sde.reverse(score_fn)
        self.score_fn = score_fn
        self.snr = snr
        self.n_steps = n_steps

    @abc.abstractmethod
    def update_fn(self, x, t, *args, **kwargs):
        """One update of the corrector: returns (next state, next state without noise)."""


@CorrectorRegistry.register("ald2")
class AnnealedLangevinDynamics2(Corrector):
    """x_mean = x + 2 snr^2 L L score ; x = x_mean + 2 snr L z, L = marginal std at t
    (correctors.py:109-128).  With n_steps = 0 it returns (x, x), as the reference does."""

    def __init__(self, sde, score_fn, snr, n_steps):
        super().__init__(sde, score_fn, snr, n_steps)
        self.sde = sde
        if not isinstance(sde, (sdes.MixSDE, sdes.PriorMixSDE)):
            raise NotImplementedError(f"SDE class {sde.__class__.__name__} not yet supported.")

    def update_fn(self, x, t, *args, **kwargs):
        x_mean = x
        for _ in range(self.n_steps):
            grad = self.score_fn(x, t, *args)
            x, x_mean = self.sde.corrector_update(x, grad, t, args[0], self.snr)
        return x, x_mean


@CorrectorRegistry.register("ald")
class AnnealedLangevinDynamics(Corrector):
    """The original annealed Langevin dynamics corrector of NCSN (correctors.py:58-91): step size
    2 (snr std)^2 with std the scalar marginal standard deviation; MixSDE only, like the reference."""

    def __init__(self, sde, score_fn, snr, n_steps):
        super().__init__(sde, score_fn, snr, n_steps)
        if not isinstance(sde, sdes.MixSDE):
            raise NotImplementedError(f"SDE class {sde.__class__.__name__} not yet supported.")
        self.sde = sde

    def update_fn(self, x, t, *args, **kwargs):
        x_mean = x
        for _ in range(self.n_steps):
            grad = self.score_fn(x, t, *args)
            x, x_mean = self.sde.ald_update(x, grad, t, self.snr)
        return x, x_mean


@CorrectorRegistry.register("langevin")
class LangevinCorrector(Corrector):
    """Langevin corrector of score_sde (correctors.py:35-55): the step size comes from the batch means
    of the per-entry score and noise norms, so batch entries are coupled (do not shard a batch across
    GPUs with it if results must match the single-GPU run)."""

    def update_fn(self, x, t, *args, **kwargs):
        x_mean = x
        B = x.shape[0]
        n = x[0].numel()
        for _ in range(self.n_steps):
            grad = self.score_fn(x, t, *args)
            z, seed, off = _noise.SOURCE.next(x.shape, x.device)
            if z is None:
                z = ops.randn(torch.empty_like(x), seed, off)
            norms = torch.empty(2, B, device=x.device, dtype=torch.float32)
            x_out, x_mean = torch.empty_like(x), torch.empty_like(x)
            ops.sde_corrector_langevin(x.contiguous(), grad.contiguous(), z, float(self.snr), B, n, norms, x_out,
                                       x_mean)
            x = x_out
        return x, x_mean


@CorrectorRegistry.register("none")
class NoneCorrector(Corrector):
    """An empty corrector that does nothing.  Deliberate deviation: the reference's returns the 1-tuple ``(x,)``
    (correctors.py:140-141), which breaks the sampler's ``xt, xt_mean = corrector.update_fn(...)`` unpacking
    (sdes/__init__.py:179) — its "none" corrector cannot run in its own PC sampler.  Here it returns ``(x, x)``
    like ``ald2`` with ``n_steps=0`` (what configs[0]'s predictor-only run uses), so the registered name works."""

    def __init__(self, *args, **kwargs):
        self.snr = 0
        self.n_steps = 0

    def update_fn(self, x, t, *args, **kwargs):
        return x, x
