"""Corrector plugins (reference ``sdes/correctors.py:11-141``).  ``ald2`` — the one the CLIs
hard-code (``separate.py:87-92``) — is implemented as one fused kernel per step; ``none`` keeps the
reference's behaviour.  ``langevin`` / ``ald`` are outside the hot path."""
from __future__ import annotations

import abc

from ..utils.registry import Registry
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


@CorrectorRegistry.register("none")
class NoneCorrector(Corrector):
    """An empty corrector that does nothing."""

    def __init__(self, *args, **kwargs):
        self.snr = 0
        self.n_steps = 0

    def update_fn(self, x, t, *args, **kwargs):
        return x, x
