"""Predictor plugins (reference ``sdes/predictors.py:10-77``): same registry names, constructor
and ``update_fn(x, t, *args) -> (x, x_mean)`` contract; the update itself is one fused kernel."""
from __future__ import annotations

import abc

from ..utils.registry import Registry

PredictorRegistry = Registry("Predictor")


class Predictor(abc.ABC):
    """The abstract class for a predictor algorithm."""

    def __init__(self, sde, score_fn, probability_flow=False):
        super().__init__()
        self.sde = sde
        # like the reference (predictors.py:13-18) the flag is stored but NOT forwarded to reverse():
        # probability_flow=True leaves the predictors unchanged (pinned by tests/golden/plugins.npz)
        self.rsde = sde.reverse(score_fn)
        self.score_fn = score_fn
        self.probability_flow = probability_flow

    @abc.abstractmethod
    def update_fn(self, x, t, *args, **kwargs):
        """One update of the predictor: returns (next state, next state without noise)."""

    def debug_update_fn(self, x, t, *args):
        raise NotImplementedError(f"Debug update function not implemented for predictor {self}.")


@PredictorRegistry.register("euler_maruyama")
class EulerMaruyamaPredictor(Predictor):
    """x_mean = x - [f - g^2 score] dt ; x = x_mean + g sqrt(dt) z, with dt = 1/N regardless of the
    ``dt`` keyword (the reference reads it with ``getattr`` on a dict, predictors.py:45)."""

    def update_fn(self, x, t, *args, **kwargs):
        return self.rsde.step(x, t, *args)


@PredictorRegistry.register("reverse_diffusion")
class ReverseDiffusionPredictor(Predictor):
    """Uses ``rsde.discretize`` (predictors.py:60-66), which for these SDEs is the base-class
    Euler-Maruyama discretisation (sdes.py:93-107): algebraically the same update."""

    def update_fn(self, x, t, *args, **kwargs):
        return self.rsde.step(x, t, *args)


@PredictorRegistry.register("none")
class NonePredictor(Predictor):
    """An empty predictor that does nothing."""

    def __init__(self, *args, **kwargs):
        pass

    def update_fn(self, x, t, *args, **kwargs):
        return x, x
