"""Predictor-corrector samplers of DiffSep (reference ``sdes/__init__.py:46-190``)."""
from __future__ import annotations

import contextlib
import math

import torch

from .correctors import Corrector, CorrectorRegistry
from .noise import injected_noise
from .predictors import Predictor, PredictorRegistry, ReverseDiffusionPredictor
from .sdes import SDE, MixSDE, PriorMixSDE, SDERegistry

__all__ = ["PredictorRegistry", "CorrectorRegistry", "SDERegistry", "Predictor", "Corrector", "MixSDE",
           "PriorMixSDE", "get_pc_sampler", "get_pc_scheduled_sampler", "injected_noise"]


def _timesteps(sde, eps, schedule, device):
    """sdes/__init__.py:175 (N points) and :92-111 (scheduled variants, N + 1 points)."""
    if schedule is None:
        return torch.linspace(sde.T, eps, sde.N, device=device)
    base = 10
    if schedule == "linear":
        return torch.linspace(sde.T, eps, sde.N + 1, device=device)
    if schedule == "log":
        return torch.logspace(math.log(sde.T) / math.log(base), math.log(eps) / math.log(base), sde.N + 1,
                              base=base, device=device)
    if schedule == "revlog":
        return torch.logspace(math.log(eps) / math.log(base), math.log(sde.T) / math.log(base), sde.N + 1,
                              base=base, device=device).flip(dims=(0,))
    raise NotImplementedError(f"Schedule '{schedule}' does not exist")


def _make_sampler(predictor_name, corrector_name, sde, score_fn, y, true_mean, denoise, eps, snr,
                  corrector_steps, probability_flow, intermediate, schedule):
    predictor_cls = PredictorRegistry.get_by_name(predictor_name)
    corrector_cls = CorrectorRegistry.get_by_name(corrector_name)
    predictor = predictor_cls(sde, score_fn, probability_flow=probability_flow)
    corrector = corrector_cls(sde, score_fn, snr=snr, n_steps=corrector_steps)

    def pc_sampler():
        """The PC sampler function: -> (x, nfe[, intermediates])."""
        im = []
        cond = true_mean if true_mean is not None else y
        # the mixture spectrogram is constant over the run: let the score model keep it
        cache = getattr(score_fn, "cached_mixture", None)
        if hasattr(sde, "reset_cache"):
            sde.reset_cache()          # sigma_mix of a previous mixture must not survive into this run
        with torch.no_grad(), (cache(y) if cache is not None else contextlib.nullcontext()):
            xt = sde.prior_sampling(cond.shape, cond)
            # one host read of the grid: every batch entry shares t (sdes/__init__.py:177-178)
            ts = _timesteps(sde, eps, schedule, "cpu").tolist()
            vec_t = torch.empty(y.shape[0], device=y.device, dtype=torch.float32)
            # the time embedding and the FiLM projections depend on t only and the grid is known: evaluate them for
            # all N times now, and tell the score model which (shared) time each step is at
            prepare = getattr(score_fn, "prepare_times", None)
            uniform = getattr(score_fn, "uniform_time", None)
            if prepare is not None:
                prepare(ts[:sde.N])
            xt_mean = xt
            for i in range(sde.N):
                vec_t.fill_(ts[i])
                with (uniform(ts[i]) if uniform is not None else contextlib.nullcontext()):
                    xt, xt_mean = corrector.update_fn(xt, vec_t, y)
                    if intermediate:
                        im.append((xt, xt_mean))
                    xt, xt_mean = predictor.update_fn(xt, vec_t, y)
            x_result = xt_mean if denoise else xt
            ns = sde.N * (corrector.n_steps + 1)
            return (x_result, ns, im) if intermediate else (x_result, ns)

    return pc_sampler


def get_pc_sampler(predictor_name, corrector_name, sde, score_fn, y, true_mean=None, denoise=True, eps=3e-2,
                   snr=0.1, corrector_steps=1, probability_flow: bool = False, intermediate=False, **kwargs):
    """Create a Predictor-Corrector sampler (same signature as the reference, sdes/__init__.py:132-146)."""
    return _make_sampler(predictor_name, corrector_name, sde, score_fn, y, true_mean, denoise, eps, snr,
                         corrector_steps, probability_flow, intermediate, None)


def get_pc_scheduled_sampler(predictor_name, corrector_name, sde, score_fn, y, denoise=True, true_mean=None,
                             eps=3e-2, snr=0.1, corrector_steps=1, probability_flow: bool = False,
                             intermediate=False, schedule="linear", **kwargs):
    """Scheduled time grid (N + 1 points); the step size stays 1/N as in the reference
    (``getattr(kwargs, "dt", ...)`` on a dict, sdes/sdes.py:103)."""
    _timesteps(sde, eps, schedule, "cpu")   # unknown schedule -> NotImplementedError up front
    return _make_sampler(predictor_name, corrector_name, sde, score_fn, y, true_mean, denoise, eps, snr,
                         corrector_steps, probability_flow, intermediate, schedule)
