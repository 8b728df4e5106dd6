"""``DiffSepModel``: the inference facade of the reference's LightningModule (``pl_model.py``).

Keeps what the separation path uses — ``forward(xt, time, mix)`` (:407-409), ``normalize_batch`` /
``denormalize_batch`` (:81-92), ``get_pc_sampler(predictor, corrector, y, N, minibatch, schedule,
**kw)`` (:687-759), ``separate`` (:148-164), ``load_from_checkpoint`` incl. the EMA swap-in
(:642-670) — without importing lightning / hydra / omegaconf.  Training (losses, optimiser, EMA
updates) is out of scope.
"""
from __future__ import annotations

import math
import warnings
from types import SimpleNamespace

import torch

from . import ops, sdes
from .score_model import ScoreModelNCSNpp
from .sdes.sdes import SDERegistry

_SDE_TARGETS = {"sdes.sdes.MixSDE": "mix", "sdes.sdes.PriorMixSDE": "priormix",
                "sdes.MixSDE": "mix", "sdes.PriorMixSDE": "priormix"}


def checkpoint_state(ckpt, no_ema=False):
    """``(config, score-model state_dict)`` of a loaded Lightning ``.ckpt`` / HF ``checkpoint.pt`` dict, with the EMA
    weights swapped in as the reference does on ``.eval()`` (pl_model.py:650-670; ``no_ema=True`` is its
    ``eval(no_ema=True)``).  ``checkpoint["ema"]`` is ``torch_ema.ExponentialMovingAverage.state_dict()``
    (``decay``, ``num_updates``, ``shadow_params``, ``collected_params``; pl_model.py:672-673); ``shadow_params`` is a
    list in ``parameters()`` order (output_layer first, then all_modules: ncsnpp.py:105,308 — the order of the
    ``state_dict`` keys).  The reference environment does not pin torch_ema, and its versions differ in what the
    list holds: <= 0.2 tracks only ``requires_grad`` parameters (the frozen Fourier ``W`` is absent), 0.3 tracks every
    entry of ``parameters()`` (``W`` included, equal to the raw ``W``).  Both layouts are accepted; any other count
    is an error — silently evaluating raw weights where the reference evaluates EMA weights is not."""
    config = ckpt.get("hyper_parameters", {}).get("config")
    sd = {k[len("score_model."):]: v for k, v in ckpt["state_dict"].items() if k.startswith("score_model.")}
    ema = ckpt.get("ema")
    if ema is None:
        warnings.warn("EMA state_dict not found in checkpoint!")       # the reference's message, pl_model.py:647
        return config, sd
    if no_ema or not ema.get("shadow_params"):
        return config, sd
    shadow = ema["shadow_params"]
    all_names = [k for k in sd if k.startswith("backbone.")]
    trainable = [k for k in all_names if not k.endswith("all_modules.0.W")]
    if len(shadow) == len(trainable):
        names = trainable
    elif len(shadow) == len(all_names):
        names = all_names
    else:
        raise ValueError(f"checkpoint holds {len(shadow)} EMA shadow parameters; the score model has "
                         f"{len(trainable)} trainable / {len(all_names)} parameter tensors (pass no_ema=True to "
                         "evaluate the raw weights)")
    for k, v in zip(names, shadow):
        if tuple(v.shape) != tuple(sd[k].shape):
            raise ValueError(f"EMA shadow parameter for {k} has shape {tuple(v.shape)}, "
                             f"the state_dict has {tuple(sd[k].shape)}")
        sd[k] = v
    return config, sd


def normalize_batch(batch):
    """(mix, tgt) -> ((mix', tgt'), mean, std): per-utterance zero mean / unit (unbiased) std with the
    std clamped at 1e-5 (pl_model.py:81-88), computed by ``dsep_normalize``."""
    mix, tgt = batch
    mix = mix.contiguous().float()
    B = mix.shape[0]
    n = mix[0].numel()
    out = torch.empty_like(mix)
    mean = torch.empty(B, 1, 1, device=mix.device, dtype=torch.float32)
    std = torch.empty(B, 1, 1, device=mix.device, dtype=torch.float32)
    ops.normalize(mix, B, n, out, mean, std)
    if tgt is not None:
        tgt = (tgt - mean) / std
    return (out, tgt), mean, std


def denormalize_batch(x, mean, std):
    return x * std + mean


def _ns(d):
    """dict tree -> attribute access (stands in for omegaconf's DictConfig)."""
    if isinstance(d, dict):
        return SimpleNamespace(**{k: _ns(v) for k, v in d.items()})
    return d


def _to_plain(cfg):
    """DictConfig / namespace / dict -> plain nested dict."""
    if isinstance(cfg, dict):
        return {k: _to_plain(v) for k, v in cfg.items()}
    if isinstance(cfg, SimpleNamespace):
        return {k: _to_plain(v) for k, v in vars(cfg).items()}
    if hasattr(cfg, "items") and not isinstance(cfg, (str, bytes)):
        return {k: _to_plain(v) for k, v in cfg.items()}
    if isinstance(cfg, (list, tuple)) or type(cfg).__name__ == "ListConfig":
        return [_to_plain(v) for v in cfg]
    return cfg


DEFAULT_CONFIG = {   # reference config/model/default.yaml + experiment/icassp-separation.yaml
    "model": {
        "fs": 8000, "n_speakers": 2, "t_eps": 0.03,
        "sde": {"_target_": "sdes.sdes.MixSDE", "ndim": 2, "d_lambda": 2.0, "sigma_min": 0.05,
                "sigma_max": 0.5, "N": 30},
        "score_model": {
            "_target_": "models.score_models.ScoreModelNCSNpp", "num_sources": 2,
            "stft_args": {"n_fft": 510, "hop_length": 128, "center": True, "pad_mode": "constant"},
            "backbone_args": {"_target_": "models.ncsnpp.NCSNpp", "nf": 128},
            "transform": "exponent", "spec_abs_exponent": 0.5, "spec_factor": 0.15,
            "spec_trans_learnable": False},
        "sampler": {"N": 30, "corrector_steps": 1, "snr": 0.5},
    }
}


class DiffSepModel(torch.nn.Module):
    def __init__(self, config=None, device="cuda", passes=None, score_state_dict=None):
        super().__init__()
        cfg = _to_plain(config) if config is not None else DEFAULT_CONFIG
        self.config = _ns(cfg)
        ops.use_device(device)       # kernels launch on the current device's stream: make `device` current
        m = cfg["model"]
        sm = dict(m["score_model"])
        sm.pop("_target_", None)
        self.score_model = ScoreModelNCSNpp(**sm, device=device, passes=passes, state_dict=score_state_dict)
        sde_cfg = dict(m["sde"])
        target = sde_cfg.pop("_target_", "sdes.sdes.MixSDE")
        if target not in _SDE_TARGETS:
            raise NotImplementedError(f"SDE '{target}' is not on the DiffSep hot path (MixSDE / PriorMixSDE)")
        self.sde = SDERegistry.get_by_name(_SDE_TARGETS[target])(**sde_cfg)
        self.t_eps = m.get("t_eps", 0.03)
        self.t_max = self.sde.T
        self.normalize_batch = normalize_batch
        self.denormalize_batch = denormalize_batch
        self.dev = torch.device(device)

    # ------------------------------------------------------------------ reference surface
    def forward(self, xt, time, mix):
        return self.score_model(xt, time, mix)

    def cached_mixture(self, mix):
        return self.score_model.cached_mixture(mix)

    # ------------------------------------------------------------------ training-side forward pieces (no autograd)
    def sample_time(self, x):
        """uniform in [t_eps, t_max] (pl_model.py:166-171, the default ``time_sampling_strategy``)."""
        return torch.empty(x.shape[0], device=x.device, dtype=torch.float32).uniform_(self.t_eps, self.t_max)

    def sample_prior(self, mix, target, time=None):
        """``(x_t, time, L, z)`` of the reference's ``sample_prior`` with the default ``init_hack = false``
        (pl_model.py:179-188, 243-247): x_t = mean + L z on ``sde.marginal_prob(target, time, mix)``.  ``L`` is returned
        as the pair (time, mix) it is a function of — ``compute_score_loss`` applies it inside its kernel instead of
        materialising the [B,n,n(,T)] tensor."""
        if getattr(self.config.model, "init_hack", False):
            raise NotImplementedError("sample_prior: only the default init_hack = false is built")
        time = self.sample_time(target) if time is None else time
        x_t, z = self.sde.marginal_sample(target, time, mix)
        return x_t, time, (time, mix), z

    def compute_score_loss(self, mix, target, time=None, reduction="mean"):
        """The score-matching loss of the reference's ``compute_score_loss`` (pl_model.py:411-424), FORWARD ONLY — a
        validation metric on this inference path (its kernels have no backward): perturb (one kernel), score network,
        ``MSE(L score, -z)`` (one kernel).  ``reduction``: "mean" (MSELoss's default, config/model/default.yaml:44-45) or
        "none" (per sample, :421-422)."""
        x_t, time, _, z = self.sample_prior(mix, target, time)
        pred_score = self(x_t, time, mix)
        loss = self.sde.score_loss(pred_score, z, time, mix)
        return loss.mean() if reduction == "mean" else loss

    def prepare_times(self, ts):
        return self.score_model.prepare_times(ts)

    def uniform_time(self, t):
        return self.score_model.uniform_time(t)

    def get_pc_sampler(self, predictor_name, corrector_name, y, N=None, minibatch=None, schedule=None,
                       **kwargs):
        N = self.sde.N if N is None else N
        sde = self.sde.copy()
        sde.N = N
        kwargs = {"eps": self.t_eps, **kwargs}

        def make(y_):
            if schedule is None:
                return sdes.get_pc_sampler(predictor_name, corrector_name, sde=sde, score_fn=self, y=y_, **kwargs)
            return sdes.get_pc_scheduled_sampler(predictor_name, corrector_name, sde=sde, score_fn=self, y=y_,
                                                 schedule=schedule, **kwargs)

        if minibatch is None:
            return make(y)
        M = y.shape[0]

        def batched_sampling_fn():
            samples, ns, intmet = [], [], []
            for i in range(int(math.ceil(M / minibatch))):
                y_mini = y[i * minibatch:(i + 1) * minibatch].contiguous()
                sample, n, *other = make(y_mini)()
                samples.append(sample)
                ns.append(n)
                if other:
                    intmet.append(other[0])
            samples = torch.cat(samples, dim=0)
            return (samples, ns, intmet) if intmet else (samples, ns)

        return batched_sampling_fn

    def separate(self, mix, **kwargs):
        """mix [B,1,T] -> (estimate [B,2,T], nfe) in the normalised domain, like the reference, whose
        ``separate`` returns a fresh ``sampler()`` call and drops the de-normalised one (:148-164)."""
        (mix, _), *stats = self.normalize_batch((mix, None))
        sampler_kwargs = dict(vars(self.config.model.sampler)) if hasattr(self.config.model, "sampler") else {}
        sampler_kwargs.update(kwargs)
        sampler = self.get_pc_sampler("reverse_diffusion", "ald2", mix, **sampler_kwargs)
        return sampler()

    # ------------------------------------------------------------------ checkpoints
    @classmethod
    def load_from_checkpoint(cls, path, map_location=None, device="cuda", passes=None, no_ema=False, **kwargs):
        """Reads a Lightning ``.ckpt`` / HF ``checkpoint.pt``: ``hyper_parameters.config``,
        ``state_dict`` (``score_model.*``), and — because the reference swaps EMA weights in on
        ``.eval()`` (pl_model.py:650-670) — ``ema.shadow_params`` in ``parameters()`` order."""
        ckpt = torch.load(path, map_location="cpu", weights_only=False)
        config, sd = checkpoint_state(ckpt, no_ema=no_ema)
        model = cls(config, device=device, passes=passes, score_state_dict=sd)
        return model
