"""ORACLE (test infrastructure, not product code): closed-form restatement of the DiffSep
SDE arithmetic and of the predictor-corrector sampler loop.

Restates (reference file:line):
  * MixSDE            sdes/sdes.py:180-349   (sde :275-284, _cov_eigval :296-310, _std :316-320,
                                               prior_sampling :334-346)
  * PriorMixSDE       sdes/sdes.py:352-590   (_std_sigma_mix :477-489, sde :451-470, _std :515-528,
                                               prior_sampling :564-587)
  * SDE.discretize / RSDE.discretize          sdes/sdes.py:93-107, 163-171
  * ReverseDiffusionPredictor.update_fn       sdes/predictors.py:60-66
  * AnnealedLangevinDynamics2.update_fn       sdes/correctors.py:109-128
  * get_pc_sampler / get_pc_scheduled_sampler sdes/__init__.py:46-190

With ``A`` the channel-averaging matrix and ``Pn = I - A`` (sdes.py:242-248), for any
``v`` [B, C, T] and channel mean ``vbar``:  ``A v = vbar``, ``Pn v = v - vbar`` and
``(a A + b Pn) v = a vbar + b (v - vbar)``.

Noise is *injected* (a list of pre-drawn tensors popped in draw order: prior, then per step
[corrector noise]*n_steps, predictor z) because ``randn_like`` on the reference's strided
tensors does not map seeds to elements reproducibly.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


class MixSDEParams:
    def __init__(self, ndim=2, d_lambda=2.0, sigma_min=0.05, sigma_max=0.5, N=30,
                 prior=False, avg_len=510):
        self.ndim, self.d_lambda = ndim, d_lambda
        self.sigma_min, self.sigma_max = sigma_min, sigma_max
        self.ratiosig = sigma_max / sigma_min
        self.logsig = math.log(self.ratiosig)
        self.N = N
        self.T = 1.0
        self.prior = prior            # True -> PriorMixSDE
        self.avg_len = avg_len


def cov_eigval(p, t):
    """sdes.py:296-310 / 497-511."""
    mult = p.sigma_min ** 2
    srp = p.ratiosig ** (2 * t)
    ev1 = mult * (srp - 1)
    ev2 = mult * (srp - torch.exp(-2.0 * p.d_lambda * t)) / (1.0 + p.d_lambda / p.logsig)
    return ev1, ev2


def sigma_mix(p, mix):
    """PriorMixSDE._std_sigma_mix, sdes.py:477-489."""
    s = F.avg_pool1d(mix ** 2, kernel_size=p.avg_len, stride=1, padding=p.avg_len // 2)
    s = s.clamp(min=1e-4).sqrt()
    if p.avg_len % 2 == 0:
        s = s[..., :-1]
    return 0.5 * s


def mult_L(p, t, v, mix=None):
    """L(t) v with L = sqrt(ev1) A + sqrt(ev2) Pn [* sigma_mix]; sdes.py:316-328 / 515-532."""
    ev1, ev2 = cov_eigval(p, t)
    vbar = v.mean(dim=1, keepdim=True)
    out = ev1.sqrt()[:, None, None] * vbar + ev2.sqrt()[:, None, None] * (v - vbar)
    if p.prior:
        out = out * sigma_mix(p, mix)
    return out


def prior_sampling(p, mix, z):
    """x_T = mean + L(T) z; sdes.py:334-346 (MixSDE: mean = 0.5 mix broadcast to TWO channels, hard-coded) /
    564-587 (PriorMixSDE: mean = 0.5 mix broadcast to ndim for a 1-channel input, the input itself for an
    ndim-channel one — the sampler's ``true_mean`` — and L = (sqrt(ev1) A + sqrt(ev2) Pn)[c,d] sigma_mix[d],
    :515-528: sigma enters per INPUT channel d, which for a 1-channel mixture is the plain scaling of mult_L)."""
    t = torch.ones(mix.shape[0], dtype=mix.dtype) * p.T
    if not p.prior:
        mean = torch.broadcast_to(0.5 * mix, (mix.shape[0], 2, mix.shape[2]))
        return mean + mult_L(p, t, z, mix)
    if mix.shape[1] == p.ndim:
        mean = mix
    elif mix.shape[1] == 1:
        mean = torch.broadcast_to(0.5 * mix, (mix.shape[0], p.ndim, mix.shape[2]))
    else:
        raise ValueError("The input provided to prior_sampling should have 1 channel, or the same as the number of "
                         f"speakers. Found {mix.shape[1]} channels instead.")
    ev1, ev2 = cov_eigval(p, t)
    w = z * sigma_mix(p, mix)
    wbar = w.mean(dim=1, keepdim=True)
    return mean + ev1.sqrt()[:, None, None] * wbar + ev2.sqrt()[:, None, None] * (w - wbar)


def marginal_mean(p, x0, t):
    """mean of p_t(x | x0) = (A + e^{-lambda t} Pn) x0; sdes.py:286-294 / 472-494."""
    decay = torch.exp(-t * p.d_lambda)[:, None, None]
    xbar = x0.mean(dim=1, keepdim=True)
    return xbar + decay * (x0 - xbar)


def sample_prior(p, mix, target, time, z):
    """DiffSepModel.sample_prior with the default ``init_hack = false`` (pl_model.py:179-188, 243-247):
    x_t = mean(target, t) + L(t) z with (mean, L) = sde.marginal_prob(target, t, mix) (sdes.py:322-324 / 560-562).
    ``time`` and ``z`` are given (the reference draws them with uniform_ / randn_like)."""
    return marginal_mean(p, target, time) + mult_L(p, time, z, mix)


def score_loss(p, score, z, time, mix=None, reduction="mean"):
    """DiffSepModel.compute_score_loss after the network call (pl_model.py:418-424): MSE between L(t) score and -z;
    ``reduction="none"`` gives the per-sample mean over (channel, time) of :421-422."""
    e = (mult_L(p, time, score, mix) + z) ** 2
    return e.mean() if reduction == "mean" else e.mean(dim=(-2, -1))


def diffusion(p, t, mix=None):
    """g(t) = sigma_min ratiosig^t sqrt(2 logsig) [* sigma_mix]; sdes.py:282-283 / 465-469."""
    g = p.sigma_min * p.ratiosig ** t * math.sqrt(2 * p.logsig)
    g = g[:, None, None]
    if p.prior:
        g = g * sigma_mix(p, mix)
    return g


def predictor_step(p, score_fn, x, t, mix, z):
    """reverse_diffusion predictor: predictors.py:60-66 with RSDE.discretize sdes.py:163-171.

    dt is always 1/N (``getattr(kwargs, "dt", ...)`` on a dict)."""
    dt = 1.0 / p.N
    xbar = x.mean(dim=1, keepdim=True)
    drift = -p.d_lambda * (x - xbar)
    f = drift * dt
    G = diffusion(p, t, mix) * math.sqrt(dt)
    rev_f = f - G ** 2 * score_fn(x, t, mix)
    x_mean = x - rev_f
    return x_mean + G * z, x_mean


def corrector_step(p, score_fn, x, t, mix, noises, snr):
    """ald2: correctors.py:109-128.  ``noises``: list of n_steps tensors."""
    x_mean = x
    for nz in noises:
        grad = score_fn(x, t, mix)
        g2 = mult_L(p, t, mult_L(p, t, grad, mix), mix)
        x_mean = x + 2 * snr ** 2 * g2
        x = x_mean + 2 * snr * mult_L(p, t, nz, mix)
    return x, x_mean


def ald_step(p, score_fn, x, t, mix, noises, snr):
    """ald: correctors.py:58-91 (MixSDE only).  std = sqrt(first-row sum of L L) = sqrt(ev1)."""
    ev1, _ = cov_eigval(p, t)
    std = ev1.sqrt()[:, None, None]
    x_mean = x
    for nz in noises:
        grad = score_fn(x, t, mix)
        step = (snr * std) ** 2 * 2
        x_mean = x + step * grad
        x = x_mean + nz * torch.sqrt(step * 2)
    return x, x_mean


def langevin_step(p, score_fn, x, t, mix, noises, snr):
    """langevin: correctors.py:35-55 (batch-mean norms couple the batch entries)."""
    x_mean = x
    for nz in noises:
        grad = score_fn(x, t, mix)
        grad_norm = torch.norm(grad.reshape(grad.shape[0], -1), dim=-1).mean()
        noise_norm = torch.norm(nz.reshape(nz.shape[0], -1), dim=-1).mean()
        step = (snr * noise_norm / grad_norm) ** 2 * 2
        x_mean = x + step * grad
        x = x_mean + nz * torch.sqrt(step * 2)
    return x, x_mean


def pc_sampler_plugins(p, score_fn, mix, noises, predictor="reverse_diffusion", corrector="ald2", eps=0.03, snr=0.5,
                       corrector_steps=1, denoise=True):
    """sdes/__init__.py:166-190 over the other registered plugins.  ``euler_maruyama``
    is algebraically ``reverse_diffusion``; ``probability_flow`` has no effect in the reference (the
    flag never reaches ``reverse()``, predictors.py:13-18), so it is not a parameter here."""
    noises = list(noises)
    xt = prior_sampling(p, mix, noises.pop(0))
    ts = timesteps(p, eps, None, mix.dtype)
    corr = {"ald2": corrector_step, "ald": ald_step, "langevin": langevin_step}[corrector]
    xt_mean = xt
    for i in range(p.N):
        vec_t = torch.ones(mix.shape[0], dtype=mix.dtype) * ts[i]
        cn = [noises.pop(0) for _ in range(corrector_steps)]
        xt, xt_mean = corr(p, score_fn, xt, vec_t, mix, cn, snr)
        if predictor == "none":
            xt_mean = xt
        else:
            xt, xt_mean = predictor_step(p, score_fn, xt, vec_t, mix, noises.pop(0))
    return xt_mean if denoise else xt


def timesteps(p, eps, schedule=None, dtype=torch.float32):
    """sdes/__init__.py:175 (plain) and :92-111 (scheduled; N+1 points, dt unchanged)."""
    if schedule is None:
        return torch.linspace(p.T, eps, p.N, dtype=dtype)
    if schedule == "linear":
        return torch.linspace(p.T, eps, p.N + 1, dtype=dtype)
    if schedule == "log":
        return torch.logspace(math.log(p.T) / math.log(10), math.log(eps) / math.log(10),
                              p.N + 1, base=10, dtype=dtype)
    if schedule == "revlog":
        return torch.logspace(math.log(eps) / math.log(10), math.log(p.T) / math.log(10),
                              p.N + 1, base=10, dtype=dtype).flip(dims=(0,))
    raise NotImplementedError(f"Schedule '{schedule}' does not exist")


def pc_sampler(p, score_fn, mix, noises, eps=0.03, snr=0.5, corrector_steps=1,
               denoise=True, schedule=None, intermediate=False, true_mean=None):
    """sdes/__init__.py:166-190.  ``noises``: flat list in draw order.  ``true_mean`` replaces the mixture in the
    prior only (:171-174)."""
    noises = list(noises)
    xt = prior_sampling(p, mix if true_mean is None else true_mean, noises.pop(0))
    ts = timesteps(p, eps, schedule, mix.dtype)
    im = []
    xt_mean = xt
    for i in range(p.N):
        vec_t = torch.ones(mix.shape[0], dtype=mix.dtype) * ts[i]
        cn = [noises.pop(0) for _ in range(corrector_steps)]
        xt, xt_mean_c = corrector_step(p, score_fn, xt, vec_t, mix, cn, snr)
        if intermediate:
            im.append((xt, xt_mean_c))
        xt, xt_mean = predictor_step(p, score_fn, xt, vec_t, mix, noises.pop(0))
    out = xt_mean if denoise else xt
    nfe = p.N * (corrector_steps + 1)
    return (out, nfe, im) if intermediate else (out, nfe)


def normalize_batch(mix):
    """pl_model.py:81-88 (unbiased std, clamp 1e-5)."""
    mean = mix.mean(dim=(1, 2), keepdim=True)
    std = mix.std(dim=(1, 2), keepdim=True).clamp(min=1e-5)
    return (mix - mean) / std, mean, std


def scale_output(mix, sep):
    """separate.py:73-78."""
    num = (mix * sep).sum(dim=-1, keepdim=True)
    den = (sep * sep + 1e-10).sum(dim=-1, keepdim=True)
    return num / den * sep
