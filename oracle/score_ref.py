"""ORACLE (test infrastructure, not product code): the waveform <-> spectrogram wrapper
around the backbone, restating ``ScoreModelNCSNpp`` (reference
``models/score_models.py:10-138``).

torchaudio's ``Spectrogram(power=None)`` / ``InverseSpectrogram`` are thin wrappers over
``torch.stft`` / ``torch.istft`` with a periodic Hann window of length ``n_fft``
(torchaudio default); those are called directly here.  ``stft_matrix`` /
``istft_matrix`` are an independent float64 DFT-matrix formulation used to pin the frame
indexing (SURVEY.md §7 hard-part 4).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from .ncsnpp_ref import ncsnpp_forward

N_FFT = 510      # reference config/model/default.yaml:18-22
HOP = 128


def n_frames(T, n_fft=N_FFT, hop=HOP):
    return 1 + (T + (n_fft - hop)) // hop


def pre_process(x, spec_factor=0.15, spec_abs_exponent=0.5, n_fft=N_FFT, hop=HOP):
    """reference score_models.py:107-116 (+ :41-48, :72-76, :83-91).

    x: [B, C, T] real -> ([B, 2C, n_fft//2+1, W] real, n_samples, n_pad)."""
    n_samples = x.shape[-1]
    x = F.pad(x, (0, n_fft - hop))
    B, C, Tp = x.shape
    win = torch.hann_window(n_fft, dtype=x.dtype)
    spec = torch.stft(x.reshape(B * C, Tp), n_fft=n_fft, hop_length=hop, win_length=n_fft,
                      window=win, center=True, pad_mode="constant", normalized=False,
                      onesided=True, return_complex=True)
    spec = spec.reshape(B, C, spec.shape[-2], spec.shape[-1])
    if spec_abs_exponent != 1:
        e = abs(spec_abs_exponent)
        spec = spec.abs() ** e * torch.exp(1j * spec.angle())
    spec = spec * spec_factor
    xr = torch.stack((spec.real, spec.imag), dim=1).flatten(1, 2)   # [re_0..re_{C-1}, im_0..]
    rem = xr.shape[-1] % 64
    n_pad = 0 if rem == 0 else 64 - rem
    if n_pad:
        xr = F.pad(xr, (0, n_pad))
    return xr, n_samples, n_pad


def post_process(x, n_samples, n_pad, spec_factor=0.15, spec_abs_exponent=0.5,
                 n_fft=N_FFT, hop=HOP):
    """reference score_models.py:118-124 (+ :59-64, :78-81, :99-105)."""
    if n_pad:
        x = x[..., :-n_pad]
    x = x.reshape((x.shape[0], 2, -1) + x.shape[2:])
    spec = torch.view_as_complex(x.moveaxis(1, -1).contiguous())
    spec = spec / abs(spec_factor)
    if spec_abs_exponent != 1:
        e = abs(spec_abs_exponent)
        spec = spec.abs() ** (1 / e) * torch.exp(1j * spec.angle())
    B, C, Fq, Fr = spec.shape
    win = torch.hann_window(n_fft, dtype=x.dtype)
    y = torch.istft(spec.reshape(B * C, Fq, Fr), n_fft=n_fft, hop_length=hop, win_length=n_fft,
                    window=win, center=True, normalized=False, onesided=True)
    y = y.reshape(B, C, -1)
    if y.shape[-1] < n_samples:
        y = F.pad(y, (0, n_samples - y.shape[-1]))
    elif y.shape[-1] > n_samples:
        y = y[..., :n_samples]
    return y


def score_forward(params, xt, t, mix, spec_factor=0.15, spec_abs_exponent=0.5, taps=None):
    """ScoreModelNCSNpp.forward, reference score_models.py:126-138.

    ``params`` holds backbone parameters keyed relative to the backbone."""
    x = torch.cat((xt, mix), dim=1)
    x, n_samples, n_pad = pre_process(x, spec_factor, spec_abs_exponent)
    if taps is not None:
        taps["spec_in"] = x
    x = ncsnpp_forward(params, x, t, taps=taps)
    if taps is not None:
        taps["spec_out"] = x
    return post_process(x, n_samples, n_pad, spec_factor, spec_abs_exponent)


# ---- independent float64 DFT-matrix formulation (pins the frame indexing) -------------

def stft_matrix(x, n_fft=N_FFT, hop=HOP):
    """Frame k covers samples [hop*k - n_fft//2, hop*k + n_fft//2) of the signal right-padded
    by n_fft-hop zeros, zero outside, times periodic Hann(n_fft); one-sided DFT.
    x: [..., T] -> complex128 [..., n_fft//2+1, Fr]."""
    x = x.to(torch.float64)
    T = x.shape[-1]
    Fr = n_frames(T, n_fft, hop)
    half = n_fft // 2
    xp = F.pad(x, (half, (n_fft - hop) + half))
    idx = (torch.arange(Fr)[:, None] * hop + torch.arange(n_fft)[None, :])
    frames = xp[..., idx]                                  # [..., Fr, n_fft]
    n = torch.arange(n_fft, dtype=torch.float64)
    w = 0.5 - 0.5 * torch.cos(2 * math.pi * n / n_fft)
    k = torch.arange(n_fft // 2 + 1, dtype=torch.float64)
    ang = -2 * math.pi * k[:, None] * n[None, :] / n_fft
    basis = torch.complex(torch.cos(ang), torch.sin(ang))  # [bins, n_fft]
    fw = (frames * w).to(torch.complex128)
    return torch.einsum("...fn,kn->...kf", fw, basis)


def istft_matrix(spec, n_fft=N_FFT, hop=HOP):
    """Inverse of the above as torch.istft defines it: irfft per frame (imaginary parts of
    DC/Nyquist ignored), times window, overlap-add, divided by the sum of squared windows,
    cropped by n_fft//2 on both sides -> length hop*(Fr-1)."""
    spec = spec.to(torch.complex128)
    Fq, Fr = spec.shape[-2], spec.shape[-1]
    n = torch.arange(n_fft, dtype=torch.float64)
    w = 0.5 - 0.5 * torch.cos(2 * math.pi * n / n_fft)
    frames = torch.fft.irfft(spec.transpose(-1, -2), n=n_fft, dim=-1) * w   # [..., Fr, n_fft]
    L = n_fft + hop * (Fr - 1)
    y = torch.zeros(spec.shape[:-2] + (L,), dtype=torch.float64)
    env = torch.zeros(L, dtype=torch.float64)
    for f in range(Fr):
        y[..., f * hop:f * hop + n_fft] += frames[..., f, :]
        env[f * hop:f * hop + n_fft] += w * w
    half = n_fft // 2
    y = y[..., half:half + hop * (Fr - 1)]
    env = env[half:half + hop * (Fr - 1)]
    return y / env
