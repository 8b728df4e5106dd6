"""Golden vectors for 3-source sampling and for the sampler's ``true_mean`` prior, produced by the
REAL reference sampler (``sdes.get_pc_sampler`` over ``PriorMixSDE(ndim=3)`` / ``MixSDE``) with the analytic score of
cases.py and injected noise.  Build container only.

    python tests/golden/make_golden_ndim.py
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(HERE.parent.parent))

import cases  # noqa: E402
from make_golden import NoiseInjector, import_reference  # noqa: E402

CASES, N, B, T = cases.NDIM_CASES, cases.NDIM_N, cases.NDIM_B, cases.NDIM_T
noises_for, true_mean_for = cases.ndim_noises, cases.ndim_true_mean


def main():
    R = import_reference()
    from oracle import sde_ref as sd
    mix, _, _ = sd.normalize_batch(cases.batch_mix(B, T))
    out = {}
    for name, sde_name, ndim, tm_ch, cs in CASES:
        cls = R["MixSDE"] if sde_name == "mix" else R["PriorMixSDE"]
        sde = cls(ndim=ndim, d_lambda=2.0, sigma_min=0.05, sigma_max=0.5, N=N)
        tm = true_mean_for(tm_ch) if tm_ch else None
        with NoiseInjector(noises_for(ndim, cs)):
            x, nfe = R["sdes"].get_pc_sampler("reverse_diffusion", "ald2", sde=sde, score_fn=cases.analytic_score,
                                              y=mix, true_mean=tm, eps=0.03, snr=0.5, corrector_steps=cs,
                                              denoise=True)()
        out[name] = x.numpy()
        print(name, nfe, tuple(x.shape), float(x.abs().mean()))
    np.savez_compressed(HERE / "ndim.npz", **out)


if __name__ == "__main__":
    main()
