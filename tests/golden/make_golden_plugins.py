"""Golden vectors for the remaining predictor / corrector plugins, produced by the REAL
reference sampler with the analytic score of cases.py and injected noise.  Build container only.

    python tests/golden/make_golden_plugins.py
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

CASES = [
    # name, predictor, corrector, sde, kwargs
    ("em_ald2", "euler_maruyama", "ald2", "mix", dict(corrector_steps=1)),
    ("rd_ald", "reverse_diffusion", "ald", "mix", dict(corrector_steps=2)),
    ("rd_langevin", "reverse_diffusion", "langevin", "mix", dict(corrector_steps=1)),
    ("rd_langevin_prior", "reverse_diffusion", "langevin", "priormix", dict(corrector_steps=1)),
    ("rd_ald2_pflow", "reverse_diffusion", "ald2", "mix", dict(corrector_steps=1, probability_flow=True)),
    ("em_ald2_pflow", "euler_maruyama", "ald2", "priormix", dict(corrector_steps=1, probability_flow=True)),
    ("none_ald2", "none", "ald2", "mix", dict(corrector_steps=1)),
]
N, B, T = 10, 2, 1024


def main():
    R = import_reference()
    from oracle import sde_ref as sd
    mix, _, _ = sd.normalize_batch(cases.batch_mix(B, T))
    out = {}
    for name, pred, corr, sde_name, kw in CASES:
        cls = R["MixSDE"] if sde_name == "mix" else R["PriorMixSDE"]
        sde = cls(ndim=2, d_lambda=2.0, sigma_min=0.05, sigma_max=0.5, N=N)
        cs = kw["corrector_steps"]
        noises = cases.sampler_noises(B, T, N, cs)
        with NoiseInjector(noises):
            x, nfe = R["sdes"].get_pc_sampler(pred, corr, sde=sde, score_fn=cases.analytic_score, y=mix, eps=0.03,
                                              snr=0.5, denoise=False, **kw)()
        out[name] = x.numpy()
        print(name, nfe, float(x.abs().mean()))
    np.savez_compressed(HERE / "plugins.npz", **out)


if __name__ == "__main__":
    main()
