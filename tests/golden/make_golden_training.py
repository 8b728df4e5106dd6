"""Golden vectors for the training-side forward pieces: the REAL reference's
``sde.marginal_prob`` / ``mult_std`` (sdes/sdes.py:322-328, 531-532, 560-562) and the arithmetic of
``DiffSepModel.sample_prior`` (default init_hack) / ``compute_score_loss`` (pl_model.py:179-247, 411-424) around them,
with given times, injected noise and the analytic score of cases.py in place of the network.  Build container only.

    python tests/golden/make_golden_training.py
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
from make_golden import import_reference  # noqa: E402

B, T = 3, 1024


def main():
    R = import_reference()
    from oracle import sde_ref as sd
    out = {}
    for name, cls, ndim in (("mix", R["MixSDE"], 2), ("priormix", R["PriorMixSDE"], 2), ("priormix3", R["PriorMixSDE"], 3)):
        sde = cls(ndim=ndim, d_lambda=2.0, sigma_min=0.05, sigma_max=0.5, N=30)
        g = cases.gen(101 + ndim)
        target = torch.randn(B, ndim, T, generator=g) * 0.3
        mix = target.sum(dim=1, keepdim=True)
        time = torch.tensor([0.03, 0.4711, 1.0])
        z = torch.randn(B, ndim, T, generator=g)
        mean, L = sde.marginal_prob(target, time, mix)
        x_t = mean + sde.mult_std(L, z)                      # pl_model.py:247
        score = torch.randn(B, ndim, T, generator=g) * 2.0    # stands for self(x_t, time, mix)
        L_score = sde.mult_std(L, score)                      # pl_model.py:419
        err = torch.nn.MSELoss(reduction="none")(L_score, -z)
        out.update({f"{name}_target": target.numpy(), f"{name}_time": time.numpy(), f"{name}_z": z.numpy(),
                    f"{name}_score": score.numpy(), f"{name}_mean": mean.numpy(), f"{name}_xt": x_t.numpy(),
                    f"{name}_loss_none": err.mean(dim=(-2, -1)).numpy(),
                    f"{name}_loss_mean": torch.nn.MSELoss()(L_score, -z).numpy()})
        print(name, tuple(L.shape), float(x_t.abs().mean()), float(err.mean()))
    np.savez_compressed(HERE / "training.npz", **out)


if __name__ == "__main__":
    main()
