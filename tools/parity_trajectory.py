"""Parity over a whole sampling run of the benchmark architecture (nf=128, 4 s @ 8 kHz, N=30, 1 corrector
step, injected noise): (a) per-step error with the CUDA path re-started from the oracle's state at every
update (the north-star "per-step output within 1e-4"), (b) the free-running final estimate (two fp32
implementations of a 60-evaluation stochastic recursion drift apart; reported for the record).
Test infrastructure: imports oracle/.  ~3 min of CPU oracle time.

    python tools/parity_trajectory.py >> profiles/parity_rXX.md
"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests" / "golden"))

import torch  # noqa: E402

import cases  # noqa: E402
from diffsep_b200 import sdes  # noqa: E402
from diffsep_b200.pl_model import DEFAULT_CONFIG, DiffSepModel, normalize_batch  # noqa: E402
from oracle import score_ref as sr, sde_ref as sd, weights as ow  # noqa: E402

N, T, NF = int(sys.argv[1]) if len(sys.argv) > 1 else 30, 32000, 128
DEV = "cuda"


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm())


PASSES = int(__import__("os").environ.get("DSEP_PASSES", "2"))     # 2: the e4m3-correction variant (DSEP_LIB)
model = DiffSepModel(DEFAULT_CONFIG, passes=PASSES, score_state_dict=ow.make_score_model_state_dict(nf=NF, seed=0))
params = ow.make_backbone_params(nf=NF, seed=0)
mix_cpu, _, _ = sd.normalize_batch(cases.batch_mix(1, T))
(mix, _), _, _ = normalize_batch((cases.batch_mix(1, T).to(DEV), None))
noises = cases.sampler_noises(1, T, N, 1)
p = sd.MixSDEParams(N=N)
sde = sdes.MixSDE(2, 2.0, 0.05, 0.5, N=N)


def score_cpu(x, t, m):
    with torch.no_grad():
        return sr.score_forward(params, x, t, m)


# oracle trajectory, keeping the state before every update
ts = sd.timesteps(p, 0.03)
nz = list(noises)
x = sd.prior_sampling(p, mix_cpu, nz.pop(0))
worst_c = worst_p = 0.0
with model.cached_mixture(mix):
    for i in range(N):
        vt = torch.ones(1) * ts[i]
        vt_d = vt.to(DEV)
        zc, zp = nz.pop(0), nz.pop(0)
        # corrector from the oracle's state
        xc, _ = sd.corrector_step(p, score_cpu, x, vt, mix_cpu, [zc], 0.5)
        with sdes.injected_noise([zc]):
            g, _ = sde.corrector_update(x.to(DEV), model(x.to(DEV), vt_d, mix), vt_d, mix, 0.5)
        worst_c = max(worst_c, rel(g, xc))
        xp, xm = sd.predictor_step(p, score_cpu, xc, vt, mix_cpu, zp)
        with sdes.injected_noise([zp]):
            g, _ = sde.predictor_update(xc.to(DEV), model(xc.to(DEV), vt_d, mix), vt_d, mix, 1.0 / N)
        worst_p = max(worst_p, rel(g, xp))
        x = xp
want = xm
with sdes.injected_noise(noises):
    got, nfe = model.get_pc_sampler("reverse_diffusion", "ald2", mix, N=N, corrector_steps=1, snr=0.5, denoise=True)()
print(f"| nf=128, T=32000, N={N}, 1 corrector step: worst per-step error over {N} corrector updates (restarted from "
      f"the oracle state) | CPU oracle | {worst_c:.2e} | 1e-04 |")
print(f"| same, worst over {N} predictor updates | CPU oracle | {worst_p:.2e} | 1e-04 |")
print(f"| same, free-running final estimate after {nfe} evaluations (drift of two fp32 implementations) | CPU oracle | "
      f"{rel(got, want):.2e} | (not a tolerance) |")
