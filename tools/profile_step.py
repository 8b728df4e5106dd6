"""One predictor-corrector step of configs[1] (batch 32 x 4 s @ 8 kHz, nf=128) for ncu: prior sample, ald2 corrector,
reverse-diffusion predictor = 2 score-network evaluations + the three fused SDE kernels + normalize / scale_output.

    ncu --profile-from-start off --metrics ... python tools/profile_step.py
"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests" / "golden"))
import torch  # noqa: E402

import cases  # noqa: E402
from diffsep_b200 import ops, synthetic  # noqa: E402
from diffsep_b200.pl_model import DEFAULT_CONFIG, DiffSepModel  # noqa: E402

B, T, N = 32, 32000, 30
model = DiffSepModel(DEFAULT_CONFIG, score_state_dict=synthetic.make_score_model_state_dict(nf=128, seed=0))
mix_raw = cases.batch_mix(B, T).cuda()
sde = model.sde.copy()
sde.N = N
vt = torch.ones(B, device="cuda")


def step():
    (mix, _), _, _ = model.normalize_batch((mix_raw, None))
    with model.cached_mixture(mix):
        x0 = sde.prior_sampling(mix.shape, mix)
        xc, _ = sde.corrector_update(x0, model(x0, vt, mix), vt, mix, 0.5)
        xp, xm = sde.predictor_update(xc, model(xc, vt, mix), vt, mix, 1.0 / N)
    out = torch.empty_like(xm)
    ops.scale_output(mix_raw, xm, B, 2, T, out)
    return out


step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("one PC step done", flush=True)
