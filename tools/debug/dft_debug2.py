import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests" / "golden"))
os.environ["DSEP_CUDA_GRAPH"] = "0"
import torch
import cases
from diffsep_b200 import ops, synthetic as ow
from diffsep_b200.score_model import ScoreModelNCSNpp, LD

def rel(a, b): return float((a.double() - b.double()).norm() / max(float(b.double().norm()), 1e-30))
T = int(os.environ.get("T", "8000")); B = 1
sm = ScoreModelNCSNpp(num_sources=2, backbone_args=dict(nf=64), passes=3, state_dict=ow.make_score_model_state_dict(nf=64, seed=0))
xt, t, mix = (v.cuda() for v in cases.score_inputs(B, T, seed=3))
keys = ("frames_mix", "dft_mix", "frames", "dft", "x_pyr", "spec_out", "frames_out")
snaps = []
for tc in (True, False):
    sm._stft_tc = tc
    sm._bufs = {}
    y = sm(xt, t, mix).clone()
    torch.cuda.synchronize()
    bf = sm._work(B, T)
    snaps.append({k: bf[k].clone() for k in keys} | {"y": y})
M = B * 2 * snaps[0]["frames"].shape[0]
for k in keys + ("y",):
    a, b = snaps[0][k], snaps[1][k]
    n = min(a.shape[0], b.shape[0])
    print(k, tuple(a.shape), "rel", rel(a[:n], b[:n]), "finite", bool(torch.isfinite(a).all()), bool(torch.isfinite(b).all()))
Mx = 2 * bf["Fr"]
a, b = snaps[0]["frames_out"], snaps[1]["frames_out"]
err = ((a[:Mx].double() - b[:Mx].double()).norm(dim=1) / b[:Mx].double().norm(dim=1).clamp_min(1e-30))
print("frames_out per-row err: worst rows", torch.topk(err, 8).indices.tolist(), [f"{v:.2e}" for v in torch.topk(err, 8).values.tolist()])
a, b = snaps[0]["dft"], snaps[1]["dft"]
err = ((a[:Mx].double() - b[:Mx].double()).norm(dim=1) / b[:Mx].double().norm(dim=1).clamp_min(1e-30))
print("dft per-row err: worst rows", torch.topk(err, 8).indices.tolist(), [f"{v:.2e}" for v in torch.topk(err, 8).values.tolist()])
