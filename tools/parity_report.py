"""Measured parity of the CUDA path against the CPU oracle / reference goldens, level by level.
Prints a markdown table (committed as profiles/parity_rXX.md).  Test infrastructure: imports oracle/.

    python tools/parity_report.py > profiles/parity_r01.md
"""
import copy
import math
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests" / "golden"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import cases  # noqa: E402
from diffsep_b200 import ops, sdes  # noqa: E402
from diffsep_b200.pl_model import DEFAULT_CONFIG, DiffSepModel, normalize_batch  # noqa: E402
from diffsep_b200.score_model import ScoreModelNCSNpp  # noqa: E402
from oracle import ncsnpp_ref as nr, score_ref as sr, sde_ref as sd, weights as ow  # noqa: E402

DEV = "cuda"
rows = []


def rel(a, b):
    a, b = torch.as_tensor(a, dtype=torch.float64), torch.as_tensor(b, dtype=torch.float64)
    return float((a - b).norm() / b.norm())


def add(what, against, err, tol):
    rows.append((what, against, err, tol))


def score_model(nf, passes=3):
    return ScoreModelNCSNpp(num_sources=2, backbone_args=dict(nf=nf), passes=passes,
                            state_dict=ow.make_score_model_state_dict(nf=nf, seed=0))


G = ROOT / "tests" / "golden"
# backbone, layer by layer (nf=128, W=64)
for nf in (64, 128):
    sm = score_model(nf)
    params = ow.make_backbone_params(nf=nf, seed=0)
    g = cases.gen(31)
    x = torch.randn(2, 6, 256, 64, generator=g) * 0.5
    t = torch.tensor([0.8, 0.05])
    taps = {}
    with torch.no_grad():
        nr.ncsnpp_forward(params, x, t, taps=taps)
        taps64 = {}
        nr.ncsnpp_forward({k: v.double() for k, v in params.items()}, x.double(), t.double(), taps=taps64)
    xin = (2 * x - 1).permute(0, 2, 3, 1).contiguous().to(DEV)
    cpad = sm.backbone.conv_in.cin_pad
    xa = torch.zeros(2, 256, 64, cpad, device=DEV)
    xa[..., :6] = xin
    planes = ops.Split.empty((2, 256, 64, cpad), DEV)
    ops.split_f16(xa, planes)
    pyr = sm.backbone(planes, xin, t.to(DEV)).permute(0, 3, 1, 2).cpu()
    add(f"NCSN++ backbone nf={nf}, [2,6,256,64] -> output pyramid", "CPU oracle fp32", rel(pyr, taps["pyr0"]), 1e-4)
    add(f"NCSN++ backbone nf={nf} (same)", "CPU oracle fp64 (truth)", rel(pyr, taps64["pyr0"]), 1e-4)
    add(f"CPU oracle fp32 itself, nf={nf}", "CPU oracle fp64 (truth)", rel(taps["pyr0"], taps64["pyr0"]), None)

g = np.load(G / "score_nf128.npz")
xt, t, mix = cases.score_inputs(1, 7680, seed=9)
y = score_model(128)(xt.to(DEV), t.to(DEV), mix.to(DEV)).cpu()
add("ScoreModelNCSNpp.forward nf=128, T=7680", "golden from the real reference (CPU)", rel(y, g["y"]), 1e-4)
y1 = score_model(128, passes=1)(xt.to(DEV), t.to(DEV), mix.to(DEV)).cpu()
add("same, 1-pass mode (11-bit operands)", "golden from the real reference (CPU)", rel(y1, g["y"]), 1e-2)

g = np.load(G / "stft.npz")
g = np.load(G / "sampler.npz")
(mixn, _), _, _ = normalize_batch((cases.batch_mix(2, 1024).to(DEV), None))
for tag, cls in (("mix", sdes.MixSDE), ("priormix", sdes.PriorMixSDE)):
    sde = cls(ndim=2, d_lambda=2.0, sigma_min=0.05, sigma_max=0.5, N=30)
    with sdes.injected_noise(cases.sampler_noises(2, 1024, 30, 1)):
        out, _ = sdes.get_pc_sampler("reverse_diffusion", "ald2", sde=sde, score_fn=cases.analytic_score, y=mixn,
                                     eps=0.03, snr=0.5, corrector_steps=1, denoise=True)()
    add(f"PC sampler, {cls.__name__}, N=30, 1 corrector step, analytic score", "golden from the real reference sampler",
        rel(out.cpu(), g[f"{tag}.cs1"]), 1e-5)

cfg = copy.deepcopy(DEFAULT_CONFIG)
cfg["model"]["score_model"]["backbone_args"]["nf"] = 64
model = DiffSepModel(cfg, score_state_dict=ow.make_score_model_state_dict(nf=64, seed=0))
params = ow.make_backbone_params(nf=64, seed=0)
mix_cpu, _, _ = sd.normalize_batch(cases.batch_mix(1, 4096))
noises = cases.sampler_noises(1, 4096, 5, 1)


def score_fn(x, t, m):
    with torch.no_grad():
        return sr.score_forward(params, x, t, m)


want, _, im_w = sd.pc_sampler(sd.MixSDEParams(N=5), score_fn, mix_cpu, noises, eps=0.03, snr=0.5, corrector_steps=1,
                              denoise=True, intermediate=True)
(mix, _), _, _ = normalize_batch((cases.batch_mix(1, 4096).to(DEV), None))
with sdes.injected_noise(noises):
    got, nfe, im = model.get_pc_sampler("reverse_diffusion", "ald2", mix, N=5, corrector_steps=1, snr=0.5,
                                        denoise=True, intermediate=True)()
for i, ((gx, _), (wx, _)) in enumerate(zip(im, im_w)):
    add(f"network-driven PC sampler nf=64, N=5, default mode, FREE-RUNNING state after corrector of step {i + 1}/5",
        "CPU oracle sampler", rel(gx.cpu(), wx), 4e-4)
add("network-driven PC sampler nf=64, N=5, default mode, free-running final estimate (10 evaluations)",
    "CPU oracle sampler", rel(got.cpu(), want), 4e-4)

print("| what | against | rel-L2 error | tolerance in tests |\n|---|---|---:|---:|")
for what, against, err, tol in rows:
    print(f"| {what} | {against} | {err:.2e} | {'' if tol is None else f'{tol:.0e}'} |")
