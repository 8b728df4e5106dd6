"""CPU: pins the oracle (oracle/) against golden vectors produced by the real reference
(tests/golden/make_golden.py).  Tolerances are fp32 round-off of re-ordered arithmetic."""
import json

import numpy as np
import pytest
import torch

import cases
from conftest import GOLDEN, rel_l2
from oracle import ncsnpp_ref as nr
from oracle import score_ref as sr
from oracle import sde_ref as sd
from oracle import weights as ow


@pytest.mark.parametrize("nf", [128, 64])
def test_structure_matches_reference(nf):
    ref = json.loads((GOLDEN / f"structure_nf{nf}.json").read_text())
    mine = [[k, list(v)] for k, v in ow.backbone_param_shapes(nf).items()]
    assert mine == ref["parameters"]           # names, shapes AND parameters() order (EMA order)
    sd_names = [k for k, _ in ref["state_dict"]]
    assert sd_names == ["backbone." + k for k, _ in mine] + ["stft.window", "stft_inv.window"]
    assert sum(int(np.prod(s)) for _, s in mine) == ref["n_params"]


def test_param_count_nf128():
    assert sum(int(np.prod(s)) for s in ow.backbone_param_shapes(128).values()) == 65_623_366


def test_fir(golden):
    g = golden("fir.npz")
    x = torch.from_numpy(g["x"])
    assert np.abs(nr.fir_down2(x).numpy() - g["down"]).max() < 1e-6
    assert np.abs(nr.fir_up2(x).numpy() - g["up"]).max() < 1e-6


@pytest.mark.parametrize("case", cases.RESBLOCK_CASES, ids=lambda c: c[0])
def test_resblock(golden, case):
    name, kw, _ = case
    g = golden("blocks.npz")
    params = {"all_modules.0." + k[len(name) + 3:]: torch.from_numpy(g[k]) for k in g.files
              if k.startswith(name + ".p.")}
    x, temb = torch.from_numpy(g[name + ".x"]), torch.from_numpy(g[name + ".temb"])
    y = nr.resblock(nr._P(params, 0), x, nr.silu(temb), up=kw["up"], down=kw["down"])
    assert rel_l2(y, g[name + ".y"]) < 2e-6


def test_attnblock(golden):
    g = golden("blocks.npz")
    params = {"all_modules.0." + k[len("attn.p."):]: torch.from_numpy(g[k]) for k in g.files
              if k.startswith("attn.p.")}
    y = nr.attnblock(nr._P(params, 0), torch.from_numpy(g["attn.x"]))
    assert rel_l2(y, g["attn.y"]) < 2e-6


def test_stft_wrapper(golden):
    g = golden("stft.npz")
    xt, t, mix = cases.score_inputs(2, 2048, seed=7)
    spec, n_samples, n_pad = sr.pre_process(torch.cat((xt, mix), dim=1))
    assert n_pad == int(g["n_pad"]) and n_samples == 2048
    assert spec.shape[-1] % 64 == 0
    assert rel_l2(spec[..., : spec.shape[-1] - n_pad], g["spec"]) < 1e-6
    assert float(spec[..., spec.shape[-1] - n_pad:].abs().max()) == 0.0
    net_out = torch.randn(2, 4, 256, spec.shape[-1], generator=cases.gen(8)) * 0.2
    wav = sr.post_process(net_out, n_samples, n_pad)
    assert rel_l2(wav, g["wav"]) < 1e-6


def test_stft_matrix_formulation_pins_frame_indexing(golden):
    """The explicit float64 DFT-matrix formulation reproduces torch.stft / torch.istft."""
    g = golden("stft.npz")
    xt, t, mix = cases.score_inputs(2, 2048, seed=7)
    x = torch.cat((xt, mix), dim=1)
    S = sr.stft_matrix(x)
    assert S.shape[-1] == sr.n_frames(2048) == 19
    mag = S.abs()
    comp = (mag ** 0.5) * torch.exp(1j * S.angle()) * 0.15
    xr = torch.stack((comp.real, comp.imag), dim=1).flatten(1, 2)
    assert rel_l2(xr, g["spec"]) < 1e-6
    # inverse
    spec = torch.randn(2, 2, 256, 19, dtype=torch.complex128, generator=cases.gen(3))
    win = torch.hann_window(510, dtype=torch.float64)
    ref = torch.istft(spec.reshape(4, 256, 19), n_fft=510, hop_length=128, window=win, center=True)
    mine = sr.istft_matrix(spec).reshape(4, -1)
    assert mine.shape == ref.shape == (4, 128 * 18)
    assert rel_l2(mine, ref) < 1e-12


def test_score_model_nf32(golden):
    g = golden("score_nf32.npz")
    params = ow.make_backbone_params(nf=32, seed=0)
    xt, t, mix = cases.score_inputs(2, 2048, seed=7)
    with torch.no_grad():
        y = sr.score_forward(params, xt, t, mix)
    assert rel_l2(y, g["y"]) < 1e-5


@pytest.mark.parametrize("tag", ["mix", "priormix"])
@pytest.mark.parametrize("cs", [0, 1, 2])
def test_sampler_analytic(golden, tag, cs):
    g = golden("sampler.npz")
    mix = cases.batch_mix(2, 1024)
    mix, _, _ = sd.normalize_batch(mix)
    p = sd.MixSDEParams(N=30, prior=(tag == "priormix"))
    out, nfe = sd.pc_sampler(p, cases.analytic_score, mix, cases.sampler_noises(2, 1024, 30, cs),
                             eps=0.03, snr=0.5, corrector_steps=cs, denoise=True)
    assert nfe == 30 * (cs + 1)
    assert rel_l2(out, g[f"{tag}.cs{cs}"]) < 5e-6


@pytest.mark.parametrize("sched", ["linear", "log", "revlog"])
def test_sampler_schedules(golden, sched):
    g = golden("sampler.npz")
    mix, _, _ = sd.normalize_batch(cases.batch_mix(2, 1024))
    p = sd.MixSDEParams(N=10)
    out, _ = sd.pc_sampler(p, cases.analytic_score, mix, cases.sampler_noises(2, 1024, 10, 1),
                           eps=0.03, snr=0.5, corrector_steps=1, denoise=False, schedule=sched)
    assert rel_l2(out, g[f"sched.{sched}"]) < 5e-6


def test_sampler_network_nf32(golden):
    g = golden("sampler.npz")
    params = ow.make_backbone_params(nf=32, seed=0)
    mix, _, _ = sd.normalize_batch(cases.batch_mix(1, 2048))
    p = sd.MixSDEParams(N=3)

    def score_fn(x, t, m):
        with torch.no_grad():
            return sr.score_forward(params, x, t, m)
    out, nfe, im = sd.pc_sampler(p, score_fn, mix, cases.sampler_noises(1, 2048, 3, 1), eps=0.03,
                                 snr=0.5, corrector_steps=1, denoise=True, intermediate=True)
    assert nfe == 6
    assert rel_l2(im[0][0], g["net32.im0"]) < 1e-5
    assert rel_l2(out, g["net32.out"]) < 1e-4


def test_misc(golden):
    g = golden("misc.npz")
    m = cases.batch_mix(3, 512) * 3.0 + 0.2
    norm, mean, std = sd.normalize_batch(m)
    assert rel_l2(norm, g["norm"]) < 1e-6
    sep = torch.randn(3, 2, 512, generator=cases.gen(5))
    assert rel_l2(sd.scale_output(m, sep), g["scaled"]) < 1e-6


def test_flop_count_matches_survey():
    assert abs(nr.count_flops(128, 256)["total"] / 1e9 - 532.891) < 0.01
    assert abs(nr.count_flops(128, 512)["total"] / 1e9 - 1066.173) < 0.01
    assert abs(nr.count_flops(64, 256)["total"] / 1e9 - 133.832) < 0.01


def test_score_model_nf128(golden):
    g = golden("score_nf128.npz")
    params = ow.make_backbone_params(nf=128, seed=0)
    xt, t, mix = cases.score_inputs(1, 7680, seed=9)
    with torch.no_grad():
        y = sr.score_forward(params, xt, t, mix)
    assert rel_l2(y, g["y"]) < 1e-5


PLUGIN_CASES = [("em_ald2", "euler_maruyama", "ald2", False, 1), ("rd_ald", "reverse_diffusion", "ald", False, 2),
                ("rd_langevin", "reverse_diffusion", "langevin", False, 1),
                ("rd_langevin_prior", "reverse_diffusion", "langevin", True, 1),
                ("rd_ald2_pflow", "reverse_diffusion", "ald2", False, 1), ("em_ald2_pflow", "euler_maruyama", "ald2", True, 1),
                ("none_ald2", "none", "ald2", False, 1)]


@pytest.mark.parametrize("case", PLUGIN_CASES, ids=lambda c: c[0])
def test_sampler_other_plugins(golden, case):
    """ald / langevin correctors, euler_maruyama / none predictors, and the reference's no-op
    probability_flow flag, vs the real reference sampler (tests/golden/make_golden_plugins.py)."""
    name, pred, corr, prior, cs = case
    g = golden("plugins.npz")
    mix, _, _ = sd.normalize_batch(cases.batch_mix(2, 1024))
    p = sd.MixSDEParams(N=10, prior=prior)
    out = sd.pc_sampler_plugins(p, cases.analytic_score, mix, cases.sampler_noises(2, 1024, 10, cs), predictor=pred,
                                corrector=corr, corrector_steps=cs, denoise=False)
    assert rel_l2(out, g[name]) < 5e-6


@pytest.mark.parametrize("case", cases.NDIM_CASES, ids=lambda c: c[0])
def test_three_sources_and_true_mean_prior(golden, case):
    """PriorMixSDE with ndim = 3 (the 3-speaker models) and the sampler's ``true_mean`` prior
    (an ndim-channel input: mean = the input, sigma_mix per input channel; MixSDE keeps its 0.5 and its two
    hard-coded channels) vs the real reference sampler (tests/golden/make_golden_ndim.py)."""
    name, sde_name, ndim, tm_ch, cs = case
    g = golden("ndim.npz")
    mix, _, _ = sd.normalize_batch(cases.batch_mix(cases.NDIM_B, cases.NDIM_T))
    p = sd.MixSDEParams(ndim=ndim, N=cases.NDIM_N, prior=sde_name == "priormix")
    tm = cases.ndim_true_mean(tm_ch) if tm_ch else None
    out, nfe = sd.pc_sampler(p, cases.analytic_score, mix, cases.ndim_noises(ndim, cs), corrector_steps=cs,
                             denoise=True, true_mean=tm)
    assert nfe == cases.NDIM_N * (cs + 1) and out.shape == (cases.NDIM_B, ndim, cases.NDIM_T)
    assert rel_l2(out, g[name]) < 5e-6


@pytest.mark.parametrize("name,prior,ndim", [("mix", False, 2), ("priormix", True, 2), ("priormix3", True, 3)])
def test_training_forward_pieces_match_reference_golden(name, prior, ndim):
    """oracle sample_prior / score_loss (pl_model.py:179-247, 411-424 on sde.marginal_prob / mult_std) vs outputs of the
    REAL reference (tests/golden/make_golden_training.py)."""
    g = np.load(GOLDEN / "training.npz")
    p = sd.MixSDEParams(ndim=ndim, prior=prior)
    t = lambda k: torch.from_numpy(g[f"{name}_{k}"])
    target = t("target")
    mix = target.sum(dim=1, keepdim=True)
    assert rel_l2(sd.marginal_mean(p, target, t("time")), t("mean")) < 1e-6
    assert rel_l2(sd.sample_prior(p, mix, target, t("time"), t("z")), t("xt")) < 1e-6
    assert rel_l2(sd.score_loss(p, t("score"), t("z"), t("time"), mix, "none"), t("loss_none")) < 1e-6
    assert abs(float(sd.score_loss(p, t("score"), t("z"), t("time"), mix)) - float(t("loss_mean"))) < 1e-6
