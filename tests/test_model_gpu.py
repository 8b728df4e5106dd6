"""GPU parity of the assembled path — STFT wrapper, NCSN++ backbone, score model, PC sampler —
through the public Python surface (which calls the C-ABI) against the CPU oracle and the golden
vectors generated from the real reference (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

import cases
from conftest import rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _score_model(nf, passes=3, seed=0):
    from diffsep_b200.score_model import ScoreModelNCSNpp
    from oracle import weights as ow
    sd = ow.make_score_model_state_dict(nf=nf, seed=seed)
    return ScoreModelNCSNpp(num_sources=2, backbone_args=dict(nf=nf), state_dict=sd, passes=passes)


def test_stft_frontend_matches_reference_golden(golden):
    """frames -> DFT-510 GEMM -> compress/pack vs the reference's pre_process golden (torch.stft):
    frame count and padding bit-exact, values < 2e-6 rel-L2 of spec (fp32 DFT round-off)."""
    from diffsep_b200 import ops
    from diffsep_b200.score_model import ScoreModelNCSNpp, n_frames, LD
    g = golden("stft.npz")
    xt, t, mix = cases.score_inputs(2, 2048, seed=7)
    x = torch.cat((xt, mix), dim=1).to(DEV)
    B, C, T = x.shape
    Fr = n_frames(T)
    assert Fr == g["spec"].shape[-1] == 19
    sm = ScoreModelNCSNpp(num_sources=2, backbone_args=dict(nf=64))
    frames = torch.empty(B * C * Fr, LD, device=DEV)
    dft = torch.empty(B * C * Fr, LD, device=DEV)
    ops.stft_frames(x, sm.window, B, C, T, Fr, frames)
    ops.sgemm(frames, LD, sm.basis_fwd, LD, dft, LD, B * C * Fr, LD, LD)
    Wp = 64
    x_f32 = torch.zeros(B, 256, Wp, 6, device=DEV)
    planes = ops.Split.zeros((B, 256, Wp, 64), DEV)
    ops.spec_pack(dft, B, C, Fr, Wp, 0, C, 64, 0.15, 0.5, x_f32, planes)
    torch.cuda.synchronize()
    got = x_f32.permute(0, 3, 1, 2).cpu()          # [B, 6, 256, Wp], after the backbone's 2x-1
    spec = (got + 1.0) / 2.0
    assert rel_l2(spec[..., :Fr], g["spec"]) < 2e-6
    assert float((got[..., Fr:] + 1.0).abs().max()) == 0.0      # zero frames -> exactly -1
    rec = (planes.hi.float() + planes.lo.float())[..., :6].permute(0, 3, 1, 2).cpu()
    assert rel_l2(rec, got) < 1e-6
    assert float(planes.hi[..., 6:].abs().max()) == 0.0


def test_istft_backend_matches_reference_golden(golden):
    """out_head (1x1 conv folded to identity here) -> inverse DFT GEMM -> OLA vs the reference's
    post_process golden (torch.istft): < 2e-6."""
    from diffsep_b200 import ops
    from diffsep_b200.score_model import ScoreModelNCSNpp, LD
    g = golden("stft.npz")
    B, ns, T, Fr, Wp = 2, 2, 2048, 19, 64
    net_out = torch.randn(2, 4, 256, Wp, generator=cases.gen(8)) * 0.2      # [re0, re1, im0, im1]
    sm = ScoreModelNCSNpp(num_sources=2, backbone_args=dict(nf=64))
    # feed net_out through out_head as a 4-channel "pyramid" with identity 1x1 conv and t = 1
    pyr = net_out.permute(0, 2, 3, 1).contiguous().to(DEV)
    w = torch.eye(4, device=DEV)
    spec = torch.zeros(B * ns * Fr, LD, device=DEV)
    ops.out_head(pyr, B, Wp, 4, ns, Fr, torch.ones(B, device=DEV), w, None, 0.15, 0.5, spec)
    frames_t = torch.empty(B * ns * Fr, LD, device=DEV)
    ops.sgemm(spec, LD, sm.basis_inv, LD, frames_t, LD, B * ns * Fr, LD, LD)
    wav = torch.empty(B, ns, T, device=DEV)
    ops.istft_ola(frames_t, sm.window, B, ns, Fr, T, wav)
    torch.cuda.synchronize()
    assert rel_l2(wav.cpu(), g["wav"]) < 2e-6


def test_stft_istft_round_trip_full_length():
    """Size-independent property at the benchmark length: iSTFT(STFT(x)) == x on [0, 128 (Fr-1))."""
    from diffsep_b200 import ops
    from diffsep_b200.score_model import ScoreModelNCSNpp, n_frames, LD
    B, C, T = 2, 2, 32000
    x = torch.randn(B, C, T, generator=cases.gen(21)).to(DEV)
    Fr = n_frames(T)
    assert Fr == 253
    sm = ScoreModelNCSNpp(num_sources=2, backbone_args=dict(nf=64))
    M = B * C * Fr
    frames, dft, back = (torch.empty(M, LD, device=DEV) for _ in range(3))
    ops.stft_frames(x, sm.window, B, C, T, Fr, frames)
    ops.sgemm(frames, LD, sm.basis_fwd, LD, dft, LD, M, LD, LD)
    ops.sgemm(dft, LD, sm.basis_inv, LD, back, LD, M, LD, LD)
    y = torch.empty(B, C, T, device=DEV)
    ops.istft_ola(back, sm.window, B, C, Fr, T, y)
    torch.cuda.synchronize()
    assert rel_l2(y.cpu(), x.cpu()) < 2e-6


@pytest.mark.parametrize("nf", [64, 128])
def test_backbone_layerwise_vs_oracle(nf):
    """Whole NCSN++ forward at W=64 against the CPU fp32 oracle, same seeded weights and inputs.
    Stated tolerance: 1e-4 rel-L2 on the output pyramid (measured ~1e-6)."""
    from diffsep_b200 import ops
    from oracle import ncsnpp_ref as nr, weights as ow
    sm = _score_model(nf)
    params = ow.make_backbone_params(nf=nf, seed=0)
    g = cases.gen(31)
    B, W = 2, 64
    x = torch.randn(B, 6, 256, W, generator=g) * 0.5
    t = torch.tensor([0.8, 0.05])
    taps = {}
    with torch.no_grad():
        nr.ncsnpp_forward(params, x, t, taps=taps)
    xin = (2 * x - 1).permute(0, 2, 3, 1).contiguous().to(DEV)
    cpad = sm.backbone.conv_in.cin_pad
    xa = torch.zeros(B, 256, W, cpad, device=DEV)
    xa[..., :6] = xin
    planes = ops.Split.empty((B, 256, W, cpad), DEV)
    ops.split_f16(xa, planes)
    pyr = sm.backbone(planes, xin, t.to(DEV))
    torch.cuda.synchronize()
    got = pyr.permute(0, 3, 1, 2).cpu()
    assert rel_l2(got, taps["pyr0"]) < 1e-4


def test_score_model_nf128_matches_reference_golden(golden):
    """ScoreModelNCSNpp.forward (nf=128, the benchmark architecture) vs the golden produced by the
    real reference on CPU: north-star tolerance 1e-4 rel-L2."""
    g = golden("score_nf128.npz")
    sm = _score_model(128)
    xt, t, mix = cases.score_inputs(1, 7680, seed=9)
    y = sm(xt.to(DEV), t.to(DEV), mix.to(DEV))
    torch.cuda.synchronize()
    assert y.shape == (1, 2, 7680)
    assert rel_l2(y.cpu(), g["y"]) < 1e-4


def test_score_model_nf32_matches_reference_golden(golden):
    """nf=32 (channel counts 32/64/96: partial output tiles, 4-channel groups) vs the real reference."""
    from diffsep_b200.backbone import cin_align
    if cin_align() > 32:
        pytest.skip("needs the 32-channel K-block build (DSEP_CONV_BK=32)")
    g = golden("score_nf32.npz")
    sm = _score_model(32)
    xt, t, mix = cases.score_inputs(2, 2048, seed=7)
    y = sm(xt.to(DEV), t.to(DEV), mix.to(DEV))
    torch.cuda.synchronize()
    assert rel_l2(y.cpu(), g["y"]) < 1e-4


def test_score_model_non_power_of_two_width_vs_oracle():
    """T = 24000 -> 190 frames -> W = 192: level widths 192 ... 3 (odd), edge tiles everywhere."""
    from oracle import score_ref as sr, weights as ow
    params = ow.make_backbone_params(nf=64, seed=0)
    xt, t, mix = cases.score_inputs(1, 24000, seed=11)
    with torch.no_grad():
        want = sr.score_forward(params, xt, t, mix)
    y = _score_model(64)(xt.to(DEV), t.to(DEV), mix.to(DEV))
    torch.cuda.synchronize()
    assert rel_l2(y.cpu(), want) < 1e-4


def test_score_model_16khz_shape_vs_oracle():
    """configs[3] shape: 4 s @ 16 kHz -> 503 frames -> W = 512 (nf=64 to keep the CPU oracle short)."""
    from oracle import score_ref as sr, weights as ow
    params = ow.make_backbone_params(nf=64, seed=0)
    xt, t, mix = cases.score_inputs(1, 64000, seed=13)
    with torch.no_grad():
        want = sr.score_forward(params, xt, t, mix)
    y = _score_model(64)(xt.to(DEV), t.to(DEV), mix.to(DEV))
    torch.cuda.synchronize()
    assert rel_l2(y.cpu(), want) < 1e-4


def test_long_form_30s_properties():
    """configs[4] shape: 30 s @ 8 kHz -> 1878 frames -> W = 1920, attention over 1920 tokens.  The CPU
    oracle takes minutes here, so the full size is checked through properties: finite output,
    batch-entry independence (entry 0 alone == entry 0 in a batch of 2) and STFT framing."""
    from diffsep_b200.score_model import n_frames
    assert n_frames(240000) == 1878
    sm = _score_model(64)
    xt, t, mix = (v.to(DEV) for v in cases.score_inputs(2, 240000, seed=17))
    y2 = sm(xt, t, mix)
    y1 = sm(xt[:1].contiguous(), t[:1].contiguous(), mix[:1].contiguous())
    torch.cuda.synchronize()
    assert y2.shape == (2, 2, 240000) and bool(torch.isfinite(y2).all())
    assert rel_l2(y1.cpu(), y2[:1].cpu()) < 1e-5


def _network_sampler_vs_oracle(sde_cfg, prior, T, N):
    """A network-driven PC sampler (nf=64, injected noise) against the CPU oracle sampler.

    (a) the north-star criterion, per step: every corrector / predictor update re-started from the ORACLE's state
        stays within 1e-4 in the default mode (passes = 2);
    (b) free-running (each implementation continues from its own state): the three-product mode stays within 1e-4
        over the whole run, which pins the loop itself (time grid, noise order, denoise); the default mode's
        per-step error (~3e-5) compounds over the 2N evaluations, bounded here at 4e-4 — drift, not a tolerance."""
    import copy
    from diffsep_b200 import sdes
    from diffsep_b200.pl_model import DEFAULT_CONFIG, DiffSepModel, normalize_batch
    from oracle import score_ref as sr, sde_ref as sd, weights as ow
    cfg = copy.deepcopy(DEFAULT_CONFIG)
    cfg["model"]["score_model"]["backbone_args"]["nf"] = 64
    if sde_cfg is not None:
        cfg["model"]["sde"] = sde_cfg
    params = ow.make_backbone_params(nf=64, seed=0)
    mix_cpu, _, _ = sd.normalize_batch(cases.batch_mix(1, T))
    noises = cases.sampler_noises(1, T, N, 1)

    def score_fn(x, t, m):
        with torch.no_grad():
            return sr.score_forward(params, x, t, m)
    p = sd.MixSDEParams(N=N, prior=prior)
    want, nfe_w, im_w = sd.pc_sampler(p, score_fn, mix_cpu, noises, eps=0.03, snr=0.5, corrector_steps=1,
                                      denoise=True, intermediate=True)
    (mix, _), _, _ = normalize_batch((cases.batch_mix(1, T).to(DEV), None))
    assert rel_l2(mix.cpu(), mix_cpu) < 1e-6
    for passes, tol in ((3, 1e-4), (None, 4e-4)):
        model = DiffSepModel(cfg, passes=passes, score_state_dict=ow.make_score_model_state_dict(nf=64, seed=0))
        assert isinstance(model.sde, sdes.PriorMixSDE if prior else sdes.MixSDE)
        with sdes.injected_noise(noises):
            got, nfe, im = model.get_pc_sampler("reverse_diffusion", "ald2", mix, N=N, corrector_steps=1, snr=0.5,
                                                denoise=True, intermediate=True)()
        torch.cuda.synchronize()
        assert nfe == nfe_w == 2 * N
        for (gx, gm), (wx, wm) in zip(im, im_w):
            assert rel_l2(gx.cpu(), wx) < tol
        assert rel_l2(got.cpu(), want) < tol
    # (a) per step in the default mode, restarted from the oracle's state (model = the passes=None one)
    sde = model.sde
    ts = sd.timesteps(p, 0.03)
    nz = list(noises)
    x = sd.prior_sampling(p, mix_cpu, nz.pop(0))
    with model.score_model.cached_mixture(mix):
        for i in range(N):
            vt = torch.ones(1) * ts[i]
            vt_d = vt.to(DEV)
            zc, zp = nz.pop(0), nz.pop(0)
            xc, _ = sd.corrector_step(p, score_fn, x, vt, mix_cpu, [zc], 0.5)
            with sdes.injected_noise([zc]):
                g, _ = sde.corrector_update(x.to(DEV), model(x.to(DEV), vt_d, mix), vt_d, mix, 0.5)
            assert rel_l2(g.cpu(), xc) < 1e-4, ("corrector", i)
            xp, _ = sd.predictor_step(p, score_fn, xc, vt, mix_cpu, zp)
            with sdes.injected_noise([zp]):
                g, _ = sde.predictor_update(xc.to(DEV), model(xc.to(DEV), vt_d, mix), vt_d, mix, 1.0 / N)
            assert rel_l2(g.cpu(), xp) < 1e-4, ("predictor", i)
            x = xp


def test_enhancement_sampler_priormix_with_network_vs_oracle():
    """configs[3] path: PriorMixSDE (sigma_mix-scaled prior / corrector / predictor) driving the network, N=2."""
    _network_sampler_vs_oracle({"_target_": "sdes.sdes.PriorMixSDE", "ndim": 2, "d_lambda": 2.0, "sigma_min": 0.05,
                                "sigma_max": 0.5, "N": 30, "avg_len": 510}, True, 8000, 2)


def test_separate_cli_end_to_end(tmp_path):
    """separate.py input_dir output_dir --model <checkpoint> -N 2: reference CLI surface (flags, s0/ s1/
    layout) on a Lightning-style checkpoint with EMA weights; output equals the API path."""
    import subprocess, sys
    from pathlib import Path
    import numpy as np
    from scipy.io import wavfile
    from oracle import weights as ow
    from diffsep_b200.pl_model import DEFAULT_CONFIG
    import copy
    root = Path(__file__).resolve().parent.parent
    cfg = copy.deepcopy(DEFAULT_CONFIG)
    cfg["model"]["score_model"]["backbone_args"]["nf"] = 64
    sd_ = ow.make_score_model_state_dict(nf=64, seed=0)
    # state_dict holds *different* (seed 1) raw weights; the EMA shadow holds the seed-0 ones that must win
    raw = ow.make_score_model_state_dict(nf=64, seed=1)
    names = [k for k in sd_ if k.startswith("backbone.") and not k.endswith("all_modules.0.W")]
    raw["backbone.all_modules.0.W"] = sd_["backbone.all_modules.0.W"]
    ckpt = {"state_dict": {"score_model." + k: v for k, v in raw.items()},
            "hyper_parameters": {"config": cfg},
            "ema": {"shadow_params": [sd_[k] for k in names], "decay": 0.999}}
    torch.save(ckpt, tmp_path / "checkpoint.pt")
    (tmp_path / "in").mkdir()
    wav = (cases.batch_mix(1, 8000)[0, 0] * 0.5).numpy().astype(np.float32)
    wavfile.write(tmp_path / "in" / "utt.wav", 8000, wav)
    env = dict(**__import__("os").environ)
    r = subprocess.run([sys.executable, str(root / "separate.py"), str(tmp_path / "in"), str(tmp_path / "out"),
                        "--model", str(tmp_path / "checkpoint.pt"), "-N", "2", "--corrector-steps", "1"],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    outs = []
    for i in range(2):
        sr_, data = wavfile.read(tmp_path / "out" / f"s{i}" / "utt.wav")
        assert sr_ == 8000 and data.shape == (8000,)
        outs.append(torch.from_numpy(data.astype(np.float32)))
    assert all(bool(torch.isfinite(o).all()) and float(o.abs().max()) > 0 for o in outs)
    # same weights through the API (seed-0 == EMA shadow) with the CLI's RNG seed give the same audio
    from diffsep_b200.pl_model import DiffSepModel
    m = DiffSepModel.load_from_checkpoint(str(tmp_path / "checkpoint.pt"))
    w_api = m.score_model.backbone.conv_in.planes.hi
    from diffsep_b200.score_model import ScoreModelNCSNpp
    w_ref = ScoreModelNCSNpp(num_sources=2, backbone_args=dict(nf=64), state_dict=sd_).backbone.conv_in.planes.hi
    assert torch.equal(w_api, w_ref)


def test_score_model_tf32_grade_mode():
    """passes=1 (11-bit operands, what cuDNN's default TF32 gives the reference on a GPU) stays
    within 1e-2 of the fp32 oracle; reported, not the parity mode."""
    from oracle import score_ref as sr, weights as ow
    params = ow.make_backbone_params(nf=64, seed=0)
    xt, t, mix = cases.score_inputs(1, 4096, seed=5)
    with torch.no_grad():
        want = sr.score_forward(params, xt, t, mix)
    y = _score_model(64, passes=1)(xt.to(DEV), t.to(DEV), mix.to(DEV))
    torch.cuda.synchronize()
    assert rel_l2(y.cpu(), want) < 1e-2


def test_score_model_e4m3_correction_mode_matches_reference_golden(golden):
    """passes=2 (the product default): fp16 hi*hi plus one e4m3 product for both correction terms in every conv on
    a map of at least 16 x 8 — two tensor-core units per MAC — meets the north-star tolerance against the real
    reference's golden (tools/numerics_study.py predicts ~5e-5)."""
    g = golden("score_nf128.npz")
    sm = _score_model(128, passes=2)
    xt, t, mix = cases.score_inputs(1, 7680, seed=9)
    y = sm(xt.to(DEV), t.to(DEV), mix.to(DEV))
    torch.cuda.synchronize()
    assert rel_l2(y.cpu(), g["y"]) < 1e-4


def test_hoisted_time_embedding_equals_per_evaluation_one():
    """prepare_times / uniform_time (the FiLM rows of the sampler's time grid evaluated once, row 0 serving the whole
    batch) reproduce the per-evaluation time-embedding + Dense_0 kernels, with and without the CUDA graph."""
    sm = _score_model(64)
    xt, _, mix = (v.to(DEV) for v in cases.score_inputs(3, 2048, seed=4))
    for tval in (1.0, 0.5166666507720947, 0.03):
        t = torch.full((3,), tval, device=DEV)
        ref = sm(xt, t, mix)
        sm.prepare_times([tval])
        with sm.uniform_time(tval):
            got = sm(xt, t, mix)
            with sm.cached_mixture(mix):
                sm(xt, t, mix)                      # first call inside the context computes the mixture spectrogram
                got_graph = sm(xt, t, mix)          # second one replays the graph
        torch.cuda.synchronize()
        # (not torch.equal: the GroupNorm sums are combined with fp64 atomics in an order that varies run to run)
        assert rel_l2(got.cpu(), ref.cpu()) < 1e-6 and rel_l2(got_graph.cpu(), ref.cpu()) < 1e-6, tval


def test_score_model_argument_errors():
    sm = _score_model(64)
    xt, t, mix = (v.to(DEV) for v in cases.score_inputs(2, 1024))
    with pytest.raises(ValueError):
        sm(xt[:, :1], t, mix)
    with pytest.raises(ValueError):
        sm(xt, t[:1], mix)
    with pytest.raises(ValueError):
        sm(xt.cpu(), t, mix)


@pytest.mark.parametrize("tag", ["mix", "priormix"])
@pytest.mark.parametrize("cs", [0, 1, 2])
def test_sampler_analytic_matches_reference_golden(golden, tag, cs):
    """Full PC sampler (registries, time grid, ald2 + reverse_diffusion, injected noise) with the
    closed-form score of SURVEY.md §4-4 vs the reference's own sampler output: < 1e-5."""
    from diffsep_b200 import sdes
    from diffsep_b200.pl_model import normalize_batch
    g = golden("sampler.npz")
    mix = cases.batch_mix(2, 1024).to(DEV)
    (mix, _), _, _ = normalize_batch((mix, None))
    cls = sdes.MixSDE if tag == "mix" else sdes.PriorMixSDE
    sde = cls(ndim=2, d_lambda=2.0, sigma_min=0.05, sigma_max=0.5, N=30)
    noises = cases.sampler_noises(2, 1024, 30, cs)
    with sdes.injected_noise(noises):
        out, nfe = sdes.get_pc_sampler("reverse_diffusion", "ald2", sde=sde, score_fn=cases.analytic_score,
                                       y=mix, eps=0.03, snr=0.5, corrector_steps=cs, denoise=True)()
    torch.cuda.synchronize()
    assert nfe == 30 * (cs + 1)
    assert rel_l2(out.cpu(), g[f"{tag}.cs{cs}"]) < 1e-5


@pytest.mark.parametrize("sched", ["linear", "log", "revlog"])
def test_sampler_schedules_match_reference_golden(golden, sched):
    from diffsep_b200 import sdes
    from diffsep_b200.pl_model import normalize_batch
    g = golden("sampler.npz")
    (mix, _), _, _ = normalize_batch((cases.batch_mix(2, 1024).to(DEV), None))
    sde = sdes.MixSDE(ndim=2, d_lambda=2.0, sigma_min=0.05, sigma_max=0.5, N=10)
    with sdes.injected_noise(cases.sampler_noises(2, 1024, 10, 1)):
        out, _ = sdes.get_pc_scheduled_sampler("reverse_diffusion", "ald2", sde=sde, score_fn=cases.analytic_score,
                                               y=mix, eps=0.03, snr=0.5, corrector_steps=1, denoise=False,
                                               schedule=sched)()
    assert rel_l2(out.cpu(), g[f"sched.{sched}"]) < 1e-5


def test_registry_errors():
    from diffsep_b200 import sdes
    sde = sdes.MixSDE(2, 2.0, 0.05, 0.5, N=3)
    y = torch.zeros(1, 1, 64, device=DEV)
    with pytest.raises(ValueError):
        sdes.get_pc_sampler("nope", "ald2", sde=sde, score_fn=cases.analytic_score, y=y)
    with pytest.raises(ValueError):
        sdes.get_pc_sampler("reverse_diffusion", "nope", sde=sde, score_fn=cases.analytic_score, y=y)
    with pytest.raises(NotImplementedError):
        sdes.get_pc_scheduled_sampler("reverse_diffusion", "ald2", sde=sde, score_fn=cases.analytic_score, y=y,
                                      schedule="fib")


def test_sampler_with_network_vs_oracle():
    """MixSDE, N=3, 1 corrector step, nf=64 network, injected noise: per-step and free-running vs the CPU oracle."""
    _network_sampler_vs_oracle(None, False, 4096, 3)


def test_minibatch_sampler_equals_full_batch():
    from diffsep_b200 import sdes
    from diffsep_b200.pl_model import DiffSepModel, DEFAULT_CONFIG, normalize_batch
    from oracle import weights as ow
    import copy
    cfg = copy.deepcopy(DEFAULT_CONFIG)
    cfg["model"]["score_model"]["backbone_args"]["nf"] = 64
    model = DiffSepModel(cfg, score_state_dict=ow.make_score_model_state_dict(nf=64, seed=0))
    (mix, _), _, _ = normalize_batch((cases.batch_mix(3, 2048).to(DEV), None))
    noises = cases.sampler_noises(3, 2048, 2, 1)
    with sdes.injected_noise(noises):
        full, _ = model.get_pc_sampler("reverse_diffusion", "ald2", mix, N=2, corrector_steps=1, snr=0.5)()
    per_item = []
    for b in range(3):
        with sdes.injected_noise([z[b:b + 1] for z in noises]):
            o, _ = model.get_pc_sampler("reverse_diffusion", "ald2", mix[b:b + 1].contiguous(), N=2,
                                        corrector_steps=1, snr=0.5)()
        per_item.append(o)
    torch.cuda.synchronize()
    # utterances are independent end to end (what makes the multi-GPU sharding exact)
    assert rel_l2(torch.cat(per_item).cpu(), full.cpu()) < 1e-5


def test_evaluate_batch_driver(tmp_path):
    """evaluate.py: per-batch nfe / runtime / len_s JSON (reference evaluate.py:394-406), wavs of each utterance's own
    length.  Default batching puts only equal-length utterances together (batch-of-one results, like the reference's
    loop); --pad-batches pads ragged batches like max_collator and crops the estimates back."""
    import copy, json, subprocess, sys
    from pathlib import Path
    import numpy as np
    from scipy.io import wavfile
    from oracle import weights as ow
    from diffsep_b200.pl_model import DEFAULT_CONFIG
    root = Path(__file__).resolve().parent.parent
    cfg = copy.deepcopy(DEFAULT_CONFIG)
    cfg["model"]["score_model"]["backbone_args"]["nf"] = 64
    sd_ = ow.make_score_model_state_dict(nf=64, seed=0)
    torch.save({"state_dict": {"score_model." + k: v for k, v in sd_.items()}, "hyper_parameters": {"config": cfg}},
               tmp_path / "ckpt.pt")
    (tmp_path / "in").mkdir()
    lens = [6000, 8000, 6000, 7000]
    for i, n in enumerate(lens):
        wavfile.write(tmp_path / "in" / f"u{i}.wav", 8000, (cases.synthetic_mix(i, n)[0] * 0.5).numpy().astype(np.float32))

    def run(out, *extra):
        r = subprocess.run([sys.executable, str(root / "evaluate.py"), str(tmp_path / "in"), str(tmp_path / out),
                            "--model", str(tmp_path / "ckpt.pt"), "-N", "2", "--batch-size", "2", *extra],
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout + r.stderr
        res = json.loads((tmp_path / out / "results.json").read_text())
        assert all(b["nfe"] == 4 and b["runtime"] > 0 for b in res)
        summ = json.loads((tmp_path / out / "results_summary.json").read_text())
        assert summ["utterances"] == len(lens) and summ["utt_per_s"] > 0
        wavs = {}
        for i, n in enumerate(lens):
            for s in (0, 1):
                sr_, data = wavfile.read(tmp_path / out / f"s{s}" / f"u{i}.wav")
                assert sr_ == 8000 and data.shape == (n,) and np.isfinite(data).all()
                wavs[i, s] = data
        return res, wavs

    res, _ = run("out")                                       # equal-length buckets: {u0, u2}, {u1}, {u3}
    assert [b["files"] for b in res] == [["u0.wav", "u2.wav"], ["u1.wav"], ["u3.wav"]]
    assert res[0]["len_s"] == [0.75, 0.75] and res[1]["len_s"] == [1.0] and res[2]["len_s"] == [0.875]
    res_p, _ = run("out_pad", "--pad-batches")                # folder order, padded: {u0, u1}, {u2, u3}
    assert [len(b["files"]) for b in res_p] == [2, 2] and res_p[0]["len_s"] == [0.75, 1.0]


@pytest.mark.parametrize("case", [("em_ald2", "euler_maruyama", "ald2", "mix", 1, False),
                                  ("rd_ald", "reverse_diffusion", "ald", "mix", 2, False),
                                  ("rd_langevin", "reverse_diffusion", "langevin", "mix", 1, False),
                                  ("rd_langevin_prior", "reverse_diffusion", "langevin", "priormix", 1, False),
                                  ("rd_ald2_pflow", "reverse_diffusion", "ald2", "mix", 1, True),
                                  ("em_ald2_pflow", "euler_maruyama", "ald2", "priormix", 1, True),
                                  ("none_ald2", "none", "ald2", "mix", 1, False)], ids=lambda c: c[0])
def test_sampler_other_plugins_match_reference_golden(golden, case):
    """The remaining registered plugins against the real reference sampler."""
    from diffsep_b200 import sdes
    from diffsep_b200.pl_model import normalize_batch
    name, pred, corr, sde_name, cs, pflow = case
    g = golden("plugins.npz")
    (mix, _), _, _ = normalize_batch((cases.batch_mix(2, 1024).to(DEV), None))
    sde = sdes.SDERegistry.get_by_name(sde_name)(ndim=2, d_lambda=2.0, sigma_min=0.05, sigma_max=0.5, N=10)
    with sdes.injected_noise(cases.sampler_noises(2, 1024, 10, cs)):
        out, nfe = sdes.get_pc_sampler(pred, corr, sde=sde, score_fn=cases.analytic_score, y=mix, eps=0.03, snr=0.5,
                                       corrector_steps=cs, denoise=False, probability_flow=pflow)()
    torch.cuda.synchronize()
    assert nfe == 10 * (cs + 1)
    assert rel_l2(out.cpu(), g[name]) < 1e-5


@pytest.mark.parametrize("case", cases.NDIM_CASES, ids=lambda c: c[0])
def test_three_sources_and_true_mean_prior_match_reference_golden(golden, case):
    """3-source sampling through PriorMixSDE (ndim = 3) and the sampler's ``true_mean`` prior branch, against the
    real reference sampler (tests/golden/make_golden_ndim.py): the fused update kernels with NC = 3, per-channel
    sigma_mix in the prior, MixSDE's 0.5 * true_mean quirk."""
    from diffsep_b200 import sdes
    from diffsep_b200.pl_model import normalize_batch
    name, sde_name, ndim, tm_ch, cs = case
    g = golden("ndim.npz")
    (mix, _), _, _ = normalize_batch((cases.batch_mix(cases.NDIM_B, cases.NDIM_T).to(DEV), None))
    sde = sdes.SDERegistry.get_by_name(sde_name)(ndim=ndim, d_lambda=2.0, sigma_min=0.05, sigma_max=0.5,
                                                 N=cases.NDIM_N)
    tm = cases.ndim_true_mean(tm_ch).to(DEV) if tm_ch else None
    with sdes.injected_noise(cases.ndim_noises(ndim, cs)):
        out, nfe = sdes.get_pc_sampler("reverse_diffusion", "ald2", sde=sde, score_fn=cases.analytic_score, y=mix,
                                       true_mean=tm, eps=0.03, snr=0.5, corrector_steps=cs, denoise=True)()
    torch.cuda.synchronize()
    assert nfe == cases.NDIM_N * (cs + 1) and tuple(out.shape) == (cases.NDIM_B, ndim, cases.NDIM_T)
    assert rel_l2(out.cpu(), g[name]) < 1e-5


def test_three_source_limits_follow_the_reference():
    """MixSDE.prior_sampling hard-codes two sources in the reference (sdes.py:344): ndim = 3 constructs but cannot
    sample; PriorMixSDE rejects inputs with neither 1 nor ndim channels with the reference's message (:579-583)."""
    from diffsep_b200 import sdes
    y = torch.zeros(1, 1, 256, device=DEV)
    with pytest.raises(RuntimeError):
        sdes.MixSDE(3, 2.0, 0.05, 0.5, N=2).prior_sampling(y.shape, y)
    with pytest.raises(ValueError, match="should have 1 channel"):
        y2 = torch.zeros(1, 2, 256, device=DEV)
        sdes.PriorMixSDE(3, 2.0, 0.05, 0.5, N=2).prior_sampling(y2.shape, y2)
    with pytest.raises(NotImplementedError):
        sdes.MixSDE(4, 2.0, 0.05, 0.5)


def test_ald_rejects_priormix_like_the_reference():
    from diffsep_b200 import sdes
    sde = sdes.PriorMixSDE(2, 2.0, 0.05, 0.5, N=3)
    with pytest.raises(NotImplementedError):
        sdes.get_pc_sampler("reverse_diffusion", "ald", sde=sde, score_fn=cases.analytic_score,
                            y=torch.zeros(1, 1, 64, device=DEV))


def test_probability_flow_kernel_switch():
    """dsep_sde_predictor(probability_flow=1): half the score term, no noise (sdes.py:143-152,167-170) —
    reachable through sde.reverse(score_fn, probability_flow=True), not through the predictors."""
    from diffsep_b200 import ops
    from oracle import sde_ref as sd
    B, T = 2, 512
    g = cases.gen(14)
    mix, x, score, z = (torch.randn(B, c, T, generator=g) for c in (1, 2, 2, 2))
    t = torch.tensor([0.9, 0.2])
    p = sd.MixSDEParams(N=10)
    want_x, want_m = sd.predictor_step(p, lambda *_: 0.5 * score, x, t, mix, torch.zeros_like(z))
    xo, xm = torch.empty(B, 2, T, device=DEV), torch.empty(B, 2, T, device=DEV)
    ops.sde_predictor(ops.sde_params(2.0, 0.05, 0.5), x.to(DEV), score.to(DEV), t.to(DEV), None, z.to(DEV), 0, 0,
                      0.1, B, T, xo, xm, probability_flow=True)
    torch.cuda.synchronize()
    assert rel_l2(xo.cpu(), want_x) < 2e-6 and rel_l2(xm.cpu(), want_m) < 2e-6


def test_per_step_parity_benchmark_architecture():
    """The north-star criterion at the benchmark shape (nf=128, 4 s @ 8 kHz, [1,6,256,256] spectrograms): each
    corrector / predictor update, restarted from the oracle's state, within 1e-4 rel-L2 of the CPU oracle
    (measured 1.5e-5 / 6e-6 over a full N=30 run, profiles/parity_r01.md; 3 steps here to stay short)."""
    from diffsep_b200 import sdes
    from diffsep_b200.pl_model import DEFAULT_CONFIG, DiffSepModel, normalize_batch
    from oracle import score_ref as sr, sde_ref as sd, weights as ow
    N, T = 3, 32000
    model = DiffSepModel(DEFAULT_CONFIG, score_state_dict=ow.make_score_model_state_dict(nf=128, seed=0))
    params = ow.make_backbone_params(nf=128, seed=0)
    mix_cpu, _, _ = sd.normalize_batch(cases.batch_mix(1, T))
    (mix, _), _, _ = normalize_batch((cases.batch_mix(1, T).to(DEV), None))
    nz = cases.sampler_noises(1, T, N, 1)
    p = sd.MixSDEParams(N=N)
    sde = sdes.MixSDE(2, 2.0, 0.05, 0.5, N=N)

    def score_cpu(x, t, m):
        with torch.no_grad():
            return sr.score_forward(params, x, t, m)
    ts = sd.timesteps(p, 0.03)
    x = sd.prior_sampling(p, mix_cpu, nz.pop(0))
    with model.cached_mixture(mix):
        for i in range(N):
            vt = torch.ones(1) * ts[i]
            vt_d = vt.to(DEV)
            zc, zp = nz.pop(0), nz.pop(0)
            xc, _ = sd.corrector_step(p, score_cpu, x, vt, mix_cpu, [zc], 0.5)
            with sdes.injected_noise([zc]):
                g, _ = sde.corrector_update(x.to(DEV), model(x.to(DEV), vt_d, mix), vt_d, mix, 0.5)
            assert rel_l2(g.cpu(), xc) < 1e-4
            xp, _ = sd.predictor_step(p, score_cpu, xc, vt, mix_cpu, zp)
            with sdes.injected_noise([zp]):
                g, _ = sde.predictor_update(xc.to(DEV), model(xc.to(DEV), vt_d, mix), vt_d, mix, 1.0 / N)
            assert rel_l2(g.cpu(), xp) < 1e-4
            x = xp


def test_shape_cache_is_bounded():
    """ragged inputs: buffers / plans / graphs of at most MAX_SHAPES (B, T) shapes stay resident"""
    sm = _score_model(64)
    for T in (1000, 1500, 2000, 2500, 3000, 9000, 1000):
        xt, t, mix = (v.to(DEV) for v in cases.score_inputs(1, T, seed=T))
        y = sm(xt, t, mix)
        assert y.shape == (1, 2, T) and bool(torch.isfinite(y).all())
    assert len(sm._bufs) <= sm.MAX_SHAPES and len(sm.backbone._plans) <= sm.MAX_SHAPES


def test_compute_score_loss_forward_vs_oracle():
    """DiffSepModel.compute_score_loss (forward only): perturb kernel + nf=64 score network + loss
    kernel against the oracle's sample_prior / score_forward / score_loss on the same times and injected noise."""
    import copy
    from diffsep_b200 import sdes
    from diffsep_b200.pl_model import DEFAULT_CONFIG, DiffSepModel
    from oracle import score_ref as sr, sde_ref as sd, weights as ow
    cfg = copy.deepcopy(DEFAULT_CONFIG)
    cfg["model"]["score_model"]["backbone_args"]["nf"] = 64
    model = DiffSepModel(cfg, score_state_dict=ow.make_score_model_state_dict(nf=64, seed=0))
    params = ow.make_backbone_params(nf=64, seed=0)
    B, T = 2, 4096
    g = cases.gen(77)
    target = torch.randn(B, 2, T, generator=g) * 0.2
    mix = target.sum(dim=1, keepdim=True)
    time = torch.tensor([0.2, 0.9])
    z = torch.randn(B, 2, T, generator=g)
    p = sd.MixSDEParams(N=30)
    xt_w = sd.sample_prior(p, mix, target, time, z)
    with torch.no_grad():
        score_w = sr.score_forward(params, xt_w, time, mix)
    want = sd.score_loss(p, score_w, z, time, mix, "none")
    with sdes.injected_noise([z]):
        got = model.compute_score_loss(mix.to(DEV), target.to(DEV), time=time.to(DEV), reduction="none")
    torch.cuda.synchronize()
    assert rel_l2(got.cpu(), want) < 1e-4
    t2 = model.sample_time(target.to(DEV))
    assert t2.shape == (B,) and float(t2.min()) >= model.t_eps and float(t2.max()) <= model.t_max


def test_stft_gemms_on_the_tensor_core_match_the_fp32_gemm():
    """ScoreModelNCSNpp._dft: the DFT-510 / inverse products as a 1x1 convolution with three fp16 tensor-core products
    (fp32-grade) against the fp32 CUDA-core GEMM and a float64 product, forward and inverse, at the benchmark's row
    count (B * ns * Fr = 64 * 251), at row counts that are not whole 8-row lines / fewer than 128 with the model's padded
    buffers (still the tensor core: the arithmetic must not depend on the batch size, so a row's result is the same
    bit for bit whatever the batch it sits in), and with exact-size buffers (falls back to the GEMM)."""
    from diffsep_b200 import ops
    from diffsep_b200.score_model import ScoreModelNCSNpp, LD
    sm = ScoreModelNCSNpp(num_sources=2, backbone_args=dict(nf=64))
    g = cases.gen(5)
    for M, padded in ((64 * 251, True), (2 * 251, True), (65, True), (2 * 251, False)):
        Mb = sm.dft_rows(M) if padded else M
        src = torch.zeros(Mb, LD, device=DEV)
        src[:M] = torch.randn(M, LD, generator=g).to(DEV)
        src[:, 510:] = 0.0
        for which, basis in (("fwd", sm.basis_fwd), ("inv", sm.basis_inv)):
            if which == "inv":      # the inverse product stays an fp32 GEMM: its operand (the decompressed score
                src[:M] *= 3.0e4    # spectrogram) is not bounded by fp16's range — 2e5 and more occur
            want = (src[:M].double() @ basis.double()).cpu()
            ref = torch.empty(M, LD, device=DEV)
            ops.sgemm(src, LD, basis, LD, ref, LD, M, LD, LD)
            got = torch.full((Mb, LD), float("nan"), device=DEV)
            sm._dft(src, which, got, M)
            torch.cuda.synchronize()
            assert rel_l2(ref.cpu(), want) < 1e-6
            assert rel_l2(got[:M].cpu(), want) < 2e-6, (M, which)
            if padded:      # a row's result does not depend on how many rows come with it
                big = torch.zeros(sm.dft_rows(4 * M), LD, device=DEV)
                big[M:2 * M] = src[:M]
                out = torch.empty_like(big)
                sm._dft(big, which, out, 4 * M)
                torch.cuda.synchronize()
                assert torch.equal(out[M:2 * M], got[:M]), (M, which)
