"""CPU: the C-ABI library loads and exports every symbol include/dsep.h declares (no compute
calls), host-side plugin logic (registries, time grids, noise injection, config handling), and the
world-size-2 sharding path over gloo."""
import ctypes
import os
import re
import subprocess
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    from diffsep_b200 import build
    build.build()
    from diffsep_b200 import _lib
    return _lib.load()


def test_abi_exports_every_declared_symbol(lib):
    from diffsep_b200 import _lib
    header = (ROOT / "include" / "dsep.h").read_text()
    declared = set(re.findall(r"\b(dsep_[a-z0-9_]+)\s*\(", header)) - {"dsep_stream_t"}
    assert len(declared) >= 26
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(_lib.PROTOTYPES) | set(_lib.OTHER_SYMBOLS)
    assert lib.dsep_abi_version() == _lib.ABI_VERSION == 8


def test_abi_argument_counts_match_header():
    from diffsep_b200 import _lib
    header = (ROOT / "include" / "dsep.h").read_text()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    for name, argtypes in _lib.PROTOTYPES.items():
        m = re.search(r"\bint\s+" + name + r"\s*\((.*?)\)\s*;", header, flags=re.S)
        assert m, name
        assert len([a for a in m.group(1).split(",") if a.strip()]) == len(argtypes), name


def test_no_fallback_without_gpu(lib):
    """The product path fails loudly instead of falling back when there is no device."""
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from diffsep_b200 import ops
    with pytest.raises(RuntimeError):
        ops.require_device()
    from diffsep_b200.score_model import ScoreModelNCSNpp
    with pytest.raises(RuntimeError):
        ScoreModelNCSNpp(num_sources=2, backbone_args=dict(nf=64))
    with pytest.raises(ValueError):
        ops.ptr(torch.zeros(4))       # CPU tensors are rejected, never computed on


def test_product_does_not_import_oracle():
    bad = []
    for p in list((ROOT / "diffsep_b200").rglob("*.py")) + [ROOT / "separate.py"]:
        if re.search(r"^\s*(from|import)\s+oracle\b", p.read_text(), flags=re.M):
            bad.append(str(p))
    assert not bad, bad


def test_synthetic_weights_match_oracle_copy():
    from diffsep_b200 import synthetic
    from oracle import weights as ow
    a = synthetic.make_score_model_state_dict(nf=32, seed=3)
    b = ow.make_score_model_state_dict(nf=32, seed=3)
    assert list(a) == list(b)
    assert all(torch.equal(a[k], b[k]) for k in a)


def test_registries_and_errors():
    from diffsep_b200 import sdes
    assert set(sdes.PredictorRegistry.get_all_names()) == {"euler_maruyama", "reverse_diffusion", "none"}
    assert set(sdes.CorrectorRegistry.get_all_names()) == {"ald2", "ald", "langevin", "none"}
    assert set(sdes.SDERegistry.get_all_names()) == {"mix", "priormix"}
    with pytest.raises(ValueError, match="unknown"):
        sdes.PredictorRegistry.get_by_name("heun")
    with pytest.warns(UserWarning, match="doubly registered"):
        @sdes.PredictorRegistry.register("none")
        class Again(sdes.predictors.NonePredictor):
            pass
    sde = sdes.MixSDE(ndim=2, d_lambda=2.0, sigma_min=0.05, sigma_max=0.5, N=30)
    c = sde.copy()
    c.N = 5
    assert sde.N == 30 and c.N == 5 and c.T == 1.0
    assert sdes.PriorMixSDE(ndim=3, d_lambda=2.0, sigma_min=0.05, sigma_max=0.5).ndim == 3    # 3-speaker models
    with pytest.raises(NotImplementedError):
        sdes.MixSDE(ndim=4, d_lambda=2.0, sigma_min=0.05, sigma_max=0.5)
    class OtherSDE(sdes.SDE):
        pass
    with pytest.raises(NotImplementedError, match="not yet supported"):
        sdes.correctors.AnnealedLangevinDynamics2(OtherSDE(2, 2.0, 0.05, 0.5), None, 0.5, 1)


def test_time_grids_match_oracle():
    from diffsep_b200 import sdes
    from oracle import sde_ref as sd
    sde = sdes.MixSDE(2, 2.0, 0.05, 0.5, N=30)
    p = sd.MixSDEParams(N=30)
    for sched in (None, "linear", "log", "revlog"):
        assert torch.equal(sdes._timesteps(sde, 0.03, sched, "cpu"), sd.timesteps(p, 0.03, sched))
    with pytest.raises(NotImplementedError):
        sdes._timesteps(sde, 0.03, "fib", "cpu")


def test_noise_injection_order_and_exhaustion():
    from diffsep_b200.sdes import noise
    zs = [torch.full((1, 2, 4), float(i)) for i in range(2)]
    with noise.injected_noise(zs):
        a, _, _ = noise.SOURCE.next((1, 2, 4), "cpu")
        b, _, _ = noise.SOURCE.next((1, 2, 4), "cpu")
        assert float(a[0, 0, 0]) == 0.0 and float(b[0, 0, 0]) == 1.0
        with pytest.raises(RuntimeError):
            noise.SOURCE.next((1, 2, 4), "cpu")
    z, seed, off = noise.SOURCE.next((1, 2, 4), "cpu")
    assert z is None and off > 0
    with noise.injected_noise([torch.zeros(1, 2, 5)]):
        with pytest.raises(ValueError):
            noise.SOURCE.next((1, 2, 4), "cpu")


def test_shard_bounds_cover_batch_exactly():
    from diffsep_b200.shard import shard_bounds
    for n in (0, 1, 7, 32, 256, 257):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) <= -(-n // world) if n else True


_GLOO_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["DSEP_ROOT"])
from diffsep_b200.shard import shard_bounds, gather_estimates
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
for n in (5, 4, 1):
    full = torch.arange(n * 2 * 3, dtype=torch.float32).reshape(n, 2, 3)
    lo, hi = shard_bounds(n, rank, world)
    local = full[lo:hi] * 2.0            # stand-in for the per-shard separation
    out = gather_estimates(local, n)
    assert out.shape == full.shape and torch.equal(out, full * 2.0), (rank, n)
dist.barrier()
dist.destroy_process_group()
print("ok", rank)
"""


def test_sharded_gather_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    env = dict(os.environ, DSEP_ROOT=str(ROOT), MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29613", str(script)],
                       env=env, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 2


def test_config_handling_without_hydra():
    from diffsep_b200.pl_model import DEFAULT_CONFIG, _ns, _to_plain
    ns = _ns(DEFAULT_CONFIG)
    assert ns.model.sampler.N == 30 and ns.model.fs == 8000
    assert _to_plain(ns) == DEFAULT_CONFIG


def test_max_collator_matches_reference_semantics():
    """centre padding with off // 2 zeros in front (reference datasets/wsj0_mix.py:95-111)"""
    from diffsep_b200.data import max_collator, uncollate
    sig = [torch.arange(1, 6, dtype=torch.float32)[None], torch.arange(1, 10, dtype=torch.float32)[None],
           torch.arange(1, 3, dtype=torch.float32)[None]]
    batch, spans = max_collator(sig)
    assert batch.shape == (3, 1, 9)
    assert spans == [(2, 5), (0, 9), (3, 2)]
    assert batch[0, 0].tolist() == [0, 0, 1, 2, 3, 4, 5, 0, 0]
    assert batch[2, 0].tolist() == [0, 0, 0, 1, 2, 0, 0, 0, 0]
    back = uncollate(batch, spans)
    assert all(torch.equal(a, b) for a, b in zip(back, sig))


def test_equal_length_bucketing_for_the_batch_driver(tmp_path):
    """evaluate.py's default batching: only utterances of the same length share a batch (the reference evaluates one
    utterance at a time, evaluate.py:340-376; zero padding would enter normalisation / STFT / GroupNorm)."""
    from diffsep_b200.data import bucket_by_length, save_wav, wav_length
    lengths = [800, 640, 800, 800, 640, 1000, 800]
    batches = bucket_by_length(lengths, 2)
    assert batches == [[0, 2], [3, 6], [1, 4], [5]]
    assert all(len({lengths[i] for i in b}) == 1 for b in batches)
    assert sorted(i for b in batches for i in b) == list(range(len(lengths)))
    save_wav(tmp_path / "a.wav", torch.zeros(1, 777), 8000)
    assert wav_length(tmp_path / "a.wav") == 777


def test_library_carries_the_hash_of_its_sources(lib):
    """build() is gated on a hash of the sources + flags, embedded in the binary (not on file times)."""
    from diffsep_b200 import build as b
    assert lib.dsep_source_hash().decode() == b.source_hash() == b.built_hash()
    assert not b.needs_build()


def test_wav_io_round_trip(tmp_path):
    import numpy as np
    from scipy.io import wavfile
    from diffsep_b200.data import load_wav, save_wav
    x = torch.linspace(-0.5, 0.5, 800)[None]
    save_wav(tmp_path / "a.wav", x, 8000)
    y, sr = load_wav(tmp_path / "a.wav")
    assert sr == 8000 and torch.allclose(x, y)
    wavfile.write(tmp_path / "b.wav", 8000, (x[0].numpy() * 32768).astype(np.int16))
    z, _ = load_wav(tmp_path / "b.wav")
    assert float((z - x).abs().max()) < 1e-4


def _lightning_checkpoint(raw, ema_sd, ema_names, tmp_path=None):
    """A checkpoint dict shaped like what the reference writes: Lightning's top-level keys, ``state_dict`` with the
    ``score_model.`` prefix (+ the STFT window buffers), ``hyper_parameters.config``, and ``ema`` =
    ``torch_ema.ExponentialMovingAverage.state_dict()`` as stored by ``on_save_checkpoint`` (pl_model.py:672-673)."""
    from diffsep_b200.pl_model import DEFAULT_CONFIG
    ckpt = {
        "epoch": 644, "global_step": 1_000_000, "pytorch-lightning_version": "1.6.4",
        "state_dict": {**{"score_model." + k: v for k, v in raw.items()}, "loss.dummy_buffer": torch.zeros(1)},
        "loops": {}, "callbacks": {}, "optimizer_states": [{}], "lr_schedulers": [],
        "hyper_parameters": {"config": DEFAULT_CONFIG},
        "ema": {"decay": 0.999, "num_updates": 1_000_000, "shadow_params": [ema_sd[k].clone() for k in ema_names],
                "collected_params": None},
    }
    if tmp_path is not None:                       # through torch.save / torch.load like load_from_checkpoint
        torch.save(ckpt, tmp_path / "checkpoint.pt")
        ckpt = torch.load(tmp_path / "checkpoint.pt", map_location="cpu", weights_only=False)
    return ckpt


def test_checkpoint_state_swaps_ema_weights_in(tmp_path):
    """load_from_checkpoint's host logic (reference pl_model.py:642-673): the `score_model.` prefix is stripped and the
    EMA shadow list (parameters() order) replaces the raw weights, for BOTH torch_ema layouts — <= 0.2 (only
    requires_grad parameters: the frozen Fourier W absent) and 0.3 (every parameter, W included).  A count that
    matches neither is an error (the reference always evaluates EMA weights), as is a shape mismatch; no_ema and a
    missing `ema` entry give the raw weights (the latter with the reference's warning)."""
    import warnings
    from oracle import weights as ow
    from diffsep_b200.pl_model import DEFAULT_CONFIG, checkpoint_state
    ema_sd = ow.make_score_model_state_dict(nf=32, seed=0)
    raw = ow.make_score_model_state_dict(nf=32, seed=1)
    w = "backbone.all_modules.0.W"
    every = [k for k in ema_sd if k.startswith("backbone.")]
    trainable = [k for k in every if k != w]
    assert len(every) == len(trainable) + 1
    # torch_ema <= 0.2 layout
    ckpt = _lightning_checkpoint(raw, ema_sd, trainable, tmp_path)
    config, sd = checkpoint_state(ckpt)
    assert config == DEFAULT_CONFIG and set(sd) == set(raw)
    assert all(torch.equal(sd[k], ema_sd[k]) for k in trainable)
    assert torch.equal(sd[w], raw[w])                      # not tracked: stays raw
    assert torch.equal(sd["stft.window"], raw["stft.window"])
    # torch_ema 0.3 layout: W is in the list too (equal to the raw W in a real checkpoint; here it shows it is taken)
    ckpt3 = _lightning_checkpoint(raw, ema_sd, every)
    _, sd3 = checkpoint_state(ckpt3)
    assert all(torch.equal(sd3[k], ema_sd[k]) for k in every)
    # eval(no_ema=True) / no EMA entry: raw weights
    _, sd_raw = checkpoint_state(ckpt, no_ema=True)
    assert all(torch.equal(sd_raw[k], raw[k]) for k in every)
    with warnings.catch_warnings(record=True) as rec:
        warnings.simplefilter("always")
        _, sd2 = checkpoint_state({"state_dict": ckpt["state_dict"]})
    assert rec and "EMA state_dict not found" in str(rec[0].message)
    assert all(torch.equal(sd2[k], raw[k]) for k in every)
    # a count matching neither layout: error, not a silent fall-back to raw weights
    bad = dict(ckpt, ema=dict(ckpt["ema"], shadow_params=ckpt["ema"]["shadow_params"][:-1]))
    with pytest.raises(ValueError, match="EMA shadow parameters"):
        checkpoint_state(bad)
    # wrong shape: error
    shadow = list(ckpt["ema"]["shadow_params"])
    shadow[3] = torch.zeros(7)
    with pytest.raises(ValueError):
        checkpoint_state(dict(ckpt, ema=dict(ckpt["ema"], shadow_params=shadow)))


def test_noise_source_restarts_with_the_seed():
    """torch.manual_seed(s) before two runs in one process gives the same Philox (seed, offset) sequence, as the
    reference's randn does; a different seed a different key."""
    from diffsep_b200.sdes.noise import NoiseSource
    src = NoiseSource()
    torch.manual_seed(11)
    a = [src.next((1, 2, 8), "cpu")[1:] for _ in range(3)]
    torch.manual_seed(11)
    b = [src.next((1, 2, 8), "cpu")[1:] for _ in range(3)]
    torch.manual_seed(12)
    c = [src.next((1, 2, 8), "cpu")[1:] for _ in range(3)]
    assert a == b and a != c and len({o for _, o in a}) == 3 and {k for k, _ in a} == {11}


def test_fp8_correction_entry_point_is_shipped_and_validates(lib):
    """The e4m3-correction mode (passes = 2, the product default) is in the shipped library; its entry point
    validates its arguments on the host before any launch, and the regular entry point rejects passes = 2."""
    import ctypes as C
    import diffsep_b200
    from diffsep_b200 import _lib
    assert lib.dsep_has_fp8_corr() == 1
    assert diffsep_b200.DEFAULT_PASSES in (1, 2, 3)
    g = _lib.ConvArgs()
    g.passes = 2
    with pytest.raises(ValueError):          # null operands
        _lib.call("dsep_conv2d_fused8", C.byref(g), 2.0 ** -11, 3, None)
    with pytest.raises(ValueError):
        _lib.call("dsep_conv2d_fused", C.byref(g), None)
    g.passes = 3
    with pytest.raises(ValueError):          # fused8 is passes = 2 only
        _lib.call("dsep_conv2d_fused8", C.byref(g), 2.0 ** -11, 3, None)


def test_fp8_correction_planes_and_scales_reproduce_the_product():
    """Host side of the e4m3-correction mode (ConvWeight.planes8 / corr_rel / A8_EXP), checked by emulating the
    kernel's arithmetic in float64: acc_scale * (A_hi . W_hi + corr_rel * ([A_lo8 | A_hi8] . [W_hi8 ; W_lo8]))
    equals A . W to ~1e-5, with the bytes laid out as include/dsep.h documents (16-byte groups of 8 channels)."""
    from diffsep_b200.backbone import ConvWeight, _prescale_exp
    from diffsep_b200.ops import Split
    g = torch.Generator().manual_seed(3)
    cout, cin = 64, 128
    w = torch.randn(cout, cin, generator=g) / cin ** 0.5
    k = _prescale_exp(float(w.abs().max()))
    ws = w * 2.0 ** k
    w_hi = ws.half()
    w_lo = (ws - w_hi.float()).half()
    cw = object.__new__(ConvWeight)
    cw.planes = Split(w_hi.reshape(1, cout, cin), w_lo.reshape(1, cout, cin))
    p8 = cw.planes8()
    assert p8.hi is cw.planes.hi and p8.lo.shape == (1, cout, cin) and p8.lo.dtype == torch.float16
    raw = p8.lo.view(torch.uint8).reshape(cout, cin // 8, 16)
    dec = raw.view(torch.float8_e4m3fn).double()
    w_hi8, w_lo8 = dec[..., :8].reshape(cout, cin), dec[..., 8:].reshape(cout, cin)
    assert torch.allclose(w_hi8, w_hi.double() * 2.0 ** cw.W8_EXP, rtol=2.0 ** -4, atol=2.0 ** -9)
    assert float(w_hi8.abs().max()) <= 448 and float(w_lo8.abs().max()) <= 448
    # activations as the kernel's builder prepares them (SiLU outputs up to a few units)
    a = torch.randn(256, cin, generator=g)
    a = a * torch.sigmoid(a) * 3.0
    a_hi = a.half().float()
    e4 = lambda v: v.clamp(-448, 448).to(torch.float8_e4m3fn).double()
    a_hi8, a_lo8 = e4(a_hi * 2.0 ** cw.A8_EXP), e4((a - a_hi) * 2.0 ** (cw.A8_EXP + 11))
    main = a_hi.double() @ w_hi.double().T
    corr = a_lo8 @ w_hi8.T + a_hi8 @ w_lo8.T
    got = 2.0 ** -k * (main + cw.corr_rel * corr)
    ref = a.double() @ w.double().T
    err = float((got - ref).norm() / ref.norm())
    err_1pass = float((2.0 ** -k * main - ref).norm() / ref.norm())
    assert err < 2e-5 < 2e-4 < err_1pass, (err, err_1pass)


def test_numerics_study_e4m3_emulator_matches_torch_float8():
    """tools/numerics_study.py's e4m3 rounding (the evidence behind the 2-unit conv mode) is bit-identical to
    torch.float8_e4m3fn on power-of-two-prescaled data, subnormals and saturation included."""
    import importlib.util
    import math
    spec = importlib.util.spec_from_file_location("numerics_study", ROOT / "tools" / "numerics_study.py")
    ns = importlib.util.module_from_spec(spec)
    argv, sys.argv = sys.argv, ["numerics_study.py"]
    try:
        spec.loader.exec_module(ns)
    finally:
        sys.argv = argv
    g = torch.Generator().manual_seed(0)
    x = torch.randn(50000, generator=g, dtype=torch.float64) * torch.exp(2 * torch.randn(50000, generator=g, dtype=torch.float64))
    k = math.floor(math.log2(448.0 / float(x.abs().max())))
    ref = (x * 2.0 ** k).float().clamp(-448, 448).to(torch.float8_e4m3fn).double() * 2.0 ** -k
    assert torch.equal(ns.e4m3(x), ref)


def test_dft_row_padding_is_batch_size_independent():
    """ScoreModelNCSNpp.dft_rows: the forward DFT product always runs in its tensor-core form on whole 8-row lines
    (at least 128 rows), so the arithmetic a spectrogram row sees cannot depend on how many utterances share the
    batch — a shard must reproduce the whole batch (the 2-GPU gather test on hardware)."""
    from diffsep_b200.score_model import ScoreModelNCSNpp
    r = ScoreModelNCSNpp.dft_rows
    for m in (1, 63, 126, 127, 128, 129, 260, 520, 16064):
        assert r(m) >= max(m, 128) and r(m) % 8 == 0 and r(m) - m < 128
    assert r(260) == 264 and r(520) == 520 and r(65) == 128


def test_narrow_conv_as_1x1_plus_tap_gather_algebra():
    """The layout convention of the output-pyramid convs on large maps (backbone.py ``pyr_taps`` + dsep_tap_gather3x3,
    ncsnpp.py:419-440): row ``tap * CO + co`` of the 1x1 weight is ``W[co, :, ky, kx]`` (tap = ky * 3 + kx) and
    ``out[h, w, co] = bias[co] + sum_tap z[h + ky - 1, w + kx - 1, tap * CO + co]`` over in-image neighbours — checked
    here in plain torch against F.conv2d, so the convention the CUDA gather implements is pinned without a GPU."""
    import torch
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(11)
    B, C, H, W, CO = 2, 16, 7, 9, 6
    a = torch.randn(B, C, H, W, generator=g, dtype=torch.float64)
    w = torch.randn(CO, C, 3, 3, generator=g, dtype=torch.float64)
    bias = torch.randn(CO, generator=g, dtype=torch.float64)
    w_taps = w.permute(2, 3, 0, 1).reshape(9 * CO, C, 1, 1)              # what backbone.py feeds the 1x1 convolution
    z = F.conv2d(a, w_taps)                                              # [B, 54, H, W]
    zp = F.pad(z, (1, 1, 1, 1))
    out = bias.view(1, CO, 1, 1).expand(B, CO, H, W).clone()
    for tap in range(9):
        ky, kx = tap // 3, tap % 3
        out += zp[:, tap * CO:(tap + 1) * CO, ky:ky + H, kx:kx + W]
    ref = F.conv2d(a, w, bias, padding=1)
    assert float((out - ref).abs().max()) < 1e-12

