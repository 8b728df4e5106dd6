"""Generate the golden vectors under tests/golden/ by running the REAL reference
(/root/reference, read-only) on seeded inputs.  Runs only in the build container -- the
GPU box has no /root/reference; tests there read the committed .npz/.json files.

    TORCH_CUDA_ARCH_LIST=10.0 TORCH_EXTENSIONS_DIR=/tmp/torch_ext python tests/golden/make_golden.py

The reference is imported with two stub modules: ``pytorch_lightning``
(only pulled in by ``utils/checkpoint_symlink.py:5``) and ``hydra.utils.instantiate``
(``models/score_models.py:7,27``).  Importing ``models.ncsnpp`` JIT-builds the reference's
two CUDA extensions (~2 min cold) even for CPU use.
"""
from __future__ import annotations

import importlib
import json
import os
import sys
import types
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(HERE))

import cases  # noqa: E402
from oracle import weights as ow  # noqa: E402

REF = os.environ.get("DSEP_REFERENCE", "/root/reference")


def import_reference():
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    os.environ.setdefault("TORCH_EXTENSIONS_DIR", "/tmp/torch_ext")
    sys.path.insert(0, REF)

    def stub(name, **a):
        m = types.ModuleType(name)
        m.__dict__.update(a)
        sys.modules[name] = m
        return m

    def instantiate(cfg, *args, **kw):
        cfg = dict(cfg)
        mod, _, cls = cfg.pop("_target_").rpartition(".")
        return getattr(importlib.import_module(mod), cls)(*args, **{**cfg, **kw})

    stub("pytorch_lightning", LightningModule=torch.nn.Module, LightningDataModule=object)
    stub("hydra")
    stub("hydra.utils", instantiate=instantiate, to_absolute_path=lambda p: p)
    import sdes  # noqa
    from sdes.sdes import MixSDE, PriorMixSDE
    from models.score_models import ScoreModelNCSNpp
    from models.ncsnpp_utils import layerspp, up_or_down_sampling
    return dict(sdes=sdes, MixSDE=MixSDE, PriorMixSDE=PriorMixSDE,
                ScoreModelNCSNpp=ScoreModelNCSNpp, layerspp=layerspp,
                uds=up_or_down_sampling)


def make_score_model(R, nf, seed=0, spec_factor=0.15):
    sm = R["ScoreModelNCSNpp"](
        num_sources=2,
        stft_args=dict(n_fft=510, hop_length=128, center=True, pad_mode="constant"),
        backbone_args=dict(_target_="models.ncsnpp.NCSNpp", nf=nf),
        spec_abs_exponent=0.5, spec_factor=spec_factor).eval()
    sd = ow.make_score_model_state_dict(nf=nf, seed=seed)
    missing, unexpected = sm.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    return sm


class NoiseInjector:
    """Replaces torch.randn_like by popping pre-drawn tensors."""

    def __init__(self, noises):
        self.noises = list(noises)
        self._orig = torch.randn_like

    def __enter__(self):
        def fake(x, *a, **k):
            z = self.noises.pop(0)
            assert z.shape == x.shape, (z.shape, x.shape)
            return z.to(x.dtype)
        torch.randn_like = fake
        return self

    def __exit__(self, *exc):
        torch.randn_like = self._orig


def save(name, **arrays):
    out = HERE / name
    np.savez_compressed(out, **{k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v))
                                for k, v in arrays.items()})
    print(f"wrote {out} ({out.stat().st_size / 1024:.1f} KiB)")


def main():
    torch.set_num_threads(os.cpu_count())
    torch.manual_seed(0)
    R = import_reference()

    # 1. structure pins: reference parameter names/shapes (parameters() order)
    for nf in (128, 64):
        sm = R["ScoreModelNCSNpp"](
            num_sources=2, stft_args=dict(n_fft=510, hop_length=128, center=True, pad_mode="constant"),
            backbone_args=dict(_target_="models.ncsnpp.NCSNpp", nf=nf))
        names = [[k, list(v.shape)] for k, v in sm.backbone.named_parameters()]
        sd_names = [[k, list(v.shape)] for k, v in sm.state_dict().items()]
        with open(HERE / f"structure_nf{nf}.json", "w") as f:
            json.dump({"parameters": names, "state_dict": sd_names,
                       "n_params": int(sum(p.numel() for p in sm.backbone.parameters())),
                       "n_modules": len(sm.backbone.all_modules)}, f)
        print(f"nf={nf}: {len(names)} parameters, {len(sd_names)} state_dict entries")

    # 2. FIR resampling (reference upsample_2d / downsample_2d -> upfirdn2d_native on CPU)
    g = cases.gen(11)
    xf = torch.randn(2, 3, 8, 12, generator=g)
    save("fir.npz", x=xf,
         down=R["uds"].downsample_2d(xf, (1, 3, 3, 1), factor=2),
         up=R["uds"].upsample_2d(xf, (1, 3, 3, 1), factor=2))

    # 3. ResBlock / AttnBlock / Combine modules with oracle-generated weights
    L = R["layerspp"]
    act = torch.nn.SiLU()
    blk = {}
    for name, kw, shape in cases.RESBLOCK_CASES:
        mod = L.ResnetBlockBigGANpp(act=act, in_ch=kw["cin"], out_ch=kw["cout"], temb_dim=32,
                                    up=kw["up"], down=kw["down"], dropout=0.0, fir=True,
                                    fir_kernel=[1, 3, 3, 1], skip_rescale=True, init_scale=0.0).eval()
        gg = cases.gen(100 + len(blk))
        sd = {k: torch.randn(v.shape, generator=gg) * (0.2 if v.ndim > 1 else 0.5) + (1.0 if "GroupNorm" in k and k.endswith("weight") else 0.0)
              for k, v in mod.state_dict().items()}
        mod.load_state_dict(sd)
        x = torch.randn(shape, generator=gg)
        temb = torch.randn(shape[0], 32, generator=gg)
        with torch.no_grad():
            y = mod(x, temb)
        blk[f"{name}.x"] = x
        blk[f"{name}.temb"] = temb
        blk[f"{name}.y"] = y
        for k, v in sd.items():
            blk[f"{name}.p.{k}"] = v
    mod = L.AttnBlockpp(channels=16, skip_rescale=True, init_scale=0.0).eval()
    gg = cases.gen(200)
    sd = {k: torch.randn(v.shape, generator=gg) * 0.3 + (1.0 if k == "GroupNorm_0.weight" else 0.0)
          for k, v in mod.state_dict().items()}
    mod.load_state_dict(sd)
    x = torch.randn(2, 16, 4, 6, generator=gg)
    with torch.no_grad():
        blk["attn.y"] = mod(x)
    blk["attn.x"] = x
    for k, v in sd.items():
        blk[f"attn.p.{k}"] = v
    save("blocks.npz", **blk)

    # 4. pre_process / post_process of the score model (STFT-510 wrapper)
    sm32 = make_score_model(R, nf=32)
    T = 2048
    xt, t, mix = cases.score_inputs(2, T, seed=7)
    with torch.no_grad():
        spec, n_samples, n_pad = sm32.pre_process(torch.cat((xt, mix), dim=1))
        gg = cases.gen(8)
        net_out = torch.randn(2, 4, 256, spec.shape[-1], generator=gg) * 0.2
        wav = sm32.post_process(net_out, n_samples, n_pad)
    save("stft.npz", spec=spec[..., : spec.shape[-1] - n_pad], n_pad=n_pad, wav=wav)

    # 5. full score model, nf=32 (fast) and nf=128 (the benchmarked width)
    with torch.no_grad():
        y32 = sm32(xt, t, mix)
    save("score_nf32.npz", y=y32)
    sm128 = make_score_model(R, nf=128)
    T128 = 7680
    xt1, t1, mix1 = cases.score_inputs(1, T128, seed=9)
    with torch.no_grad():
        y128 = sm128(xt1, t1, mix1)
        spec1, _, _ = sm128.pre_process(torch.cat((xt1, mix1), dim=1))
        b128 = sm128.backbone(spec1, t1)
    save("score_nf128.npz", y=y128, backbone_out=b128)

    # 6. samplers: analytic score (MixSDE, PriorMixSDE, schedules) and network (nf=32)
    sam = {}
    Ts = 1024
    mixs = cases.batch_mix(2, Ts)
    mixs = (mixs - mixs.mean(dim=(1, 2), keepdim=True)) / mixs.std(dim=(1, 2), keepdim=True)
    for tag, cls, kw in (("mix", R["MixSDE"], {}), ("priormix", R["PriorMixSDE"], {})):
        for cs in (0, 1, 2):
            sde = cls(ndim=2, d_lambda=2.0, sigma_min=0.05, sigma_max=0.5, N=30, **kw)
            noises = cases.sampler_noises(2, Ts, 30, cs)
            with NoiseInjector(noises):
                out, nfe = R["sdes"].get_pc_sampler(
                    "reverse_diffusion", "ald2", sde=sde, score_fn=cases.analytic_score, y=mixs,
                    eps=0.03, snr=0.5, corrector_steps=cs, denoise=True)()
            sam[f"{tag}.cs{cs}"] = out.contiguous()
            assert nfe == 30 * (cs + 1)
    for sched in ("linear", "log", "revlog"):
        sde = R["MixSDE"](ndim=2, d_lambda=2.0, sigma_min=0.05, sigma_max=0.5, N=10)
        noises = cases.sampler_noises(2, Ts, 10, 1)
        with NoiseInjector(noises):
            out, nfe = R["sdes"].get_pc_scheduled_sampler(
                "reverse_diffusion", "ald2", sde=sde, score_fn=cases.analytic_score, y=mixs,
                eps=0.03, snr=0.5, corrector_steps=1, denoise=False, schedule=sched)()
        sam[f"sched.{sched}"] = out.contiguous()
    # network-driven sampler, nf=32, N=3, cs=1, denoise True, with intermediates
    mixn = cases.batch_mix(1, T)
    mixn = (mixn - mixn.mean(dim=(1, 2), keepdim=True)) / mixn.std(dim=(1, 2), keepdim=True)
    sde = R["MixSDE"](ndim=2, d_lambda=2.0, sigma_min=0.05, sigma_max=0.5, N=3)
    noises = cases.sampler_noises(1, T, 3, 1)
    with NoiseInjector(noises):
        out, nfe, im = R["sdes"].get_pc_sampler(
            "reverse_diffusion", "ald2", sde=sde, score_fn=sm32, y=mixn, eps=0.03, snr=0.5,
            corrector_steps=1, denoise=True, intermediate=True)()
    sam["net32.out"] = out
    sam["net32.im0"] = im[0][0]
    save("sampler.npz", **sam)

    # 7. normalize_batch (pl_model.py:81-88) / scale_output (separate.py:73-78): those modules
    # cannot be imported (lightning/hydra/hf_hub absent), so the two free functions are
    # extracted from the reference source by AST and executed as-is.
    import ast

    def extract(path, fname):
        tree = ast.parse(Path(path).read_text())
        fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == fname)
        ns = {"torch": torch}
        exec(compile(ast.Module([fn], []), path, "exec"), ns)
        return ns[fname]

    normalize_batch = extract(f"{REF}/pl_model.py", "normalize_batch")
    scale_output = extract(f"{REF}/separate.py", "scale_output")
    m = cases.batch_mix(3, 512) * 3.0 + 0.2
    sep = torch.randn(3, 2, 512, generator=cases.gen(5))
    (norm, _), mean, std = normalize_batch((m, None))
    save("misc.npz", norm=norm, mean=mean, std=std, scaled=scale_output(m, sep))


if __name__ == "__main__":
    main()
