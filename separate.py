#!/usr/bin/env python
"""Separate all the wav files in a folder — same command line as the reference ``separate.py``
(:102-134): ``separate.py input_dir output_dir [--model PATH] [-d DEV] [-N int] [--snr f]
[--corrector-steps int] [--denoise bool] [-s {linear,log,revlog}]`` writing
``output_dir/s{0,1}/<stem>.wav``.

Differences forced by the environment, not by design: wav I/O goes through ``scipy.io.wavfile``
(torchaudio's load/save need torchcodec here), checkpoints are local files (no network for the
HF hub), and there is no CPU fallback — the hot path exists only as sm_100a kernels.  ``--model
synthetic[:nf]`` runs with seeded random weights (pipeline check without a checkpoint).
"""
from __future__ import annotations

import argparse
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

DEFAULT_MODEL = "fakufaku/diffsep"


def str_or_int(x):
    try:
        x = int(x)
    except ValueError:
        pass
    return x


def get_model(args):
    from diffsep_b200.pl_model import DEFAULT_CONFIG, DiffSepModel
    name = str(args.model)
    device = f"cuda:{args.device}" if isinstance(args.device, int) else args.device
    if name.startswith("synthetic"):
        import copy
        from diffsep_b200.synthetic import make_score_model_state_dict
        nf = int(name.split(":")[1]) if ":" in name else 128
        cfg = copy.deepcopy(DEFAULT_CONFIG)
        cfg["model"]["score_model"]["backbone_args"]["nf"] = nf
        model = DiffSepModel(cfg, device=device, score_state_dict=make_score_model_state_dict(nf=nf, seed=0))
    elif Path(name).exists():
        model = DiffSepModel.load_from_checkpoint(name, device=device)
    else:
        raise FileNotFoundError(
            f"'{name}' is not a local checkpoint; downloading '{name}' from the Hugging Face hub needs network "
            "access — fetch checkpoint.pt yourself and pass its path with --model")
    model.eval()
    sk = model.config.model.sampler
    kwargs = {
        "N": sk.N if args.N is None else args.N,
        "denoise": args.denoise,
        "intermediate": False,
        "corrector_steps": sk.corrector_steps if args.corrector_steps is None else args.corrector_steps,
        "snr": sk.snr if args.snr is None else args.snr,
        "schedule": args.schedule,
    }
    return model, kwargs


def scale_output(mix, sep):
    """Project the mixture onto each separated signal (reference separate.py:73-78)."""
    from diffsep_b200 import ops
    B, nsrc, T = sep.shape
    out = torch.empty_like(sep)
    ops.scale_output(mix.contiguous(), sep.contiguous(), B, nsrc, T, out)
    return out


def separate(mix, model, sampler_kwargs, device):
    """mix [C, T] as loaded from the file -> [1, n_src, T] on the CPU (reference separate.py:81-99).  Like the
    reference, the whole waveform is passed on: the model takes single-channel mixtures and a multi-channel file is
    an error there (channel mismatch in the backbone's first conv) and here (ValueError from the score model)."""
    mix = mix.to(device=device, dtype=torch.float32)[None]
    (mix_norm, _), *__ = model.normalize_batch((mix, None))
    sampler = model.get_pc_sampler("reverse_diffusion", "ald2", mix_norm, **sampler_kwargs)
    with torch.no_grad():
        sep, nfe, *_ = sampler()
    sep = scale_output(mix, sep)
    return sep.cpu()


def load_wav(path):
    from diffsep_b200.data import load_wav as _load
    return _load(path)


def save_wav(path, wav, sr):
    from diffsep_b200.data import save_wav as _save
    _save(path, wav, sr)


def main(argv=None):
    parser = argparse.ArgumentParser(description="Separate all the wav files in a specified folder")
    parser.add_argument("input_dir", type=Path, help="Path to the input folder")
    parser.add_argument("output_dir", type=Path, help="Path to the output folder")
    parser.add_argument("--model", type=Path, default=DEFAULT_MODEL, help="Path to model or Huggingface model")
    parser.add_argument("-d", "--device", type=str_or_int, default="cuda:0", help="Device to use (default: cuda:0)")
    parser.add_argument("-N", type=int, default=None, help="Number of steps")
    parser.add_argument("--snr", type=float, default=None, help="Step size of corrector")
    parser.add_argument("--corrector-steps", type=int, default=None, help="Number of corrector steps")
    parser.add_argument("--denoise", type=bool, default=True, help="Use denoising in solver")
    parser.add_argument("-s", "--schedule", type=str, help="Pick a different schedule for the inference")
    args = parser.parse_args(argv)

    if not torch.cuda.is_available():
        raise SystemExit("No CUDA device: this build of the DiffSep hot path runs on B200 (sm_100a) only, "
                         "there is no CPU fallback")
    model, sampler_kwargs = get_model(args)
    model_sr = model.config.model.fs
    device = f"cuda:{args.device}" if isinstance(args.device, int) else args.device

    if not args.output_dir.exists():
        args.output_dir.mkdir(parents=True, exist_ok=True)
    elif args.output_dir.is_file():
        raise ValueError("Output directory is a file")

    for wavpath in sorted(args.input_dir.glob("*.wav")):
        waveform, sr = load_wav(wavpath)
        if sr != model_sr:
            print(f"Skipping {wavpath.stem} due to mismatched sample rate. "
                  f"This model expects {model_sr} Hz, but the file is {sr} Hz.")
        sep = separate(waveform, model, sampler_kwargs, device)
        for i in range(sep.shape[1]):
            spkr_dir = args.output_dir / f"s{i}"
            spkr_dir.mkdir(parents=True, exist_ok=True)
            save_wav(spkr_dir / f"{wavpath.stem}.wav", sep[:, i, :], sr)


if __name__ == "__main__":
    main()
