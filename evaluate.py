#!/usr/bin/env python
"""Batch separation of a folder of mixtures with per-utterance timing: the inference loop of the
reference's ``evaluate.py`` (:340-406) / ``evaluate_mp.py`` (:154-326, dataset indices sharded over
devices) without the parts that need datasets or metric packages (SI-SDR / PESQ / STOI are out of
scope here).  Results JSON keeps the reference's timing fields: ``nfe``, ``runtime``, ``len_s``.

    python evaluate.py mixtures/ results/ --model checkpoint.pt [--batch-size 8] [-N 30] [--snr 0.5]
        [--corrector-steps 1] [-s linear|log|revlog] [--save-n K] [--limit M]
    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 evaluate.py ...     # files sharded over GPUs

Differences from the reference loop, on purpose: batches of more than one utterance — by default only utterances
of the SAME length share a batch (``bucket_by_length``), which reproduces the reference's batch-of-one results;
``--pad-batches`` instead pads ragged batches like ``max_collator`` (faster on ragged folders, but the zero padding
enters normalisation, STFT frames and GroupNorm statistics, so short utterances then differ from a batch-of-one
run) — and ``runtime`` is taken with the device synchronised (the reference's timer has no ``cuda.synchronize()``,
evaluate.py:374-376).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))


def summarize(results):
    n = sum(len(r["files"]) for r in results)
    t = sum(r["runtime"] for r in results)
    audio = sum(sum(r["len_s"]) for r in results)
    return {"utterances": n, "runtime_s": t, "audio_s": audio, "utt_per_s": n / t if t else None,
            "real_time_factor": t / audio if audio else None,
            "nfe_per_utt": results[0]["nfe"] if results else None}


def main(argv=None):
    import separate as sep_cli
    from diffsep_b200.data import bucket_by_length, load_wav, max_collator, save_wav, uncollate, wav_length
    from diffsep_b200.shard import shard_bounds
    import torch.distributed as dist

    ap = argparse.ArgumentParser(description="Separate a folder of mixtures in batches and time it")
    ap.add_argument("input_dir", type=Path)
    ap.add_argument("output_dir", type=Path)
    ap.add_argument("--model", type=Path, default=sep_cli.DEFAULT_MODEL)
    ap.add_argument("-d", "--device", type=sep_cli.str_or_int, default=None)
    ap.add_argument("--batch-size", type=int, default=1)
    ap.add_argument("-N", type=int, default=None)
    ap.add_argument("--snr", type=float, default=None)
    ap.add_argument("--corrector-steps", type=int, default=None)
    ap.add_argument("--denoise", type=bool, default=True)
    ap.add_argument("-s", "--schedule", type=str, default=None)
    ap.add_argument("--save-n", type=int, default=None, help="save wavs of the first K batches only (default: all)")
    ap.add_argument("-l", "--limit", type=int, default=None, help="stop after M batches")
    ap.add_argument("--pad-batches", action="store_true",
                    help="batch files in folder order and zero-pad ragged batches (results of padded utterances "
                         "differ from a batch-of-one run); default: only equal-length utterances share a batch")
    args = ap.parse_args(argv)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("No CUDA device: the DiffSep hot path here runs on B200 (sm_100a) only")
    if args.device is None:
        args.device = f"cuda:{local}"
    device = f"cuda:{args.device}" if isinstance(args.device, int) else args.device
    torch.cuda.set_device(device)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(device))

    model, kw = sep_cli.get_model(args)
    fs = model.config.model.fs
    files = sorted(args.input_dir.glob("*.wav"))
    lo, hi = shard_bounds(len(files), rank, world)       # contiguous blocks, like evaluate_mp's task split
    files = files[lo:hi]
    args.output_dir.mkdir(parents=True, exist_ok=True)

    if args.pad_batches:
        batches = [list(range(k, min(k + args.batch_size, len(files)))) for k in range(0, len(files), args.batch_size)]
    else:
        batches = bucket_by_length([wav_length(f) for f in files], args.batch_size)
    results = []
    for bno, idx in enumerate(batches):
        bidx = bno * args.batch_size
        if args.limit is not None and bno >= args.limit:
            break
        names = [files[i] for i in idx]
        wavs = []
        for f in names:
            w, sr = load_wav(f)
            if sr != fs:
                print(f"Skipping check: {f.stem} is {sr} Hz, the model expects {fs} Hz")
            wavs.append(w[:1])
        mix, spans = max_collator(wavs)
        mix = mix.to(device)
        (mix_n, _), *_ = model.normalize_batch((mix, None))
        sampler = model.get_pc_sampler("reverse_diffusion", "ald2", mix_n, **kw)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        est, nfe, *others = sampler()
        torch.cuda.synchronize()
        t_proc = time.perf_counter() - t0
        est = sep_cli.scale_output(mix, est)
        results.append({"batch_idx": lo // max(args.batch_size, 1) + bidx // args.batch_size,
                        "files": [f.name for f in names], "nfe": nfe, "runtime": t_proc,
                        "len_s": [n / fs for _, n in spans]})
        print(f"rank {rank} batch {results[-1]['batch_idx']}: {len(names)} utt, nfe={nfe}, runtime={t_proc:.3f} s")
        if args.save_n is None or bidx // args.batch_size < args.save_n:
            for f, e in zip(names, uncollate(est, spans)):
                for i in range(e.shape[0]):
                    d = args.output_dir / f"s{i}"
                    d.mkdir(parents=True, exist_ok=True)
                    save_wav(d / f.name, e[i:i + 1], fs)

    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, results)
        results = [r for part in gathered for r in part]
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        (args.output_dir / "results.json").write_text(json.dumps(results, indent=2))
        (args.output_dir / "results_summary.json").write_text(json.dumps(summarize(results), indent=2))
        print(json.dumps(summarize(results)))


if __name__ == "__main__":
    main()
