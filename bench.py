#!/usr/bin/env python
"""DiffSep reverse-diffusion throughput: separated utterances / second.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one complete predictor-corrector sampling run (normalize -> prior -> N x (ald2
corrector, reverse-diffusion predictor) -> denoised estimate) over one batch of synthetic
2-speaker mixtures: BASELINE.json configs[1] — batch 32 x 4 s @ 8 kHz, N=30, 1 corrector step,
snr 0.5, NCSN++ nf=128 — i.e. 60 score-network evaluations of [32, 6, 256, 256] spectrograms.
With N GPUs every rank runs its own batch of 32 (weak scaling; utterances are independent, the
only collective is the final all-gather of the estimates).

Prints ONE JSON line (rank 0).  ``value`` is measured with the mixtures resident in HBM; ``e2e``
is the same job through the public API from pinned HOST buffers (H2D + D2H inside the timed
region).  ``roofline`` is the dominant kernel (level-0 3x3 conv on tcgen05) timed alone with CUDA
events against MEASURED_PEAKS.json; ``cpu_baseline`` is the CPU oracle (a restatement of the
reference's PyTorch path) on this box's host cores over a bounded sample.

``--impl reference`` times that CPU path alone (rank 0 only), on the same metric / config.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests" / "golden"))

import torch  # noqa: E402

FS, SECONDS, N_STEPS, CORR_STEPS, SNR, NF = 8000, 4, 30, 1, 0.5, 128
B_DEFAULT, GFLOP_PER_EVAL, PRIOR, WORKLOAD = 32, 532.891, False, "configs[1]"   # GFLOP/sample/eval: SURVEY.md §8d
METRIC = "separated utterances/sec (4 s, 8 kHz, 2-spk, N=30 PC steps)"
# The driver's contract is configs[1] (the default).  The other single-GPU configurations of BASELINE.json can
# be timed for the record with DSEP_BENCH_WORKLOAD; they print the same line with their own `config`.
_W = os.environ.get("DSEP_BENCH_WORKLOAD", "")
if _W == "configs[3]":      # enhancement: 16 kHz, PriorMixSDE, 4 s -> [B,6,256,512]
    FS, B_DEFAULT, GFLOP_PER_EVAL, PRIOR, WORKLOAD = 16000, 16, 1066.173, True, _W
    METRIC = "enhanced utterances/sec (4 s, 16 kHz, PriorMixSDE, N=30 PC steps)"
elif _W == "configs[4]":    # long form: 30 s, N=50, 2 corrector steps -> [B,6,256,1920]
    SECONDS, N_STEPS, CORR_STEPS, B_DEFAULT, GFLOP_PER_EVAL, WORKLOAD = 30, 50, 2, 8, 4006.432, _W
    METRIC = "separated utterances/sec (30 s, 8 kHz, 2-spk, N=50 PC steps, 2 corrector steps)"
T = FS * SECONDS
B_PER_GPU = int(os.environ.get("DSEP_BENCH_BATCH", str(B_DEFAULT)))
NFE = N_STEPS * (CORR_STEPS + 1)


def config_dict(n_gpus):
    return {
        "workload": f"{WORKLOAD}: batch={B_PER_GPU} x {SECONDS} s {FS // 1000} kHz mixtures per GPU, N={N_STEPS}, "
                    f"{CORR_STEPS} corrector step(s), snr=0.5, NCSN++ nf=128 ({NFE} score evaluations of "
                    f"[{B_PER_GPU},6,256,{-(-(1 + (T + 382) // 128) // 64) * 64}])",
        "global_batch": B_PER_GPU * n_gpus, "samples": T, "n_fft": 510, "hop": 128, "N": N_STEPS,
        "corrector_steps": CORR_STEPS, "snr": SNR, "nf": NF, "sde": "PriorMixSDE" if PRIOR else "MixSDE",
        "predictor": "reverse_diffusion",
        "corrector": "ald2", "passes": int(os.environ.get("DSEP_PASSES", "2")),
        "l2": "working set (>4 GB of activations per evaluation) exceeds the 126 MB L2; no flush needed",
        "parallelism": f"dp{n_gpus} (utterance sharding, one all-gather of outputs)",
    }


def synthetic_batch(first, count):
    import cases
    return cases.batch_mix(count, T, first=first)        # [count, 1, T] fp32, CPU generator


# ------------------------------------------------------------------------------ CPU oracle arm
def cpu_sample(n_evals_N, threads):
    """Times the CPU oracle on 1 utterance for N=n_evals_N steps, 1 corrector step
    (2 n network evaluations); returns (seconds, utt/s scaled to the full 60-evaluation job)."""
    from oracle import score_ref as sr, sde_ref as sd, weights as ow
    import cases
    torch.set_num_threads(threads)
    params = ow.make_backbone_params(nf=NF, seed=0)
    mix, _, _ = sd.normalize_batch(synthetic_batch(0, 1))
    p = sd.MixSDEParams(N=n_evals_N, prior=PRIOR)
    noises = cases.sampler_noises(1, T, n_evals_N, CORR_STEPS)

    def score_fn(x, t, m):
        with torch.no_grad():
            return sr.score_forward(params, x, t, m)
    t0 = time.perf_counter()
    out, nfe = sd.pc_sampler(p, score_fn, mix, noises, eps=0.03, snr=SNR, corrector_steps=CORR_STEPS)
    dt = time.perf_counter() - t0
    assert torch.isfinite(out).all()
    return dt, 1.0 / (dt * NFE / nfe), nfe


REFERENCE_BUDGET_S = float(os.environ.get("DSEP_REF_BUDGET_S", "240"))


def run_reference(args):
    """The reference arm: the CPU restatement of the reference's own path (the oracle; the Python reference cannot
    travel to the GPU box) on all host cores.  Each step is 1 utterance x n PC steps (corrector + predictor = 2n
    network evaluations) with n chosen from a calibration evaluation so that warm-up + timed steps fit in
    DSEP_REF_BUDGET_S (240 s); when the budget allows n = 30 a step IS the whole job of one utterance (no
    extrapolation), otherwise the value is scaled by 30 / n (the job is n-linear: 60 identical evaluations)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    t_cal, _, nfe_cal = cpu_sample(1, threads)                  # calibration (also warms oneDNN / the allocator up)
    per_eval = t_cal / nfe_cal
    n_pc = int(REFERENCE_BUDGET_S / (max(args.steps + args.warmup, 1) * (CORR_STEPS + 1) * per_eval))
    n_pc = max(1, min(N_STEPS, n_pc))
    for _ in range(args.warmup):
        cpu_sample(n_pc, threads)
    t0 = time.perf_counter()
    nfe = 0
    for _ in range(args.steps):
        _, _, n = cpu_sample(n_pc, threads)
        nfe += n
    dt = time.perf_counter() - t0
    utt_s = args.steps / (dt * NFE / (nfe / args.steps)) if dt > 0 else 0.0
    scale = NFE * args.steps / nfe
    sample = (f"per step: 1 utterance x {nfe // args.steps} of the job's {NFE} score evaluations "
              f"(N={n_pc} PC steps incl. corrector) through the CPU oracle"
              + (", the whole job: no extrapolation" if scale == 1 else f", scaled x{scale:g}")
              + f"; {per_eval:.2f} s per evaluation at calibration")
    line = {
        "impl": "reference", "metric": METRIC, "value": utt_s, "unit": "utt/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args.gpus),
        "cpu_baseline": {"value": utt_s, "unit": "utt/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": utt_s, "unit": "utt/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "200"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = float(r[2])
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7),
                              ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        med = sm[len(sm) // 2] if sm else None
        return {"sm_mhz": med, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------ GPU arm
def run_gpu(args):
    import torch.distributed as dist
    from diffsep_b200 import _lib, ops
    from diffsep_b200.pl_model import DEFAULT_CONFIG, DiffSepModel
    from diffsep_b200 import synthetic as ow

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ops.require_device()
    passes = int(os.environ.get("DSEP_PASSES", "2"))
    if passes not in (1, 2, 3):
        raise SystemExit("DSEP_PASSES must be 2 (default: fp16 hi*hi + e4m3 corrections), 3 (three fp16 products) "
                         "or 1 (TF32-grade, not a parity mode)")

    import copy
    cfg = copy.deepcopy(DEFAULT_CONFIG)
    cfg["model"]["fs"] = FS
    if PRIOR:
        cfg["model"]["sde"] = {"_target_": "sdes.sdes.PriorMixSDE", "ndim": 2, "d_lambda": 2.0, "sigma_min": 0.05,
                               "sigma_max": 0.5, "N": 30, "avg_len": 510}
    model = DiffSepModel(cfg, device=dev, passes=passes,
                         score_state_dict=ow.make_score_model_state_dict(nf=NF, seed=0))
    B = B_PER_GPU
    host_mix = synthetic_batch(rank * B, B).pin_memory()
    dev_mix = host_mix.to(dev)
    host_out = torch.empty(B, 2, T).pin_memory()
    gathered = torch.empty(world * B, 2, T, device=dev) if world > 1 else None
    torch.manual_seed(1234 + rank)
    kw = dict(N=N_STEPS, corrector_steps=CORR_STEPS, snr=SNR, denoise=True)

    def job(mix_dev):
        """the public path a user calls: normalize -> sampler -> per-source rescale -> gather"""
        (mix_n, _), mean, std = model.normalize_batch((mix_dev, None))
        est, nfe = model.get_pc_sampler("reverse_diffusion", "ald2", mix_n, **kw)()
        out = torch.empty_like(est)
        ops.scale_output(mix_dev, est, B, 2, T, out)      # against the raw mixture, separate.py:97
        if world > 1:
            dist.all_gather_into_tensor(gathered, out)
        return out, nfe

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    for _ in range(args.warmup):
        out, nfe = job(dev_mix)
    barrier()
    assert nfe == NFE and bool(torch.isfinite(out).all())

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    calls0 = _lib.N_CALLS
    ms = timed(lambda: job(dev_mix), args.steps)
    launches = _lib.N_CALLS - calls0
    clock_info = clocks.stop() if rank == 0 else None

    def job_e2e():
        d = host_mix.to(dev, non_blocking=True)
        o, _ = job(d)
        host_out.copy_(o, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    job_e2e()
    ms_e2e = timed(job_e2e, args.steps)

    total_utts = world * B * args.steps
    value = total_utts / (ms / 1e3)
    e2e = total_utts / (ms_e2e / 1e3)

    line = None
    if rank == 0:
        peaks = {}
        pk = ROOT / "MEASURED_PEAKS.json"
        if pk.exists():
            peaks = json.loads(pk.read_text())
        roof = dominant_kernel_roofline(model, dev, peaks, passes)
        step_tflops = GFLOP_PER_EVAL * 1e-3 * NFE * B * args.steps / (ms / 1e3)
        sustained = peaks.get("bf16_tflops_sustained", 1400.0)
        roof["step"] = {"achieved": step_tflops, "peak": sustained, "unit": "TFLOP/s",
                        "frac": step_tflops / sustained,
                        "note": f"whole job: algorithmic FLOPs ({GFLOP_PER_EVAL} GFLOP/sample/eval x {NFE}) / wall, "
                                "per GPU, vs sustained measured bf16 peak"}
        threads = os.cpu_count() or 1
        # the CPU baseline is taken on rank 0 at N=1 only: under torchrun the other ranks' NCCL barrier
        # spin-waits would share the host cores with it (measured: 0.027 -> 0.002 utt/s at N=2)
        cpu_base = None
        if world == 1:
            cpu_dt, cpu_utt_s, cpu_nfe = cpu_sample(3 if WORKLOAD == "configs[1]" else 1, threads)
            cpu_base = {"value": cpu_utt_s, "unit": "utt/s", "cores": threads, "kind": "port",
                        "sample": f"1 utterance x {cpu_nfe} of {NFE} score evaluations through "
                                  f"the CPU oracle in {cpu_dt:.1f} s, scaled x{NFE / cpu_nfe:g}"}
        line = {
            "metric": METRIC, "value": value, "unit": "utt/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": {3: "f32 (3 x fp16 tensor-core passes, fp32 accumulate)",
                                           2: "f32 (fp16 hi*hi + e4m3 correction products, fp32 accumulate)",
                                           1: "fp16 operands (TF32-grade), fp32 accumulate"}[passes],
            "data": "synthetic", "config": config_dict(world),
            "e2e": {"value": e2e, "unit": "utt/s", "h2d_bytes_per_step": host_mix.numel() * 4,
                    "d2h_bytes_per_step": host_out.numel() * 4, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "clocks": clock_info, "roofline": roof,
        }
        if cpu_base is not None:
            line["cpu_baseline"] = cpu_base
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line), flush=True)


def dominant_kernel_roofline(model, dev, peaks, passes):
    """Times the level-0 ResBlock 3x3 conv (128 -> 128 at 256x256, batch 32: 57.6 % of the network's
    FLOPs run at this shape) alone, exactly as the network launches it — GroupNorm-apply + SiLU + fp16
    split prologue in-kernel, FiLM bias and next-GroupNorm statistics in the epilogue: 20 launches
    bracketed by CUDA events on the launching stream."""
    from diffsep_b200 import ops
    bb = model.score_model.backbone
    rb = bb.down[0]["blocks"][0]
    B, H, W, Cc = B_PER_GPU, 256, 256, NF
    x = torch.randn(B, H, W, Cc, device=dev)
    sc = torch.ones(B, Cc, device=dev)
    sh = torch.zeros(B, Cc, device=dev)
    film = torch.zeros(B, Cc, device=dev)
    stats = torch.zeros(B, Cc, 2, dtype=torch.float64, device=dev)
    out = torch.empty(B, H, W, Cc, device=dev)
    cw = rb["conv0"]
    if bb.fuse and passes == 2:
        run = lambda: ops.conv2d_fused(B, H, W, Cc, cw.planes8(), cw.cout_pad, 3, out, Cc, x0=x, C0=Cc, sc=sc, sh=sh,
                                       act=1, bias=cw.bias, film=film, film_stride=Cc, acc_scale=cw.acc_scale,
                                       stats=stats, passes=2, corr_rel=cw.corr_rel, a8_exp=cw.A8_EXP)
        variant = ("fp32 input, GN+SiLU+split prologue in-kernel (fp16 hi + e4m3 correction planes), FiLM + GN "
                   "statistics in the epilogue")
    elif bb.fuse:
        run = lambda: ops.conv2d_fused(B, H, W, Cc, cw.planes, cw.cout_pad, 3, out, Cc, x0=x, C0=Cc, sc=sc, sh=sh,
                                       act=1, bias=cw.bias, film=film, film_stride=Cc, acc_scale=cw.acc_scale,
                                       stats=stats, passes=passes)
        variant = "fp32 input, GN+SiLU+split prologue in-kernel, FiLM + GN statistics in the epilogue"
    else:
        a = ops.Split.empty((B, H, W, Cc), dev)
        ops.split_f16(x, a)
        run = lambda: ops.conv2d_tc(a, B, H, W, Cc, cw.planes, cw.cout_pad, 3, out, Cc, bias=cw.bias, film=film,
                                    film_stride=Cc, acc_scale=cw.acc_scale, stats=stats, passes=passes)
        variant = "fp16 hi/lo operand planes by TMA, FiLM + GN statistics in the epilogue"
    def time_it(fn, reps=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    from diffsep_b200 import _lib
    n_wide = _lib.load().dsep_conv_wide_launches()
    ms = time_it(run)
    # which kernel the dispatch picked for this shape (conv_wide.cu: 8 x 32 pixel tiles on the MMA's N side)
    kname = ("conv_wide_kernel" if _lib.load().dsep_conv_wide_launches() > n_wide
             else ("conv_fused_kernel<128>" if bb.fuse else "conv_tc_kernel<128, halo>"))
    # for reference: the same convolution fed with ready-made operand planes (no prologue, no statistics)
    ap = ops.Split.empty((B, H, W, Cc), dev)
    ops.split_f16(x, ap)
    plane_passes = 3 if passes == 2 else passes       # operand planes carry fp16 (hi, lo) only
    ms_planes = time_it(lambda: ops.conv2d_tc(ap, B, H, W, Cc, cw.planes, cw.cout_pad, 3, out, Cc, bias=cw.bias,
                                              acc_scale=cw.acc_scale, passes=plane_passes))
    flops = 2.0 * B * H * W * 9 * Cc * Cc
    achieved = flops / (ms * 1e-3) / 1e12
    peak = peaks.get("bf16_tflops", 1590.0)
    traffic = None
    tf = ROOT / "profiles" / "conv_traffic.json"      # dram bytes per launch from the committed ncu --set full capture
    if tf.exists() and B == 32:
        traffic = json.loads(tf.read_text()).get("dram_bytes_per_launch")
    return {"bound": "tensor", "kernel": "%s (3x3, 128->128, 256x256, batch %d): %s" % (kname, B, variant),
            "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
            "peak_source": "MEASURED_PEAKS.json bf16_tflops (burst)" if "bf16_tflops" in peaks else "fallback",
            "ms_per_launch": ms, "algorithmic_gflop_per_launch": flops / 1e9, "mma_passes": passes,
            # tensor-core time units per MAC: 3 fp16 products, or 1 fp16 + 2 e4m3 products at twice the rate
            "issued_frac": passes * achieved / peak, "traffic": traffic,
            "bare_conv": {"ms_per_launch": ms_planes, "achieved": flops / (ms_planes * 1e-3) / 1e12,
                          "frac": flops / (ms_planes * 1e-3) / 1e12 / peak,
                          "note": "same conv with precomputed operand planes, no prologue / FiLM / statistics"},
            "algorithmic_bytes_per_launch": 2 * 4.0 * B * H * W * Cc}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
