"""Runs the dominant kernel alone — the level-0 ResBlock 3x3 convolution (128 -> 128 channels,
256x256 map, batch 32) — for `ncu --set full`, and prints its CUDA-event time.

    ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 3 -c 2 \
        -o gpurun_out/conv python tools/profile_conv.py
"""
import math
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

from diffsep_b200 import ops  # noqa: E402
from diffsep_b200.backbone import ConvWeight  # noqa: E402

B = int(os.environ.get("DSEP_BENCH_BATCH", "32"))
H = W = int(os.environ.get("DSEP_HW", "256"))
CIN = int(os.environ.get("DSEP_CIN", "128"))
COUT = int(os.environ.get("DSEP_COUT", "128"))
K = int(os.environ.get("DSEP_K", "3"))
passes = int(os.environ.get("DSEP_PASSES", "2"))
with_res = int(os.environ.get("DSEP_RES", "0"))
reps = int(os.environ.get("DSEP_REPS", "5"))
dev = "cuda"
g = torch.Generator().manual_seed(0)
w = torch.randn(COUT, CIN, K, K, generator=g) / math.sqrt(CIN * K * K)
short = int(os.environ.get("DSEP_SHORT", "0"))      # channels of a fused fp16 1x1 shortcut on a raw fp32 operand
skw = {}
if short:
    w2 = torch.randn(COUT, short, 1, 1, generator=g) / math.sqrt(short)
    cw = ConvWeight(w, torch.zeros(COUT), dev, shortcut=(w2, None))
    xs = torch.randn(B, H, W, short, device=dev)
    skw = dict(s0=xs, S0=short, Cin2=cw.cin2_pad, w2=cw.planes2)
else:
    cw = ConvWeight(w, torch.zeros(COUT), dev)
x = torch.randn(B, H, W, CIN, device=dev)
a = ops.Split.empty((B, H, W, CIN), dev)
ops.split_f16(x, a)
out = torch.empty(B, H, W, COUT, device=dev)
res = torch.randn(B, H, W, COUT, device=dev) if with_res else None
with_stats = int(os.environ.get("DSEP_STATS", "0"))
fused_in = int(os.environ.get("DSEP_FUSEDIN", "0"))
stats = torch.zeros(B, COUT, 2, dtype=torch.float64, device=dev) if with_stats else None
sc = torch.ones(B, CIN, device=dev)
sh = torch.zeros(B, CIN, device=dev)
if fused_in and passes == 2:      # experimental e4m3-correction mode (DSEP_LIB = a -DDSEP_FP8_CORR=1 build)
    run = lambda: ops.conv2d_fused(B, H, W, CIN, cw.planes8(), cw.cout_pad, K, out, COUT, x0=x, C0=CIN, sc=sc, sh=sh,
                                   act=1, bias=cw.bias, residual=res, scale=0.7071 if with_res else 1.0,
                                   acc_scale=cw.acc_scale, stats=stats, passes=2, corr_rel=cw.corr_rel,
                                   a8_exp=cw.A8_EXP, **skw)
elif fused_in:
    run = lambda: ops.conv2d_fused(B, H, W, CIN, cw.planes, cw.cout_pad, K, out, COUT, x0=x, C0=CIN, sc=sc, sh=sh,
                                   act=1, bias=cw.bias, residual=res, scale=0.7071 if with_res else 1.0,
                                   acc_scale=cw.acc_scale, stats=stats, passes=passes)
else:
    run = lambda: ops.conv2d_tc(a, B, H, W, CIN, cw.planes, cw.cout_pad, K, out, COUT, bias=cw.bias, residual=res,
                                scale=0.7071 if with_res else 1.0, acc_scale=cw.acc_scale, passes=passes,
                                stats=stats)
for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
fl = 2.0 * B * H * W * (K * K * CIN + short) * COUT
print(f"conv {K}x{K} {CIN}->{COUT} {H}x{W} B={B} passes={passes} res={with_res} short={short} stats={with_stats} "
      f"fused_in={fused_in}: {ms:.3f} ms  "
      f"{fl / ms / 1e9:.1f} TFLOP/s algorithmic, {passes * fl / ms / 1e9:.1f} TFLOP/s issued", flush=True)
