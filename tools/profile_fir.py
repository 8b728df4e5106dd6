"""CUDA-event timing of the up / down ResBlocks' FIR pass (dsep_fir_resample_f32) at the level-0 shapes."""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402
from diffsep_b200 import ops  # noqa: E402

B, C = int(os.environ.get("DSEP_BENCH_BATCH", "32")), 128
reps = int(os.environ.get("DSEP_REPS", "20"))
dev = "cuda"
for mode, H in ((2, 256), (1, 128)):
    W = H
    Ho, Wo = (2 * H, 2 * W) if mode == 1 else (H // 2, W // 2)
    x = torch.randn(B, H, W, C, device=dev)
    st = torch.empty(B, C, 2, dtype=torch.float64, device=dev)
    ops.channel_stats(x, C, B, H * W, st)
    gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    af, y = torch.empty(B, Ho, Wo, C, device=dev), torch.empty(B, Ho, Wo, C, device=dev)
    run = lambda: ops.fir_resample_f32(x, B, H, W, C, mode, 32, st, gamma, beta, 1e-6, af, y=y)
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
    gb = (x.numel() + af.numel() + y.numel()) * 4 / 1e9
    print(f"fir_resample_f32 mode={mode} {H}x{W}x{C} B={B}: {ms:.3f} ms, {gb / ms * 1e3:.0f} GB/s algorithmic "
          f"({gb:.2f} GB)", flush=True)
