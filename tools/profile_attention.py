"""Times dsep_attention alone (CUDA events): DSEP_ATTN_TC=1 the tcgen05 flash-style kernel, 0 the fp32 CUDA-core one.

    python tools/profile_attention.py [B S C]
"""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

from diffsep_b200 import ops  # noqa: E402

B, S, C = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (8, 1920, 256)
qkv = torch.randn(B, S, 3 * C, device="cuda")
o = ops.Split.empty((B, S, C), "cuda")
run = lambda: ops.attention(qkv, B, S, C, C ** -0.5, o)
for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
fl = 4.0 * B * S * S * C
print(f"attention B={B} S={S} C={C} tc={os.environ.get('DSEP_ATTN_TC', '1')}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s "
      f"algorithmic (2 contractions)", flush=True)
