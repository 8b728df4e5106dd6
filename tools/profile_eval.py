"""One score-network evaluation at the benchmark shape, bracketed by cudaProfilerStart/Stop so that
`ncu --profile-from-start off` captures exactly the launches of one evaluation (after a warm-up).

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches.csv python tools/profile_eval.py
"""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests" / "golden"))

import torch  # noqa: E402

import cases  # noqa: E402
from diffsep_b200.score_model import ScoreModelNCSNpp  # noqa: E402
from diffsep_b200 import synthetic as ow  # noqa: E402

B = int(os.environ.get("DSEP_BENCH_BATCH", "32"))
NF = int(os.environ.get("DSEP_NF", "128"))
T = int(os.environ.get("DSEP_T", "32000"))
passes = int(os.environ.get("DSEP_PASSES", "2"))
sm = ScoreModelNCSNpp(num_sources=2, backbone_args=dict(nf=NF), passes=passes,
                      state_dict=ow.make_score_model_state_dict(nf=NF, seed=0))
xt, t, mix = (v.cuda() for v in cases.score_inputs(B, T, seed=3))
for _ in range(2):
    y = sm(xt, t, mix)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.profiler.start()
e0.record()
y = sm(xt, t, mix)
e1.record()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(f"one evaluation B={B} nf={NF} T={T} passes={passes}: {e0.elapsed_time(e1):.2f} ms", flush=True)
