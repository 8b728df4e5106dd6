"""In-graph time of sub-ranges of the backbone's launch plan (the ncu launch list times every kernel cold and alone,
which overstates the small, latency-bound launches of the 4x4 ... 32x32 levels): a range of plan steps is captured
into its own CUDA graph and replayed.

    python tools/profile_levels.py [a:b ...]      (ranges of LAUNCH indices of tools/profile_eval.py's launch list)
"""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests" / "golden"))
os.environ["DSEP_CUDA_GRAPH"] = "0"

import torch  # noqa: E402

import cases  # noqa: E402
from diffsep_b200.score_model import ScoreModelNCSNpp  # noqa: E402
from diffsep_b200 import synthetic as ow  # noqa: E402

B = int(os.environ.get("DSEP_BENCH_BATCH", "32"))
NF = int(os.environ.get("DSEP_NF", "128"))
T = int(os.environ.get("DSEP_T", "32000"))
reps = int(os.environ.get("DSEP_REPS", "20"))
FIRST = int(os.environ.get("DSEP_FIRST_PLAN_LAUNCH", "8"))     # launch index of plan step 0 in the launch list
sm = ScoreModelNCSNpp(num_sources=2, backbone_args=dict(nf=NF), passes=2,
                      state_dict=ow.make_score_model_state_dict(nf=NF, seed=0))
xt, t, mix = (v.cuda() for v in cases.score_inputs(B, T, seed=3))
for _ in range(2):
    y = sm(xt, t, mix)
torch.cuda.synchronize()
(plan,) = sm.backbone._plans.values()
steps = plan.steps
print(f"{len(steps)} plan steps", flush=True)
ranges = [tuple(int(v) for v in a.split(":")) for a in sys.argv[1:]] or [(FIRST, FIRST + len(steps))]
s = torch.cuda.Stream()
for a, b in ranges:
    sub = steps[a - FIRST:b - FIRST]
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        for st in sub:
            st()
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            for st in sub:
                st()
        for _ in range(3):
            g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(reps):
            g.replay()
        e1.record(s)
    torch.cuda.synchronize()
    print(f"launches [{a}:{b}) = {len(sub)} steps: {e0.elapsed_time(e1) / reps * 1e3:.1f} us per replay", flush=True)
