set -x
python -m pytest tests -q -m gpu 2>&1 | tail -3
python __graft_entry__.py --smoke 2>&1 | tail -2
python bench.py 2>&1 | tail -1 > gpurun_out/bench_r1h.json; cut -c1-200 gpurun_out/bench_r1h.json
python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-200
python - <<'PY'
import sys; sys.path.insert(0,'.')
import torch
from diffsep_b200.score_model import ScoreModelNCSNpp
from diffsep_b200 import synthetic as ow
sm = ScoreModelNCSNpp(num_sources=2, backbone_args=dict(nf=128), state_dict=ow.make_score_model_state_dict(nf=128, seed=0))
x = torch.randn(32, 2, 32000, device="cuda"); t = torch.full((32,), 0.5, device="cuda"); m = torch.randn(32, 1, 32000, device="cuda")
sm(x, t, m); torch.cuda.synchronize()
pl = sm.backbone.plan(32, 256)
print("plan arena GB", pl.arena.total_bytes / 1e9, "steps", len(pl.steps), "torch reserved GB", torch.cuda.memory_reserved() / 1e9)
PY
