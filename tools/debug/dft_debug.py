import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests" / "golden"))
import torch
import cases
from diffsep_b200 import ops, synthetic as ow
from diffsep_b200.score_model import ScoreModelNCSNpp, LD, n_frames

def rel(a, b): return float((a.double() - b.double()).norm() / b.double().norm())
T = int(os.environ.get("T", "8000")); B = 1
sm = ScoreModelNCSNpp(num_sources=2, backbone_args=dict(nf=64), passes=3, state_dict=ow.make_score_model_state_dict(nf=64, seed=0))
xt, t, mix = (v.cuda() for v in cases.score_inputs(B, T, seed=3))
bf = sm._work(B, T); Fr = bf["Fr"]; ns = 2
for name, src_t, C_, key_f, key_d in (("mix", mix, 1, "frames_mix", "dft_mix"), ("xt", xt, ns, "frames", "dft")):
    M = B * C_ * Fr
    ops.stft_frames(src_t.contiguous(), sm.window, B, C_, T, Fr, bf[key_f])
    torch.cuda.synchronize()
    fr = bf[key_f]
    print(name, "M", M, "rows", fr.shape[0], "pad rows abs max", float(fr[M:].abs().max()) if fr.shape[0] > M else None,
          "cols>=510 max", float(fr[:M, 510:].abs().max()), "finite", bool(torch.isfinite(fr).all()))
    ref = torch.empty(M, LD, device="cuda")
    ops.sgemm(fr, LD, sm.basis_fwd, LD, ref, LD, M, LD, LD)
    sm._dft(fr, "fwd", bf[key_d], M)
    torch.cuda.synchronize()
    print("   fwd tc vs sgemm", rel(bf[key_d][:M], ref), "finite", bool(torch.isfinite(bf[key_d]).all()))
y1 = sm(xt, t, mix).clone()
torch.cuda.synchronize()
y1b = sm(xt, t, mix).clone()
sm._stft_tc = False
sm._bufs = {}
y2 = sm(xt, t, mix).clone()
torch.cuda.synchronize()
print("forward tc vs sgemm", rel(y1, y2), "second call", rel(y1b, y2))
