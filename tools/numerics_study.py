"""How many tensor-core products per MAC does the 1e-4 parity budget allow?  (CPU, float64 oracle.)

Emulates operand rounding of the convolutions of the NCSN++ oracle (`oracle/ncsnpp_ref.py`, float64) exactly as
a tensor-core pass structure would produce it, and reports the rel-L2 error of the network output against the
unrounded float64 evaluation.  With x = hi + lo (fp16 planes, 11 + 11 significand bits):

    3 products  hi*hi + lo*hi + hi*lo   -> operands effectively 22 bits        (the shipped parity mode)
    2 products  drop lo*hi              -> activations rounded to 11 bits, weights 22
    2 products  drop hi*lo              -> weights rounded to 11 bits, activations 22
    1 product   hi*hi                   -> both rounded to 11 bits             (TF32-grade mode)
    3 products with fp8 corrections     -> correction operands rounded to 4 bits (e4m3), at 2x tensor rate

Policies can be restricted to the layers that hold most FLOPs (the 256-row level-0 3x3 convs: 57.6 %).
Accumulation is float64 here, so this isolates OPERAND precision (the tensor core's truncating fp32
accumulator adds 0.3e-6 ... 9e-6 per conv on top, see tests/test_ops_gpu.py).

    python tools/numerics_study.py [W=128] [t=0.6]      # a few minutes on 8 cores; writes nothing, prints a table
    python tools/numerics_study.py --sampler            # per-step error of the PC sampler's updates (N = 30)
"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests" / "golden"))

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from oracle import ncsnpp_ref as nr, weights as ow  # noqa: E402

SAMPLER = "--sampler" in sys.argv
if SAMPLER:
    sys.argv.remove("--sampler")
W = int(sys.argv[1]) if len(sys.argv) > 1 else 128
torch.manual_seed(0)
torch.set_num_threads(8)
_real_conv2d = F.conv2d


def round_bits(x, bits):
    """round to `bits` significand bits (round-to-nearest-even on the float64 value), no range limits"""
    m, e = torch.frexp(x)
    return torch.ldexp(torch.round(m * (1 << bits)) / (1 << bits), e)


def hi16(x):
    return round_bits(x, 11)


def e4m3(x):
    """what a power-of-two-prescaled e4m3 plane holds: the tensor is scaled so that its largest magnitude lands in
    [224, 448], then 4 significand bits above 2^-6, multiples of 2^-9 below (subnormals), saturation at 448"""
    amax = float(x.abs().max())
    if amax == 0.0:
        return x
    import math
    k = math.floor(math.log2(448.0 / amax))
    s = torch.ldexp(x, torch.tensor(k))
    normal = round_bits(s, 4)
    sub = torch.round(s * 512.0) / 512.0
    q = torch.where(s.abs() >= 2.0 ** -6, normal, sub).clamp(-448.0, 448.0)
    return torch.ldexp(q, torch.tensor(-k))


class Policy:
    """mode per conv call: '3', 'dropA' (activations 11 bits), 'dropW', '1', 'fp8corr'; select(weight, x) -> bool"""
    def __init__(self, mode, select=lambda w, x: True):
        self.mode, self.select, self.calls, self.hit = mode, select, 0, 0

    def conv(self, x, w, b=None, **kw):
        self.calls += 1
        if w.shape[1] < 16 or kw.get("groups", 1) != 1 or not self.select(w, x) or self.mode == "3":
            return _real_conv2d(x, w, b, **kw)          # FIR (grouped) and the tiny 4/6-channel heads stay exact
        self.hit += 1
        if self.mode == "dropA":
            return _real_conv2d(hi16(x), w, b, **kw)
        if self.mode == "dropW":
            return _real_conv2d(x, hi16(w), b, **kw)
        if self.mode == "1":
            return _real_conv2d(hi16(x), hi16(w), b, **kw)
        if self.mode == "fp8corr":       # hi*hi exact + both correction products with 4-bit operands
            xh, wh = hi16(x), hi16(w)
            y = _real_conv2d(xh, wh, b, **kw)
            y = y + _real_conv2d(e4m3(x - xh), e4m3(wh), None, **kw)
            return y + _real_conv2d(e4m3(xh), e4m3(w - wh), None, **kw)
        if self.mode == "fp8corr_ideal":   # 4 significand bits, unlimited range
            xh, wh = hi16(x), hi16(w)
            y = _real_conv2d(xh, wh, b, **kw)
            y = y + _real_conv2d(round_bits(x - xh, 4), round_bits(wh, 4), None, **kw)
            return y + _real_conv2d(round_bits(xh, 4), round_bits(w - wh, 4), None, **kw)
        if self.mode == "e5m2corr":        # 3 significand bits
            xh, wh = hi16(x), hi16(w)
            y = _real_conv2d(xh, wh, b, **kw)
            y = y + _real_conv2d(round_bits(x - xh, 3), round_bits(wh, 3), None, **kw)
            return y + _real_conv2d(round_bits(xh, 3), round_bits(w - wh, 3), None, **kw)
        raise ValueError(self.mode)


def run(policy, params, x, t):
    nr.F.conv2d = policy.conv if policy is not None else _real_conv2d
    try:
        with torch.no_grad():
            return nr.ncsnpp_forward(params, x, t)
    finally:
        nr.F.conv2d = _real_conv2d


def rel(a, b):
    return float((a - b).norm() / b.norm())


def main():
    params = {k: v.double() for k, v in ow.make_backbone_params(nf=128, seed=0).items()}
    g = torch.Generator().manual_seed(5)
    x = torch.rand(1, 6, 256, W, generator=g, dtype=torch.float64)
    t = torch.tensor([float(sys.argv[2]) if len(sys.argv) > 2 else 0.6], dtype=torch.float64)
    truth = run(None, params, x, t)
    level0_3x3 = lambda w, xx: xx.shape[-2] == 256 and w.shape[-1] == 3
    level01_3x3 = lambda w, xx: xx.shape[-2] >= 128 and w.shape[-1] == 3
    one_conv = {"n": 0}

    def first_level0(w, xx):
        if level0_3x3(w, xx):
            one_conv["n"] += 1
            return one_conv["n"] == 2
        return False
    rows = [("1 product everywhere (TF32-grade mode)", Policy("1")),
            ("2 products everywhere, lo*hi dropped (activations 11 bits)", Policy("dropA")),
            ("2 products everywhere, hi*lo dropped (weights 11 bits)", Policy("dropW")),
            ("2 products (activations 11 bits) in the level-0 3x3 convs only (57.6 % of FLOPs)", Policy("dropA", level0_3x3)),
            ("2 products (weights 11 bits) in the level-0 3x3 convs only", Policy("dropW", level0_3x3)),
            ("2 products (activations 11 bits) in levels 0-1 3x3 convs (80.6 % of FLOPs)", Policy("dropA", level01_3x3)),
            ("2 products (activations 11 bits) in ONE level-0 3x3 conv", Policy("dropA", first_level0)),
            ("3 products, both corrections with e4m3 operands (per-tensor 2^k prescale, saturating), everywhere", Policy("fp8corr")),
            ("same with unlimited exponent range (4 significand bits)", Policy("fp8corr_ideal")),
            ("3 products, e4m3 corrections in the level-0 3x3 convs only", Policy("fp8corr", level0_3x3)),
            ("3 products, corrections with 3-significand-bit operands (e5m2 precision), everywhere", Policy("e5m2corr"))]
    print(f"NCSN++ nf=128, input [1, 6, 256, {W}], float64 accumulation; rel-L2 of the output vs unrounded float64\n")
    print("| operand scheme | convs affected | rel-L2 error |\n|---|---:|---:|")
    for name, pol in rows:
        out = run(pol, params, x, t)
        print(f"| {name} | {pol.hit} of {pol.calls} | {rel(out, truth):.2e} |", flush=True)


def sampler_study(T=8000, N=30, steps=(0, 9, 19, 29)):
    """Per-step view (the north-star criterion: per-step output within 1e-4 rel-L2): the reference's PC sampler
    (reverse_diffusion + ald2, 1 corrector step, N = 30) runs in float64 with exact convolutions; at a few steps the
    corrector and the predictor update are repeated FROM THE EXACT STATE with each operand scheme, and the state
    they produce is compared with the exact one."""
    from oracle import score_ref as sr, sde_ref as sd
    import cases
    params = {k: v.double() for k, v in ow.make_backbone_params(nf=128, seed=0).items()}
    mix = sd.normalize_batch(cases.batch_mix(1, T).double())[0]
    noises = [n.double() for n in cases.sampler_noises(1, T, N, 1)]
    p = sd.MixSDEParams(N=N)
    ts = sd.timesteps(p, 0.03, None, mix.dtype)

    def score(policy):
        def fn(x, t, m):
            nr.F.conv2d = policy.conv if policy is not None else _real_conv2d
            try:
                with torch.no_grad():
                    return sr.score_forward(params, x, t, m)
            finally:
                nr.F.conv2d = _real_conv2d
        return fn
    schemes = (("1 product", "1"), ("fp16 hi*hi + e4m3 corrections", "fp8corr"))
    print(f"\nPC sampler N={N}, 1 corrector step, T={T}, float64: rel-L2 of ONE update from the exact state\n")
    print("| step (t) | " + " | ".join(f"{n}: corrector | predictor" for n, _ in schemes) + " |")
    print("|---|" + "---:|" * (2 * len(schemes)))
    xt = sd.prior_sampling(p, mix, noises[0])
    for i in range(N):
        vec_t = torch.ones(1, dtype=mix.dtype) * ts[i]
        zc, zp = noises[1 + 2 * i], noises[2 + 2 * i]
        xc, _ = sd.corrector_step(p, score(None), xt, vec_t, mix, [zc], 0.5)
        xp, _ = sd.predictor_step(p, score(None), xc, vec_t, mix, zp)
        if i in steps:
            cells = []
            for _, mode in schemes:
                gc, _ = sd.corrector_step(p, score(Policy(mode)), xt, vec_t, mix, [zc], 0.5)
                gp, _ = sd.predictor_step(p, score(Policy(mode)), xc, vec_t, mix, zp)
                cells += [f"{rel(gc, xc):.2e}", f"{rel(gp, xp):.2e}"]
            print(f"| {i + 1} ({float(ts[i]):.2f}) | " + " | ".join(cells) + " |", flush=True)
        xt = xp


if __name__ == "__main__":
    if SAMPLER:
        sampler_study()
    else:
        main()
