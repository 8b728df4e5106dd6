"""NCSN++ score backbone executed as a static plan of libdsep kernel launches.

Mirrors ``NCSNpp.forward`` (reference ``models/ncsnpp.py:319-478``) for the one configuration the
DiffSep hot path uses (BigGAN blocks, FIR resampling, input_skip / output_skip pyramids combined by
sum, Gaussian-Fourier time embedding, ch_mult (1,1,2,2,2,2,2), 2 res-blocks per level, attention
at H == 16).  Parameters come keyed as in the reference ``state_dict`` (``all_modules.{i}.…``).

B200-first design, not a module tree:
  * activations are channels-last fp32 ``[B, H, W, C]``; every conv operand is a pair of fp16
    planes (hi, lo) written by the pass that applies GroupNorm+SiLU (and, for the up/down blocks,
    the FIR resampling), so GN/SiLU/FIR/concat never cost an extra HBM round trip;
  * 3x3 / 1x1 convolutions and the NIN projections run on tcgen05 tensor cores
    (``dsep_conv2d_tc``) with bias, FiLM time-embedding bias, residual add and 1/sqrt(2) fused
    into the epilogue; ``torch.cat([h, skip])`` is never materialised in fp32;
  * all 49 ``Dense_0`` FiLM projections are one stacked GEMV per evaluation;
  * a plan (buffers + launch list) is built once per (B, W) and replayed; buffers are recycled
    through a liveness-aware arena so the working set stays small.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch

from . import ops
from .ops import Split

CH_MULT = (1, 1, 2, 2, 2, 2, 2)   # reference models/ncsnpp.py:50
NUM_RES_BLOCKS = 2                # :51
ATTN_RESOLUTIONS = (16,)          # :52
GN_EPS = 1e-6                     # layerspp.py:264-266
INV_SQRT2 = 1.0 / math.sqrt(2.0)


def _round_up(v, m):
    return (v + m - 1) // m * m


def gn_groups(c):
    return min(c // 4, 32)


class ConvWeight:
    """Tensor-core weight: split fp16 planes [taps, Cout_pad, Cin_pad] + fp32 bias."""

    def __init__(self, w_oihw, bias, device):
        w = w_oihw.detach().to(device=device, dtype=torch.float32)
        cout, cin, kh, kw = w.shape
        assert kh == kw and kh in (1, 3)
        self.ksize, self.cout, self.cin = kh, cout, cin
        self.cin_pad = _round_up(cin, 64)
        self.cout_pad = 16 if cout <= 16 else _round_up(cout, 64)
        wt = torch.zeros(kh * kw, self.cout_pad, self.cin_pad, device=device, dtype=torch.float32)
        wt[:, :cout, :cin] = w.permute(2, 3, 0, 1).reshape(kh * kw, cout, cin)
        # power-of-two pre-scale: largest |w| lands in [1024, 2048) so both fp16 planes stay normal
        amax = float(wt.abs().max())
        k = 0 if amax == 0.0 else 10 - math.floor(math.log2(amax))
        self.acc_scale = 2.0 ** (-k)
        self.planes = Split.empty(wt.shape, device)
        ops.split_f16(wt, self.planes, prescale=2.0 ** k)
        self.bias = None
        if bias is not None:
            b = torch.zeros(self.cout_pad, device=device, dtype=torch.float32)
            b[:cout] = bias.detach().to(device=device, dtype=torch.float32)
            self.bias = b


class Arena:
    """Recycles plan buffers by exact byte size (single stream, in-order execution)."""

    def __init__(self, device):
        self.device = device
        self.free = {}
        self.total_bytes = 0

    def _get(self, nbytes):
        lst = self.free.get(nbytes)
        if lst:
            return lst.pop()
        self.total_bytes += nbytes
        return torch.empty(nbytes, dtype=torch.uint8, device=self.device)

    def f32(self, *shape):
        n = 4 * math.prod(shape)
        raw = self._get(_round_up(n, 256))
        t = raw[:n].view(torch.float32).view(*shape)
        t._dsep_raw = raw
        return t

    def f64(self, *shape):
        n = 8 * math.prod(shape)
        raw = self._get(_round_up(n, 256))
        t = raw[:n].view(torch.float64).view(*shape)
        t._dsep_raw = raw
        return t

    def split(self, *shape):
        n = 2 * math.prod(shape)
        planes = []
        for _ in range(2):
            raw = self._get(_round_up(n, 256))
            t = raw[:n].view(torch.float16).view(*shape)
            t._dsep_raw = raw
            planes.append(t)
        return Split(*planes)

    def release(self, t):
        if t is None:
            return
        if isinstance(t, Split):
            self.release(t.hi)
            self.release(t.lo)
            return
        raw = t._dsep_raw
        self.free.setdefault(raw.numel(), []).append(raw)


class NCSNppB200:
    """The NCSN++ backbone: ``__call__(x_planes, x_pyramid, t) -> pyramid`` over a replayed plan."""

    def __init__(self, params, nf=128, ch_in=6, ch_out=4, device="cuda", passes=3):
        ops.require_device()
        self.device = torch.device(device)
        self.nf, self.ch_in, self.ch_out, self.passes = nf, ch_in, ch_out, passes
        self.temb_dim = 4 * nf
        self._plans = {}
        self._load(params)

    # ------------------------------------------------------------------ parameters
    def _load(self, params):
        dev = self.device
        P = {k: v for k, v in params.items()}

        def f32(name):
            return P[name].detach().to(device=dev, dtype=torch.float32).contiguous()

        def mod(i, k):
            return f"all_modules.{i}.{k}"

        self.Wf = f32(mod(0, "W"))
        self.t_w1, self.t_b1 = f32(mod(1, "weight")), f32(mod(1, "bias"))
        self.t_w2, self.t_b2 = f32(mod(2, "weight")), f32(mod(2, "bias"))
        self.conv_in = ConvWeight(P[mod(3, "weight")], P[mod(3, "bias")], dev)
        self.out_w = f32("output_layer.weight").reshape(self.ch_out, self.ch_in).contiguous()
        self.out_b = f32("output_layer.bias")

        dense_w, dense_b = [], []
        self._film_rows = 0

        def resblock(i):
            rb = {"idx": i}
            rb["gn0"] = (f32(mod(i, "GroupNorm_0.weight")), f32(mod(i, "GroupNorm_0.bias")))
            rb["gn1"] = (f32(mod(i, "GroupNorm_1.weight")), f32(mod(i, "GroupNorm_1.bias")))
            rb["conv0"] = ConvWeight(P[mod(i, "Conv_0.weight")], P[mod(i, "Conv_0.bias")], dev)
            rb["conv1"] = ConvWeight(P[mod(i, "Conv_1.weight")], P[mod(i, "Conv_1.bias")], dev)
            rb["conv2"] = None
            if mod(i, "Conv_2.weight") in P:
                rb["conv2"] = ConvWeight(P[mod(i, "Conv_2.weight")], P[mod(i, "Conv_2.bias")], dev)
            rb["film_off"] = self._film_rows
            dense_w.append(f32(mod(i, "Dense_0.weight")))
            dense_b.append(f32(mod(i, "Dense_0.bias")))
            self._film_rows += dense_w[-1].shape[0]
            rb["cin"], rb["cout"] = rb["conv0"].cin, rb["conv0"].cout
            return rb

        def attn(i):
            c = P[mod(i, "NIN_0.W")].shape[0]
            # NIN: y = x @ W + b with W stored (in, out) (layers.py:686-689) -> 1x1 conv weight W^T
            wqkv = torch.cat([P[mod(i, f"NIN_{j}.W")].detach().t() for j in range(3)], dim=0)
            bqkv = torch.cat([P[mod(i, f"NIN_{j}.b")].detach() for j in range(3)], dim=0)
            return {
                "idx": i, "c": c,
                "gn": (f32(mod(i, "GroupNorm_0.weight")), f32(mod(i, "GroupNorm_0.bias"))),
                "qkv": ConvWeight(wqkv.reshape(3 * c, c, 1, 1), bqkv, dev),
                "proj": ConvWeight(P[mod(i, "NIN_3.W")].detach().t().reshape(c, c, 1, 1), P[mod(i, "NIN_3.b")], dev),
            }

        # walk the module list exactly as NCSNpp.__init__ registers it (ncsnpp.py:104-308)
        m = 4
        nres = len(CH_MULT)
        self.down = []
        for lvl in range(nres):
            res = 256 // (2 ** lvl)
            level = {"blocks": [], "attn": [], "down": None, "combine": None}
            for _ in range(NUM_RES_BLOCKS):
                level["blocks"].append(resblock(m)); m += 1
                if res in ATTN_RESOLUTIONS:
                    level["attn"].append(attn(m)); m += 1
                else:
                    level["attn"].append(None)
            if lvl != nres - 1:
                level["down"] = resblock(m); m += 1
                w = f32(mod(m, "Conv_0.weight"))
                level["combine"] = (w.reshape(w.shape[0], w.shape[1]).contiguous(), f32(mod(m, "Conv_0.bias")))
                m += 1
            self.down.append(level)
        self.mid = [resblock(m), attn(m + 1), resblock(m + 2)]
        m += 3
        self.up = []
        for lvl in reversed(range(nres)):
            res = 256 // (2 ** lvl)
            level = {"lvl": lvl, "blocks": [], "attn": None, "up": None}
            for _ in range(NUM_RES_BLOCKS + 1):
                level["blocks"].append(resblock(m)); m += 1
            if res in ATTN_RESOLUTIONS:
                level["attn"] = attn(m); m += 1
            level["pyr_gn"] = (f32(mod(m, "weight")), f32(mod(m, "bias"))); m += 1
            level["pyr_conv"] = ConvWeight(P[mod(m, "weight")], P[mod(m, "bias")], dev); m += 1
            if lvl != 0:
                level["up"] = resblock(m); m += 1
            self.up.append(level)
        if mod(m, "weight") in P or mod(m, "GroupNorm_0.weight") in P:
            raise ValueError("unexpected extra modules in the backbone state dict")
        self.dense_w = torch.cat(dense_w, dim=0).contiguous()
        self.dense_b = torch.cat(dense_b, dim=0).contiguous()

    # ------------------------------------------------------------------ plan construction
    def plan(self, B, W):
        key = (B, W)
        if key not in self._plans:
            self._plans[key] = _Plan(self, B, W)
        return self._plans[key]

    def __call__(self, x_planes: Split, x_pyramid, t):
        """x_planes: split [B,256,W,64] network input (2x-1 applied, channels >= ch_in zero);
        x_pyramid: the same input as fp32 [B,256,W,ch_in]; t: [B].  Returns the output pyramid
        [B,256,W,ch_in] fp32 (before the /t scaling and the output 1x1 conv, which
        ``dsep_out_head`` fuses with the spectrogram decompression)."""
        B, H, W, _ = x_pyramid.shape
        assert H == 256 and W % 64 == 0
        return self.plan(B, W).run(x_planes, x_pyramid, t)


class _Plan:
    def __init__(self, net: NCSNppB200, B, W):
        self.net, self.B, self.W = net, B, W
        self.steps = []          # list of zero-arg callables
        self.arena = Arena(net.device)
        dev = net.device
        self.temb_act = torch.empty(B, net.temb_dim, device=dev, dtype=torch.float32)
        self.film = torch.empty(B, net._film_rows, device=dev, dtype=torch.float32)
        self.t_in = torch.empty(B, device=dev, dtype=torch.float32)
        self.x_planes = None     # bound at run time
        self.x_pyramid = None
        self._build()

    # -- helpers that append launches ------------------------------------------------------
    def _conv(self, a, H, W, cin_pad, cw: ConvWeight, out, cout_store, film=None, residual=None, scale=1.0):
        net, B = self.net, self.B
        film_v, stride = None, 0
        if film is not None:
            film_v, stride = self.film[:, film:], self.film.shape[1]
        self.steps.append(lambda: ops.conv2d_tc(
            a() if callable(a) else a, B, H, W, cin_pad, cw.planes, cw.cout_pad, cw.ksize, out, cout_store,
            bias=cw.bias, film=film_v, film_stride=stride, residual=residual, scale=scale,
            acc_scale=cw.acc_scale, passes=net.passes))

    def _resblock(self, rb, x0, C0, x1, C1, H, W, mode=0):
        """mode 0 plain, 1 up, 2 down.  Returns (out, Ho, Wo)."""
        ar, B = self.arena, self.B
        Cin, Cout = C0 + C1, rb["cout"]
        assert Cin == rb["cin"], (Cin, rb["cin"])
        g0 = gn_groups(Cin)
        stats0 = ar.f64(B, g0, 2)
        self.steps.append(lambda: ops.gn_stats(x0, C0, x1, C1, B, H * W, g0, stats0))
        Ho, Wo = (H * 2, W * 2) if mode == 1 else ((H // 2, W // 2) if mode == 2 else (H, W))
        a = ar.split(B, Ho, Wo, Cin)
        r = ar.split(B, Ho, Wo, Cin) if rb["conv2"] is not None else None
        gam0, bet0 = rb["gn0"]
        if mode == 0:
            self.steps.append(lambda: ops.gn_act_split(x0, C0, x1, C1, B, H * W, g0, stats0, gam0, bet0, GN_EPS,
                                                       1, a=a, r=r))
        else:
            assert x1 is None and r is not None
            self.steps.append(lambda: ops.fir_resample(x0, B, H, W, Cin, mode, g0, stats0, gam0, bet0, GN_EPS,
                                                       a=a, r=r))
        h = ar.f32(B, Ho, Wo, Cout)
        self._conv(a, Ho, Wo, Cin, rb["conv0"], h, Cout, film=rb["film_off"])
        ar.release(a)
        ar.release(stats0)
        g1 = gn_groups(Cout)
        stats1 = ar.f64(B, g1, 2)
        self.steps.append(lambda: ops.gn_stats(h, Cout, None, 0, B, Ho * Wo, g1, stats1))
        a2 = ar.split(B, Ho, Wo, Cout)
        gam1, bet1 = rb["gn1"]
        self.steps.append(lambda: ops.gn_act_split(h, Cout, None, 0, B, Ho * Wo, g1, stats1, gam1, bet1, GN_EPS,
                                                   1, a=a2))
        ar.release(stats1)
        if rb["conv2"] is not None:
            xs = ar.f32(B, Ho, Wo, Cout)
            self._conv(r, Ho, Wo, Cin, rb["conv2"], xs, Cout)
            ar.release(r)
        else:
            assert x1 is None and Cin == Cout
            xs = x0
        out = h   # conv1 never reads h (only a2), so its buffer is reused for the block output
        self._conv(a2, Ho, Wo, Cout, rb["conv1"], out, Cout, residual=xs, scale=INV_SQRT2)
        ar.release(a2)
        if rb["conv2"] is not None:
            ar.release(xs)
        return out, Ho, Wo

    def _attn(self, at, x, H, W):
        ar, B, Cc = self.arena, self.B, at["c"]
        S = H * W
        g = gn_groups(Cc)
        stats = ar.f64(B, g, 2)
        self.steps.append(lambda: ops.gn_stats(x, Cc, None, 0, B, S, g, stats))
        a = ar.split(B, H, W, Cc)
        gam, bet = at["gn"]
        self.steps.append(lambda: ops.gn_act_split(x, Cc, None, 0, B, S, g, stats, gam, bet, GN_EPS, 0, a=a))
        ar.release(stats)
        qkv = ar.f32(B, S, 3 * Cc)
        self._conv(a, H, W, Cc, at["qkv"], qkv, 3 * Cc)
        o = a   # the GN'd input planes are dead once q, k, v exist
        self.steps.append(lambda: ops.attention(qkv, B, S, Cc, float(Cc) ** -0.5, o))
        out = ar.f32(B, H, W, Cc)
        self._conv(o, H, W, Cc, at["proj"], out, Cc, residual=x, scale=INV_SQRT2)
        ar.release(o)
        ar.release(qkv)
        return out

    def _build(self):
        net, ar, B, W0 = self.net, self.arena, self.B, self.W
        nf, ch_in = net.nf, net.ch_in
        nres = len(CH_MULT)
        # refcounts for skip tensors: released after their last consumer
        H, W = 256, W0
        h = ar.f32(B, H, W, nf)
        self._conv(lambda: self.x_planes, H, W, net.conv_in.cin_pad, net.conv_in, h, nf)
        hs = [(h, nf)]
        pyr_in = None            # running input pyramid (fp32, ch_in channels); level 0 = x_pyramid
        cur_c = nf
        for lvl, level in enumerate(net.down):
            for rb, at in zip(level["blocks"], level["attn"]):
                x_prev, c_prev = hs[-1]
                h, _, _ = self._resblock(rb, x_prev, c_prev, None, 0, H, W)
                cur_c = rb["cout"]
                if at is not None:
                    h2 = self._attn(at, h, H, W)
                    ar.release(h)
                    h = h2
                hs.append((h, cur_c))
            if lvl != nres - 1:
                x_prev, c_prev = hs[-1]
                h, Hn, Wn = self._resblock(level["down"], x_prev, c_prev, None, 0, H, W, mode=2)
                new_pyr = ar.f32(B, Hn, Wn, ch_in)
                src = pyr_in
                Hc, Wc = H, W
                if src is None:
                    self.steps.append(lambda Hc=Hc, Wc=Wc, new_pyr=new_pyr: ops.fir_resample(
                        self.x_pyramid, B, Hc, Wc, ch_in, 2, y=new_pyr))
                else:
                    self.steps.append(lambda src=src, Hc=Hc, Wc=Wc, new_pyr=new_pyr: ops.fir_resample(
                        src, B, Hc, Wc, ch_in, 2, y=new_pyr))
                    ar.release(src)
                pyr_in = new_pyr
                cw, cb = level["combine"]
                self.steps.append(lambda pyr=new_pyr, cw=cw, cb=cb, h=h, P=Hn * Wn, c=cur_c: ops.combine(
                    pyr, ch_in, cw, cb, h, h, B, P, c))
                H, W = Hn, Wn
                hs.append((h, cur_c))
        if pyr_in is not None:
            ar.release(pyr_in)

        h, cur_c = hs[-1]
        h_mid, _, _ = self._resblock(net.mid[0], h, cur_c, None, 0, H, W)
        h2 = self._attn(net.mid[1], h_mid, H, W)
        ar.release(h_mid)
        h3, _, _ = self._resblock(net.mid[2], h2, cur_c, None, 0, H, W)
        ar.release(h2)
        h = h3

        pyramid = None
        for level in net.up:
            for rb in level["blocks"]:
                skip, c_skip = hs.pop()
                h_new, _, _ = self._resblock(rb, h, cur_c, skip, c_skip, H, W)
                ar.release(h)
                ar.release(skip)
                h, cur_c = h_new, rb["cout"]
            if level["attn"] is not None:
                h2 = self._attn(level["attn"], h, H, W)
                ar.release(h)
                h = h2
            # output pyramid: conv3x3(SiLU(GN(h))) (+ FIR-up of the running pyramid), ncsnpp.py:419-440
            g = gn_groups(cur_c)
            stats = ar.f64(B, g, 2)
            self.steps.append(lambda h=h, c=cur_c, P=H * W, g=g, stats=stats: ops.gn_stats(h, c, None, 0, B, P, g, stats))
            a = ar.split(B, H, W, cur_c)
            gam, bet = level["pyr_gn"]
            self.steps.append(lambda h=h, c=cur_c, P=H * W, g=g, stats=stats, gam=gam, bet=bet, a=a:
                              ops.gn_act_split(h, c, None, 0, B, P, g, stats, gam, bet, GN_EPS, 1, a=a))
            ar.release(stats)
            new_pyr = ar.f32(B, H, W, ch_in)
            up = None
            if pyramid is not None:
                up = ar.f32(B, H, W, ch_in)
                self.steps.append(lambda src=pyramid, Hs=H // 2, Ws=W // 2, up=up: ops.fir_resample(
                    src, B, Hs, Ws, ch_in, 1, y=up))
                ar.release(pyramid)
            self._conv(a, H, W, cur_c, level["pyr_conv"], new_pyr, ch_in, residual=up)
            ar.release(a)
            if up is not None:
                ar.release(up)
            pyramid = new_pyr
            if level["up"] is not None:
                h_new, H, W = self._resblock(level["up"], h, cur_c, None, 0, H, W, mode=1)
                ar.release(h)
                h = h_new
        assert not hs and H == 256 and W == W0
        ar.release(h)
        self.out = pyramid

    def run(self, x_planes, x_pyramid, t):
        net, B = self.net, self.B
        self.x_planes, self.x_pyramid = x_planes, x_pyramid
        ops.time_embedding(t, net.Wf, net.t_w1, net.t_b1, net.t_w2, net.t_b2, B, net.nf, self.temb_act)
        ops.film(self.temb_act, net.dense_w, net.dense_b, B, net.temb_dim, net._film_rows, self.film)
        for step in self.steps:
            step()
        return self.out
