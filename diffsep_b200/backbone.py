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
import os

import torch

from . import ops
from .ops import Split

CH_MULT = (1, 1, 2, 2, 2, 2, 2)   # reference models/ncsnpp.py:50
NUM_RES_BLOCKS = 2                # :51
ATTN_RESOLUTIONS = (16,)          # :52
GN_EPS = 1e-6                     # layerspp.py:264-266
INV_SQRT2 = 1.0 / math.sqrt(2.0)
# output-pyramid convs (C -> 6) on maps of at least this many pixels run as 1x1 conv + tap gather (0: never)
PYR_TAPS_MIN_PIXELS = int(os.environ.get("DSEP_PYR_TAPS_MIN", str(64 * 64))) or (1 << 62)


def _round_up(v, m):
    return (v + m - 1) // m * m


def cin_align():
    """channel granularity of tensor-core operands (the conv kernel's K-block)"""
    from . import _lib
    return _lib.load().dsep_conv_kblock()


def gn_groups(c):
    return min(c // 4, 32)


def _prescale_exp(amax):
    """power of two k with amax * 2^k in [2^14, 2^15): both fp16 planes of a weight stay normal, and the e4m3 planes
    of the 2-unit mode (W_hi * 2^-11 < 16, W_lo <= 16) keep 13 binades below the largest weight."""
    return 0 if amax == 0.0 else 14 - math.floor(math.log2(amax))


class ConvWeight:
    """Tensor-core weight: split fp16 planes [taps, Cout_pad, Cin_pad] + fp32 bias.

    ``shortcut`` (a 1x1 ConvWeight source: weight, bias) is folded in as the fused second operand of
    ``dsep_conv2d_tc``: both weights share one power-of-two pre-scale and the biases are summed."""

    def __init__(self, w_oihw, bias, device, shortcut=None, cout_pad=None):
        w = w_oihw.detach().to(device=device, dtype=torch.float32)
        cout, cin, kh, kw = w.shape
        assert kh == kw and kh in (1, 3)
        self.ksize, self.cout, self.cin = kh, cout, cin
        self.cin_pad = _round_up(cin, cin_align())
        self.cout_pad = (16 if cout <= 16 else _round_up(cout, 64)) if cout_pad is None else cout_pad
        assert self.cout_pad >= cout
        wt = torch.zeros(kh * kw, self.cout_pad, self.cin_pad, device=device, dtype=torch.float32)
        wt[:, :cout, :cin] = w.permute(2, 3, 0, 1).reshape(kh * kw, cout, cin)
        amax = float(wt.abs().max())
        w2t = None
        if shortcut is not None:
            w2 = shortcut[0].detach().to(device=device, dtype=torch.float32)
            assert w2.shape[0] == cout and w2.shape[2:] == (1, 1)
            self.cin2 = w2.shape[1]
            self.cin2_pad = _round_up(self.cin2, cin_align())
            w2t = torch.zeros(self.cout_pad, self.cin2_pad, device=device, dtype=torch.float32)
            w2t[:cout, :self.cin2] = w2.reshape(cout, self.cin2)
            amax = max(amax, float(w2t.abs().max()))
        else:
            self.cin2 = self.cin2_pad = 0
        k = _prescale_exp(amax)
        self.acc_scale = 2.0 ** (-k)
        self.planes = Split.empty(wt.shape, device)
        ops.split_f16(wt, self.planes, prescale=2.0 ** k)
        self.planes2 = None
        if w2t is not None:
            self.planes2 = Split.empty(w2t.shape, device)
            ops.split_f16(w2t, self.planes2, prescale=2.0 ** k)
        self.bias = None
        if bias is not None or (shortcut is not None and shortcut[1] is not None):
            b = torch.zeros(self.cout_pad, device=device, dtype=torch.float32)
            if bias is not None:
                b[:cout] += bias.detach().to(device=device, dtype=torch.float32)
            if shortcut is not None and shortcut[1] is not None:
                b[:cout] += shortcut[1].detach().to(device=device, dtype=torch.float32)
            self.bias = b


    # ---- passes = 2: e4m3 correction plane
    A8_EXP = 0            # activations enter the e4m3 planes as A_hi (saturating above 448) and A_lo * 2^11; with no
                          # prescale the builder converts A_hi8 straight from the packed fp16 pairs
    W8_EXP = -11          # W_hi8 = e4m3(W_hi * 2^-11), W_lo8 = e4m3(W_lo): with A_lo8 = e4m3(A_lo * 2^11) and A_hi8 =
                          # e4m3(A_hi) both correction products come out at the scale of the fp16 product, so all three
                          # accumulate into ONE TMEM accumulator (corr_rel == 1; conv_wide.cu relies on it)

    def planes8(self) -> Split:
        """(hi = the fp16 hi plane, lo = the e4m3 plane [taps, Cout_pad, 2 * Cin_pad] bytes viewed as fp16): per 8
        input channels the 16 bytes [W_hi8 x 8 | W_lo8 x 8] the kernel pairs with [A_lo8 x 8 | A_hi8 x 8]."""
        if getattr(self, "_planes8", None) is None:
            hi, lo = self.planes.hi.float(), self.planes.lo.float()
            t, n, c = hi.shape
            h8 = (hi * 2.0 ** self.W8_EXP).clamp(-448.0, 448.0).to(torch.float8_e4m3fn).view(torch.uint8)
            l8 = (lo * 2.0 ** (self.W8_EXP + 11)).clamp(-448.0, 448.0).to(torch.float8_e4m3fn).view(torch.uint8)
            w8 = torch.cat([h8.reshape(t, n, c // 8, 8), l8.reshape(t, n, c // 8, 8)], dim=-1).contiguous()
            self._planes8 = Split(self.planes.hi, w8.view(torch.float16).reshape(t, n, c))
        return self._planes8

    @property
    def corr_rel(self):
        return 2.0 ** -(self.W8_EXP + self.A8_EXP + 11)


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

    def __init__(self, params, nf=128, ch_in=6, ch_out=4, device="cuda", passes=None, fuse=None):
        ops.require_device()
        self.device = torch.device(device)
        from . import DEFAULT_PASSES
        passes = DEFAULT_PASSES if passes is None else int(passes)
        self.nf, self.ch_in, self.ch_out, self.passes = nf, ch_in, ch_out, passes
        if passes == 2:
            from . import _lib
            if not _lib.load().dsep_has_fp8_corr():
                raise NotImplementedError("passes=2 (e4m3 correction products): this libdsep was built with "
                                          "-DDSEP_FP8_CORR=0")
        # convs that still take operand planes (small maps, FIR-resampled inputs, narrow outputs) stay at 3 products
        self.plane_passes = 3 if passes == 2 else passes
        # in-kernel GroupNorm/SiLU/concat prologue (dsep_conv2d_fused) where shapes allow.  With three
        # MMA passes per tile the worker warps have time to build the patches for free; with one pass
        # they would become the bottleneck (measured), so that mode keeps the separate pass.
        default_fuse = "1" if passes in (2, 3) else "0"
        self.fuse = bool(int(os.environ.get("DSEP_FUSE", default_fuse))) if fuse is None else bool(fuse)
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
        # the input conv as a 1x1 convolution over im2col rows (dsep_im2col3x3): 9 * ch_in = 54 -> ONE 64-channel
        # K-block through the fused kernel instead of nine taps of a 6 -> 64 padded plane operand
        self.conv_in_col = None
        if self.fuse and self.passes == 2 and cin_align() == 64 and 9 * self.ch_in <= 64 \
                and int(os.environ.get("DSEP_CONV_IN_COL", "1")):
            w_in = P[mod(3, "weight")].detach().to(device=dev, dtype=torch.float32)       # [nf, ch_in, 3, 3]
            w_col = torch.zeros(w_in.shape[0], 64, 1, 1, device=dev, dtype=torch.float32)
            w_col[:, :9 * self.ch_in, 0, 0] = w_in.permute(0, 2, 3, 1).reshape(w_in.shape[0], 9 * self.ch_in)
            self.conv_in_col = ConvWeight(w_col, P[mod(3, "bias")], dev)
        self.out_w = f32("output_layer.weight").reshape(self.ch_out, self.ch_in).contiguous()
        self.out_b = f32("output_layer.bias")

        dense_w, dense_b = [], []
        self._film_rows = 0

        def resblock(i):
            rb = {"idx": i}
            rb["gn0"] = (f32(mod(i, "GroupNorm_0.weight")), f32(mod(i, "GroupNorm_0.bias")))
            rb["gn1"] = (f32(mod(i, "GroupNorm_1.weight")), f32(mod(i, "GroupNorm_1.bias")))
            rb["conv0"] = ConvWeight(P[mod(i, "Conv_0.weight")], P[mod(i, "Conv_0.bias")], dev)
            shortcut = None
            if mod(i, "Conv_2.weight") in P:     # 1x1 shortcut: fused into Conv_1's accumulation
                shortcut = (P[mod(i, "Conv_2.weight")], P[mod(i, "Conv_2.bias")])
            rb["conv1"] = ConvWeight(P[mod(i, "Conv_1.weight")], P[mod(i, "Conv_1.bias")], dev, shortcut=shortcut)
            rb["has_shortcut"] = shortcut is not None
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
            level["pyr_conv"] = ConvWeight(P[mod(m, "weight")], P[mod(m, "bias")], dev)
            # the same conv as "1x1 to 9 * ch_in tap-major channels, then a 9-term gather" (dsep_tap_gather3x3): row
            # tap * ch_in + co of the 1x1 weight is W[co, :, ky, kx]; 54 rows padded to the wide-tile kernel's 128
            wp = P[mod(m, "weight")]
            if self.passes == 2 and wp.shape[2:] == (3, 3) and 9 * wp.shape[0] <= 128 and wp.shape[1] % 64 == 0:
                w_taps = wp.detach().permute(2, 3, 0, 1).reshape(9 * wp.shape[0], wp.shape[1], 1, 1)
                level["pyr_taps"] = ConvWeight(w_taps, None, dev, cout_pad=128)
            m += 1
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

    def drop_plan(self, B, W, keep=()):
        """forget the launch plan (and its arena) of a shape no caller uses any more"""
        if (B, W) not in keep:
            self._plans.pop((B, W), None)

    def film_rows(self, ts):
        """FiLM rows [len(ts), R] of the given times (one embedding-MLP launch + one stacked Dense_0 launch)."""
        n = len(ts)
        t = torch.tensor(list(ts), dtype=torch.float32, device=self.device)
        temb = torch.empty(n, self.temb_dim, device=self.device, dtype=torch.float32)
        out = torch.empty(n, self._film_rows, device=self.device, dtype=torch.float32)
        ops.time_embedding(t, self.Wf, self.t_w1, self.t_b1, self.t_w2, self.t_b2, n, self.nf, temb)
        ops.film(temb, self.dense_w, self.dense_b, n, self.temb_dim, self._film_rows, out)
        return out

    def __call__(self, x_planes: Split, x_pyramid, t, uniform=False):
        """x_planes: split [B,256,W,conv_in.cin_pad] network input (2x-1 applied, channels >= ch_in zero);
        x_pyramid: the same input as fp32 [B,256,W,ch_in]; t: [B].  Returns the output pyramid
        [B,256,W,ch_in] fp32 (before the /t scaling and the output 1x1 conv, which
        ``dsep_out_head`` fuses with the spectrogram decompression)."""
        B, H, W, _ = x_pyramid.shape
        assert H == 256 and W % 64 == 0
        return self.plan(B, W).run(x_planes, x_pyramid, t, uniform=uniform)


class Act:
    """An fp32 activation [B, H, W, C] plus (once known) its per-channel GroupNorm sums [B, C, 2]."""

    __slots__ = ("t", "C", "H", "W", "st")

    def __init__(self, t, C, H, W, st=None):
        self.t, self.C, self.H, self.W, self.st = t, C, H, W, st


def _tile_has_one_batch_entry(H, W):
    """mirrors the 128-pixel tile choice of dsep_conv2d_tc: fused statistics need tb == 1"""
    def p2(v):
        l = 1
        while l < v:
            l *= 2
        return l
    tw = min(16, p2(W))
    th = min(128 // tw, p2(H))
    return tw * th == 128


class _Plan:
    STATS_CHUNK = 8 << 20

    def __init__(self, net: NCSNppB200, B, W):
        self.net, self.B, self.W = net, B, W
        self.steps = []          # list of zero-arg callables
        self.arena = Arena(net.device)
        dev = net.device
        self.temb_act = torch.empty(B, net.temb_dim, device=dev, dtype=torch.float32)
        self.film = torch.empty(B, net._film_rows, device=dev, dtype=torch.float32)
        self.film_stride = net._film_rows    # 0 while every batch entry shares one time: row 0 serves them all
        self.x_planes = None     # bound at run time
        self.x_pyramid = None
        self._stat_chunks, self._stat_used = [], self.STATS_CHUNK
        self._build()

    # -- GroupNorm statistics slots (zeroed once per evaluation, never recycled within one) ----------
    def _slot(self, C):
        n = self.B * C * 2
        if self._stat_used + n * 8 > self.STATS_CHUNK:
            self._stat_chunks.append(torch.empty(max(self.STATS_CHUNK, n * 8) // 8, device=self.net.device,
                                                 dtype=torch.float64))
            self._stat_used = 0
        off = self._stat_used // 8
        self._stat_used += _round_up(n * 8, 256)
        return self._stat_chunks[-1][off:off + n].view(self.B, C, 2)

    def _ensure_stats(self, x: Act):
        if x.st is None:
            x.st = self._slot(x.C)
            t, C, P, st, B = x.t, x.C, x.H * x.W, x.st, self.B
            self.steps.append(lambda: ops.channel_stats(t, C, B, P, st))
        return x.st

    def _stats_or_defer(self, x: Act):
        """-> (statistics slot, compute flag).  On small maps (at most 64 pixels per batch entry: the conv tiles there
        span several batch entries, so producers leave no fused statistics) the consuming gn_act_split computes the
        sums itself and writes them to the slot (dsep_gn_stats_act_split): no channel_stats launch."""
        if x.st is None and x.H * x.W <= 64 and x.C % 64 == 0 and 256 % (x.C // 4) == 0 \
                and int(os.environ.get("DSEP_STATS_INLINE", "1")):
            x.st = self._slot(x.C)
            return x.st, 1
        return self._ensure_stats(x), 0

    def _fused_slot(self, H, W, C):
        return self._slot(C) if (C >= 64 and _tile_has_one_batch_entry(H, W)) else None

    # -- helpers that append launches ------------------------------------------------------
    def _conv(self, a, H, W, cin_pad, cw: ConvWeight, out, cout_store, film=None, residual=None, scale=1.0,
              a2=None, stats=None, e4m3=False):
        """e4m3: ``a`` holds (fp16 hi, e4m3 correction) planes (fir_resample with a8_exp) -> passes = 2."""
        net, B = self.net, self.B
        film_v, has_film = None, film is not None
        if has_film:
            film_v = self.film[:, film:]
        w2 = cw.planes2 if a2 is not None else None
        cin2 = cw.cin2_pad if a2 is not None else 0
        if e4m3:
            w8 = cw.planes8()
            self.steps.append(lambda: ops.conv2d_tc(
                a, B, H, W, cin_pad, w8, cw.cout_pad, cw.ksize, out, cout_store, bias=cw.bias, film=film_v,
                film_stride=self.film_stride if has_film else 0, residual=residual, scale=scale, acc_scale=cw.acc_scale, passes=2, a2=a2,
                Cin2=cin2, w2=w2, stats=stats, corr_rel=cw.corr_rel, a8_exp=cw.A8_EXP))
            return
        self.steps.append(lambda: ops.conv2d_tc(
            a() if callable(a) else a, B, H, W, cin_pad, cw.planes, cw.cout_pad, cw.ksize, out, cout_store,
            bias=cw.bias, film=film_v, film_stride=self.film_stride if has_film else 0, residual=residual, scale=scale,
            acc_scale=cw.acc_scale, passes=net.plane_passes, a2=a2, Cin2=cin2, w2=w2, stats=stats))

    def _fusable(self, H, W, *channels):
        """dsep_conv2d_fused's in-kernel prologue: map of at least 16 x 8, 64-channel granularity"""
        return (self.net.fuse and cin_align() == 64 and H >= 16 and W >= 8
                and all(c % 64 == 0 for c in channels if c))

    def _tables(self, st0, C0, st1, C1, P, gamma, beta):
        """GroupNorm scale/shift tables [B, C0+C1] for the fused prologue (tiny kernel)."""
        ar, B = self.arena, self.B
        sc, sh = ar.f32(B, C0 + C1), ar.f32(B, C0 + C1)
        g = gn_groups(C0 + C1)
        self.steps.append(lambda: ops.gn_tables(st0, C0, st1, C1, B, P, g, gamma, beta, GN_EPS, sc, sh))
        return sc, sh

    def _conv_fused(self, H, W, cin, cw: ConvWeight, out, cout_store, x0, C0, x1, C1, sc, sh, act, film=None,
                    residual=None, scale=1.0, shortcut_raw=None, stats=None):
        net, B = self.net, self.B
        film_v, has_film = None, film is not None
        if has_film:
            film_v = self.film[:, film:]
        kw = {}
        if shortcut_raw is not None:
            s0, S0, s1, S1 = shortcut_raw
            kw = dict(s0=s0, S0=S0, s1=s1, S1=S1, Cin2=cw.cin2_pad, w2=cw.planes2)
        if net.passes == 2 and cw.cout_pad >= 64:      # fp16 hi*hi + one e4m3 product for both correction terms
            w8 = cw.planes8()                          # built now, not inside a replayed / captured step
            self.steps.append(lambda: ops.conv2d_fused(
                B, H, W, cin, w8, cw.cout_pad, cw.ksize, out, cout_store, x0=x0, C0=C0, x1=x1, C1=C1,
                sc=sc, sh=sh, act=act, bias=cw.bias, film=film_v, film_stride=self.film_stride if has_film else 0, residual=residual,
                scale=scale, acc_scale=cw.acc_scale, stats=stats, passes=2, corr_rel=cw.corr_rel,
                a8_exp=cw.A8_EXP, **kw))
            return
        self.steps.append(lambda: ops.conv2d_fused(
            B, H, W, cin, cw.planes, cw.cout_pad, cw.ksize, out, cout_store, x0=x0, C0=C0, x1=x1, C1=C1, sc=sc,
            sh=sh, act=act, bias=cw.bias, film=film_v, film_stride=self.film_stride if has_film else 0, residual=residual, scale=scale,
            acc_scale=cw.acc_scale, stats=stats, passes=net.plane_passes, **kw))

    def _resblock(self, rb, x: Act, skip: Act = None, mode=0, want_stats=True) -> Act:
        """ResnetBlockBigGANpp (layerspp.py:291-323).  mode 0 plain, 1 up, 2 down.

        On maps of at least 16 x 8 a plain block is TWO launches (+ two tiny table kernels): both
        GroupNorm+SiLU prologues, the channel concat and the 1x1 shortcut run inside the conv kernels.
        Up/down blocks keep one FIR pass in front (it also applies GN0+SiLU)."""
        ar, B = self.arena, self.B
        C0, C1 = x.C, (skip.C if skip is not None else 0)
        Cin, Cout, H, W = C0 + C1, rb["cout"], x.H, x.W
        assert Cin == rb["cin"], (Cin, rb["cin"])
        g0 = gn_groups(Cin)
        x0, x1 = x.t, (skip.t if skip is not None else None)
        Ho, Wo = (H * 2, W * 2) if mode == 1 else ((H // 2, W // 2) if mode == 2 else (H, W))
        gam0, bet0 = rb["gn0"]
        gam1, bet1 = rb["gn1"]
        fused = self._fusable(Ho, Wo, C0, C1, Cout)
        mask0 = 0
        if not fused and mode == 0:      # small maps: the gn_act_split in front of Conv_0 takes the sums itself
            st0, m0 = self._stats_or_defer(x)
            st1, m1 = self._stats_or_defer(skip) if skip is not None else (None, 0)
            mask0 = m0 | (m1 << 1)
        else:
            st0 = self._ensure_stats(x)
            st1 = self._ensure_stats(skip) if skip is not None else None
        h = Act(ar.f32(B, Ho, Wo, Cout), Cout, Ho, Wo, self._fused_slot(Ho, Wo, Cout))
        out_slot = self._fused_slot(Ho, Wo, Cout) if want_stats else None

        if fused and mode == 0:
            sc0, sh0 = self._tables(st0, C0, st1, C1, H * W, gam0, bet0)
            self._conv_fused(H, W, Cin, rb["conv0"], h.t, Cout, x0, C0, x1, C1, sc0, sh0, 1, film=rb["film_off"],
                             stats=h.st)
            ar.release(sc0); ar.release(sh0)
            st_h = self._ensure_stats(h)
            sc1, sh1 = self._tables(st_h, Cout, None, 0, H * W, gam1, bet1)
            out = Act(ar.f32(B, H, W, Cout), Cout, H, W, out_slot)
            if rb["has_shortcut"]:
                self._conv_fused(H, W, Cout, rb["conv1"], out.t, Cout, h.t, Cout, None, 0, sc1, sh1, 1,
                                 scale=INV_SQRT2, shortcut_raw=(x0, C0, x1, C1), stats=out.st)
            else:
                assert x1 is None and Cin == Cout
                self._conv_fused(H, W, Cout, rb["conv1"], out.t, Cout, h.t, Cout, None, 0, sc1, sh1, 1,
                                 residual=x0, scale=INV_SQRT2, stats=out.st)
            ar.release(sc1); ar.release(sh1)
            ar.release(h.t)
            return out

        if fused:
            # up / down block on a large map: one FIR pass writes a = FIR(SiLU(GN0(x))) as planes for
            # Conv_0 and FIR(x) in fp32; Conv_1 applies GN1+SiLU to h and splits FIR(x) for the
            # shortcut in-kernel (all of a launch's A patches come from ONE agent: TMA or workers)
            assert mode != 0 and x1 is None and rb["has_shortcut"]
            xr = ar.f32(B, Ho, Wo, Cin)
            e4 = self.net.passes == 2 and rb["conv0"].cout_pad >= 64
            if e4 and int(os.environ.get("DSEP_FIR_F32", "1")):
                # the FIR pass writes FIR(SiLU(GN0(x))) in fp32 and Conv_0 builds its operand planes itself (same
                # bytes through HBM as the two planes): it runs on the wide-tile fused kernel like the plain blocks
                af = ar.f32(B, Ho, Wo, Cin)
                self.steps.append(lambda: ops.fir_resample_f32(x0, B, H, W, Cin, mode, g0, st0, gam0, bet0, GN_EPS,
                                                               af, y=xr))
                self._conv_fused(Ho, Wo, Cin, rb["conv0"], h.t, Cout, af, Cin, None, 0, None, None, 0,
                                 film=rb["film_off"], stats=h.st)
                ar.release(af)
            else:
                a = ar.split(B, Ho, Wo, Cin)
                a8 = ConvWeight.A8_EXP if e4 else None
                self.steps.append(lambda: ops.fir_resample(x0, B, H, W, Cin, mode, g0, st0, gam0, bet0, GN_EPS,
                                                           a=a, y=xr, a8_exp=a8))
                self._conv(a, Ho, Wo, Cin, rb["conv0"], h.t, Cout, film=rb["film_off"], stats=h.st, e4m3=e4)
                ar.release(a)
            st_h = self._ensure_stats(h)
            sc1, sh1 = self._tables(st_h, Cout, None, 0, Ho * Wo, gam1, bet1)
            out = Act(ar.f32(B, Ho, Wo, Cout), Cout, Ho, Wo, out_slot)
            self._conv_fused(Ho, Wo, Cout, rb["conv1"], out.t, Cout, h.t, Cout, None, 0, sc1, sh1, 1,
                             scale=INV_SQRT2, shortcut_raw=(xr, Cin, None, 0), stats=out.st)
            ar.release(sc1); ar.release(sh1)
            ar.release(xr)
            ar.release(h.t)
            return out
        a = ar.split(B, Ho, Wo, Cin)
        r = ar.split(B, Ho, Wo, Cin) if rb["has_shortcut"] else None
        if mode == 0:
            self.steps.append(lambda: ops.gn_act_split(x0, C0, st0, x1, C1, st1, B, H * W, g0, gam0, bet0, GN_EPS,
                                                       1, a=a, r=r, compute_mask=mask0))
        else:
            assert x1 is None and r is not None
            self.steps.append(lambda: ops.fir_resample(x0, B, H, W, Cin, mode, g0, st0, gam0, bet0, GN_EPS,
                                                       a=a, r=r))
        self._conv(a, Ho, Wo, Cin, rb["conv0"], h.t, Cout, film=rb["film_off"], stats=h.st)
        ar.release(a)
        g1 = gn_groups(Cout)
        st_h, mh = self._stats_or_defer(h)
        a2 = ar.split(B, Ho, Wo, Cout)
        ht = h.t
        self.steps.append(lambda: ops.gn_act_split(ht, Cout, st_h, None, 0, None, B, Ho * Wo, g1, gam1, bet1,
                                                   GN_EPS, 1, a=a2, compute_mask=mh))
        # conv1 never reads h (only a2), so its buffer is reused for the block output
        out = Act(h.t, Cout, Ho, Wo, out_slot)
        if rb["has_shortcut"]:
            self._conv(a2, Ho, Wo, Cout, rb["conv1"], out.t, Cout, scale=INV_SQRT2, a2=r, stats=out.st)
            ar.release(r)
        else:
            assert x1 is None and Cin == Cout and mode == 0
            self._conv(a2, Ho, Wo, Cout, rb["conv1"], out.t, Cout, residual=x0, scale=INV_SQRT2, stats=out.st)
        ar.release(a2)
        return out

    def _attn(self, at, x: Act) -> Act:
        """AttnBlockpp (layerspp.py:76-92)."""
        ar, B, Cc, H, W = self.arena, self.B, at["c"], x.H, x.W
        S = H * W
        g = gn_groups(Cc)
        st = self._ensure_stats(x)
        gam, bet = at["gn"]
        xt = x.t
        qkv = ar.f32(B, S, 3 * Cc)
        o = ar.split(B, H, W, Cc)
        if self._fusable(H, W, Cc):     # GroupNorm (no activation) inside the stacked q/k/v projection
            sc, sh = self._tables(st, Cc, None, 0, S, gam, bet)
            self._conv_fused(H, W, Cc, at["qkv"], qkv, 3 * Cc, xt, Cc, None, 0, sc, sh, 0)
            ar.release(sc); ar.release(sh)
        else:
            a = ar.split(B, H, W, Cc)
            self.steps.append(lambda: ops.gn_act_split(xt, Cc, st, None, 0, None, B, S, g, gam, bet, GN_EPS, 0, a=a))
            self._conv(a, H, W, Cc, at["qkv"], qkv, 3 * Cc)
            ar.release(a)
        self.steps.append(lambda: ops.attention(qkv, B, S, Cc, float(Cc) ** -0.5, o))
        out = Act(ar.f32(B, H, W, Cc), Cc, H, W, self._fused_slot(H, W, Cc))
        self._conv(o, H, W, Cc, at["proj"], out.t, Cc, residual=xt, scale=INV_SQRT2, stats=out.st)
        ar.release(o)
        ar.release(qkv)
        return out

    def _build(self):
        net, ar, B, W0 = self.net, self.arena, self.B, self.W
        nf, ch_in = net.nf, net.ch_in
        nres = len(CH_MULT)
        H, W = 256, W0
        h = Act(ar.f32(B, H, W, nf), nf, H, W, self._fused_slot(H, W, nf))
        if net.conv_in_col is not None:
            col = ar.f32(B, H, W, 64)
            # (H, W are rebound below as the levels go by: bind them now)
            self.steps.append(lambda H=H, W=W, col=col: ops.im2col3x3(self.x_pyramid, B, H, W, ch_in, 64, col))
            self._conv_fused(H, W, 64, net.conv_in_col, h.t, nf, col, 64, None, 0, None, None, 0, stats=h.st)
            ar.release(col)
        else:
            self._conv(lambda: self.x_planes, H, W, net.conv_in.cin_pad, net.conv_in, h.t, nf, stats=h.st)
        hs = [h]
        pyr_in = None            # running input pyramid (fp32, ch_in channels); level 0 = x_pyramid
        for lvl, level in enumerate(net.down):
            for rb, at in zip(level["blocks"], level["attn"]):
                h = self._resblock(rb, hs[-1])
                if at is not None:
                    h2 = self._attn(at, h)
                    ar.release(h.t)
                    h = h2
                hs.append(h)
            if lvl != nres - 1:
                # the Combine below rewrites h in place, so its statistics are taken afterwards
                h = self._resblock(level["down"], hs[-1], mode=2, want_stats=False)
                new_pyr = ar.f32(B, h.H, h.W, ch_in)
                src = pyr_in
                if src is None:
                    self.steps.append(lambda Hc=H, Wc=W, new_pyr=new_pyr: ops.fir_resample(
                        self.x_pyramid, B, Hc, Wc, ch_in, 2, y=new_pyr))
                else:
                    self.steps.append(lambda src=src, Hc=H, Wc=W, new_pyr=new_pyr: ops.fir_resample(
                        src, B, Hc, Wc, ch_in, 2, y=new_pyr))
                    ar.release(src)
                pyr_in = new_pyr
                cw, cb = level["combine"]
                # maps whose consumers would otherwise launch channel_stats on the combined tensor: the sums are taken
                # here (small maps: the consuming gn_act_split computes them itself, _stats_or_defer)
                if h.st is None and h.H * h.W > 64 and int(os.environ.get("DSEP_COMBINE_STATS", "1")):
                    h.st = self._slot(h.C)
                self.steps.append(lambda pyr=new_pyr, cw=cw, cb=cb, t=h.t, P=h.H * h.W, c=h.C, st=h.st: ops.combine(
                    pyr, ch_in, cw, cb, t, t, B, P, c, stats=st))
                H, W = h.H, h.W
                hs.append(h)
        if pyr_in is not None:
            ar.release(pyr_in)

        h_mid = self._resblock(net.mid[0], hs[-1])
        h2 = self._attn(net.mid[1], h_mid)
        ar.release(h_mid.t)
        h = self._resblock(net.mid[2], h2)
        ar.release(h2.t)

        pyramid = None
        for level in net.up:
            for rb in level["blocks"]:
                skip = hs.pop()
                h_new = self._resblock(rb, h, skip)
                ar.release(h.t)
                ar.release(skip.t)
                h = h_new
            if level["attn"] is not None:
                h2 = self._attn(level["attn"], h)
                ar.release(h.t)
                h = h2
            # output pyramid: conv3x3(SiLU(GN(h))) (+ FIR-up of the running pyramid), ncsnpp.py:419-440
            g = gn_groups(h.C)
            st = self._ensure_stats(h)
            gam, bet = level["pyr_gn"]
            pyr_fused = self._fusable(h.H, h.W, h.C)
            a = None
            if not pyr_fused:
                a = ar.split(B, h.H, h.W, h.C)
                self.steps.append(lambda t=h.t, c=h.C, P=h.H * h.W, g=g, st=st, gam=gam, bet=bet, a=a:
                                  ops.gn_act_split(t, c, st, None, 0, None, B, P, g, gam, bet, GN_EPS, 1, a=a))
            new_pyr = ar.f32(B, h.H, h.W, ch_in)
            up = None
            if pyramid is not None:
                up = ar.f32(B, h.H, h.W, ch_in)
                self.steps.append(lambda src=pyramid, Hs=h.H // 2, Ws=h.W // 2, up=up: ops.fir_resample(
                    src, B, Hs, Ws, ch_in, 1, y=up))
                ar.release(pyramid)
            taps = level.get("pyr_taps") if pyr_fused else None
            if taps is not None and not (h.H % 32 == 0 and h.W % 8 == 0 and h.H * h.W >= PYR_TAPS_MIN_PIXELS):
                taps = None
            if taps is not None:
                # maps of the wide-tile kernel: 1x1 conv (GN + SiLU + split in its prologue) to the 54 tap-major
                # channels, then the 9-term gather adds bias and the up-sampled running pyramid
                zc = _round_up(9 * ch_in, 4)
                sc, sh = self._tables(st, h.C, None, 0, h.H * h.W, gam, bet)
                z = ar.f32(B, h.H, h.W, zc)
                self._conv_fused(h.H, h.W, h.C, taps, z, zc, h.t, h.C, None, 0, sc, sh, 1)
                bias = level["pyr_conv"].bias
                bias = bias[:ch_in] if bias is not None else None
                self.steps.append(lambda z=z, Hc=h.H, Wc=h.W, zc=zc, out=new_pyr, up=up, bias=bias:
                                  ops.tap_gather3x3(z, B, Hc, Wc, zc, ch_in, out, bias=bias, residual=up))
                ar.release(z)
            elif pyr_fused:     # GN + SiLU + split in the conv's own prologue (halo kernel, 16-column output tile)
                sc, sh = self._tables(st, h.C, None, 0, h.H * h.W, gam, bet)
                self._conv_fused(h.H, h.W, h.C, level["pyr_conv"], new_pyr, ch_in, h.t, h.C, None, 0, sc, sh, 1,
                                 residual=up)
            else:
                self._conv(a, h.H, h.W, h.C, level["pyr_conv"], new_pyr, ch_in, residual=up)
                ar.release(a)
            if up is not None:
                ar.release(up)
            pyramid = new_pyr
            if level["up"] is not None:
                h_new = self._resblock(level["up"], h, mode=1)
                ar.release(h.t)
                h = h_new
        assert not hs and h.H == 256 and h.W == W0
        ar.release(h.t)
        self.out = pyramid

    def run(self, x_planes, x_pyramid, t, uniform=False):
        """``uniform``: every entry of ``t`` is the same time and the caller has already put its FiLM row into
        ``self.film[0]`` (ScoreModelNCSNpp.prepare_times): the embedding MLP and the 49 Dense_0 projections — which
        depend on ``t`` only — are not re-evaluated, and the convolutions read row 0 for every batch entry."""
        net, B = self.net, self.B
        self.x_planes, self.x_pyramid = x_planes, x_pyramid
        for chunk in self._stat_chunks:
            ops.zero(chunk)
        if uniform:
            self.film_stride = 0
        else:
            self.film_stride = net._film_rows
            ops.time_embedding(t, net.Wf, net.t_w1, net.t_b1, net.t_w2, net.t_b2, B, net.nf, self.temb_act)
            ops.film(self.temb_act, net.dense_w, net.dense_b, B, net.temb_dim, net._film_rows, self.film)
        for step in self.steps:
            step()
        return self.out
