"""``ScoreModelNCSNpp``: waveform -> STFT -> NCSN++ -> iSTFT, the score network of DiffSep.

Keeps the reference's constructor arguments and ``forward(xt, time, mix)`` signature
(``models/score_models.py:10-138``); everything inside runs in libdsep kernels:

  pre_process  (:107-116)  frames (window, centre padding, right padding by n_fft-hop) -> real DFT-510
                           as an fp32 GEMM against a precomputed basis -> |S|^e e^{j arg S} * factor,
                           re/im channel stacking, frame padding to x64 and the backbone's 2x-1
                           affine, written straight as tensor-core operand planes + the fp32 pyramid
  backbone                 ``NCSNppB200``
  post_process (:118-124)  /t, output 1x1 conv, complex packing, decompression -> inverse real DFT
                           (GEMM) -> windowed overlap-add / window envelope, crop to T

The mixture's spectrogram channel does not change during sampling, so it is computed once per
mixture tensor and reused by every evaluation (the reference recomputes it 60 times per utterance).
"""
from __future__ import annotations

import contextlib
import math
import os

import torch

from . import ops
from .backbone import ConvWeight, NCSNppB200, cin_align
from .ops import Split

N_BINS = 256
LD = 512   # padded frame / spectrum row length


def n_frames(T, n_fft=510, hop=128):
    return 1 + (T + (n_fft - hop)) // hop


def _dft_bases(n_fft, device):
    """Forward basis [LD, LD]: column 2k = cos(2 pi k n / n_fft), 2k+1 = -sin(...); rows >= n_fft zero.
    Inverse basis [LD, LD]: row 2k / 2k+1 = c_k cos / -c_k sin over n, c_0 = c_{n_fft/2} = 1/n_fft,
    else 2/n_fft (one-sided irfft; imaginary parts of DC and Nyquist are ignored like torch.istft)."""
    n = torch.arange(n_fft, dtype=torch.float64)
    k = torch.arange(N_BINS, dtype=torch.float64)
    ang = 2.0 * math.pi * torch.outer(n, k) / n_fft        # [n, k]
    fwd = torch.zeros(LD, LD, dtype=torch.float64)
    fwd[:n_fft, 0::2] = torch.cos(ang)
    fwd[:n_fft, 1::2] = -torch.sin(ang)
    ck = torch.full((N_BINS,), 2.0 / n_fft, dtype=torch.float64)
    ck[0] = 1.0 / n_fft
    ck[n_fft // 2] = 1.0 / n_fft
    inv = torch.zeros(LD, LD, dtype=torch.float64)
    inv[0::2, :n_fft] = (torch.cos(ang) * ck).t()
    inv[1::2, :n_fft] = (-torch.sin(ang) * ck).t()
    inv[1, :] = 0.0
    inv[2 * (n_fft // 2) + 1, :] = 0.0
    return fwd.to(torch.float32).to(device), inv.to(torch.float32).to(device)


class ScoreModelNCSNpp(torch.nn.Module):
    """Drop-in for the reference class of the same name (inference only)."""

    def __init__(self, num_sources=2, stft_args=None, backbone_args=None, transform="exponent",
                 spec_abs_exponent=0.5, spec_factor=0.15, spec_trans_learnable=False,
                 state_dict=None, device="cuda", passes=None, **kwargs):
        super().__init__()
        stft_args = dict(stft_args or dict(n_fft=510, hop_length=128, center=True, pad_mode="constant"))
        if stft_args.get("n_fft", 510) != 510 or stft_args.get("hop_length", 128) != 128:
            raise NotImplementedError("the B200 path implements the reference's n_fft=510 / hop=128 STFT")
        if not stft_args.get("center", True) or stft_args.get("pad_mode", "constant") != "constant":
            raise NotImplementedError("only center=True, pad_mode='constant' (reference default.yaml:18-22)")
        if transform not in ("exponent", "none"):
            raise NotImplementedError(f"transform '{transform}' (reference uses 'exponent')")
        if spec_trans_learnable:
            raise NotImplementedError("spec_trans_learnable is a training feature")
        ops.use_device(device)
        ops.require_device()
        self.num_sources = num_sources
        self.n_fft, self.hop = 510, 128
        self.spec_abs_exponent = float(abs(spec_abs_exponent)) if transform == "exponent" else 1.0
        # transform == "none": neither the exponent nor the factor is applied (score_models.py:53-54, 68-69)
        self.spec_factor = float(spec_factor) if transform == "exponent" else 1.0
        backbone_args = dict(backbone_args or {})
        backbone_args.pop("_target_", None)
        self.nf = int(backbone_args.pop("nf", 128))
        if backbone_args:
            unsupported = {k: v for k, v in backbone_args.items() if k not in _BACKBONE_DEFAULTS
                           or _BACKBONE_DEFAULTS[k] != v}
            if unsupported:
                raise NotImplementedError(f"backbone options outside the hot path: {unsupported}")
        self.dev = torch.device(device)
        from . import DEFAULT_PASSES
        self.passes = DEFAULT_PASSES if passes is None else int(passes)
        self.ch_in = 2 * num_sources + 2
        self.ch_out = 2 * num_sources
        self.backbone = None
        self.window = torch.hann_window(self.n_fft, dtype=torch.float32).to(self.dev)
        self.basis_fwd, self.basis_inv = _dft_bases(self.n_fft, self.dev)
        # forward DFT-510 as a 1x1 "convolution" on the tensor core (three fp16 products, fp32-grade: the row
        # matrix [M, 512], rows padded to whole 8-row lines, is the image [1, M/8, 8, 512]) instead of the fp32 CUDA-core
        # GEMM.  Only the FORWARD product: its operand is windowed audio, O(1).  The inverse product's operand is the
        # decompressed score spectrogram ((|z| / 0.15)^2: 2e5 and more with the synthetic weights of the tests, unbounded
        # in general), beyond what an fp16 (hi, lo) pair represents (2 x 65504) — it stays an fp32 GEMM.
        self._stft_tc = bool(int(os.environ.get("DSEP_STFT_TC", "1"))) and cin_align() == 64
        self._basis_cw = {}
        if self._stft_tc:      # built now (allocations + a host sync), never inside a CUDA-graph capture
            self._basis_cw["fwd"] = ConvWeight(self.basis_fwd.t().contiguous().reshape(LD, LD, 1, 1), None, self.dev)
        self._bufs = {}
        self._mix_cache = None     # (data_ptr, B, T) of the mixture whose spectrogram is resident
        self._mix_cache_on = False
        self._uniform_t = None     # host value of the time all batch entries share (set by the sampler), or None
        self._film_cache = {}      # time -> its FiLM row [R] (device)
        self.use_cuda_graph = bool(int(os.environ.get("DSEP_CUDA_GRAPH", "1")))
        if state_dict is not None:
            self.load_state_dict(state_dict)

    # -------------------------------------------------------------- weights
    def load_state_dict(self, state_dict, strict=True):
        """Accepts the reference layout: ``backbone.*`` (+ ``stft.window``, ``stft_inv.window``)."""
        bb = {k[len("backbone."):]: v for k, v in state_dict.items() if k.startswith("backbone.")}
        if not bb:
            bb = dict(state_dict)
        if "stft.window" in state_dict:
            self.window = state_dict["stft.window"].detach().to(self.dev, torch.float32).contiguous()
        self.backbone = NCSNppB200(bb, nf=self.nf, ch_in=self.ch_in, ch_out=self.ch_out, device=self.dev,
                                   passes=self.passes)
        self._bufs = {}        # drops captured graphs that reference the old weights
        self._film_cache = {}
        return self

    @staticmethod
    def dft_rows(M):
        """rows the tensor-core form of the DFT product works on: whole 8-row image lines, at least 16 of them"""
        return max(128, (M + 7) // 8 * 8)

    def _dft(self, src, which, dst, M):
        """dst[M, LD] = src[M, LD] @ basis (``which``: "fwd" / "inv").

        Forward product: tensor-core form whenever the buffers hold ``dft_rows(M)`` rows (the model's own always do: the pad
        rows are zero and each output row depends on its own input row only), so that the arithmetic does NOT depend
        on the batch size — a shard of a batch must reproduce the whole batch's results (tests/test_graded_gpu.py,
        2-GPU gather).  Exact-size buffers of other callers fall back to the fp32 GEMM when M is not such a count."""
        basis = self.basis_fwd if which == "fwd" else self.basis_inv
        Mp = self.dft_rows(M)
        if not (which in self._basis_cw and src.shape[0] >= Mp and dst.shape[0] >= Mp):
            ops.sgemm(src, LD, basis, LD, dst, LD, M, LD, LD)
            return
        cw = self._basis_cw[which]
        ops.conv2d_fused(1, Mp // 8, 8, LD, cw.planes, cw.cout_pad, 1, dst, LD, x0=src, C0=LD, act=0,
                         acc_scale=cw.acc_scale, passes=3)

    # -------------------------------------------------------------- buffers per (B, T)
    MAX_SHAPES = 4   # distinct (batch, length) shapes kept resident (buffers, launch plan, CUDA graph)

    def _work(self, B, T):
        key = (B, T)
        if key in self._bufs:
            self._bufs[key] = self._bufs.pop(key)        # most recently used last
            return self._bufs[key]
        while len(self._bufs) >= self.MAX_SHAPES:        # ragged evaluation sets: do not hoard old shapes
            old = next(iter(self._bufs))
            del self._bufs[old]
            if self.backbone is not None:
                self.backbone.drop_plan(old[0], (n_frames(old[1]) + 63) // 64 * 64,
                                        keep={(b, (n_frames(t) + 63) // 64 * 64) for b, t in self._bufs})
        dev, ns = self.dev, self.num_sources
        Fr = n_frames(T)
        Wp = (Fr + 63) // 64 * 64
        f32 = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)
        rows = lambda m: torch.zeros(self.dft_rows(m), LD, device=dev, dtype=torch.float32)   # operands of _dft
        b = dict(
            Fr=Fr, Wp=Wp,
            frames=rows(B * ns * Fr), dft=rows(B * ns * Fr),
            frames_mix=rows(B * Fr), dft_mix=rows(B * Fr),
            # operand planes of the input conv — not needed when it runs over im2col rows of x_pyr
            x_planes=(None if self.backbone.conv_in_col is not None
                      else Split.zeros((B, N_BINS, Wp, self.backbone.conv_in.cin_pad), dev)),
            x_pyr=f32(B, N_BINS, Wp, self.ch_in),
            spec_out=rows(B * ns * Fr), frames_out=rows(B * ns * Fr),
            score=f32(B, ns, T),
        )
        self._bufs[key] = b
        return b

    @contextlib.contextmanager
    def cached_mixture(self, mix):
        """Within the context, evaluations with this very ``mix`` tensor (same storage) reuse its
        spectrogram channel instead of recomputing it; the sampler wraps its loop in this."""
        prev = (self._mix_cache_on, self._mix_cache)
        self._mix_cache_on, self._mix_cache = True, None
        try:
            yield self
        finally:
            self._mix_cache_on, self._mix_cache = prev

    # -------------------------------------------------------------- time embedding hoisted out of the loop
    def prepare_times(self, ts):
        """The Fourier embedding, its two Linear layers and the 49 ``Dense_0`` FiLM projections depend on ``t`` only
        (ncsnpp.py:324-343, layerspp.py:311-313) and the sampler's time grid is known up front: evaluate them for the
        whole grid in one go instead of once per score evaluation."""
        if self.backbone is None:
            return
        new = [float(t) for t in ts if float(t) not in self._film_cache]
        if new:
            rows = self.backbone.film_rows(new)
            for i, t in enumerate(new):
                self._film_cache[t] = rows[i]
            while len(self._film_cache) > 4096:
                self._film_cache.pop(next(iter(self._film_cache)))

    @contextlib.contextmanager
    def uniform_time(self, t):
        """Within the context the caller guarantees that every entry of the ``time`` argument equals ``t`` (the PC
        sampler: sdes/__init__.py:177-178), so the cached FiLM row of ``t`` serves the whole batch."""
        prev = self._uniform_t
        self._uniform_t = float(t)
        try:
            yield self
        finally:
            self._uniform_t = prev

    def _film_row(self):
        if self._uniform_t is None:
            return None
        if self._uniform_t not in self._film_cache:
            self.prepare_times([self._uniform_t])
        return self._film_cache[self._uniform_t]

    # -------------------------------------------------------------- forward
    @torch.no_grad()
    def forward(self, xt, time, mix):
        """xt: [B, n_src, T], time: [B], mix: [B, 1, T] (CUDA fp32) -> score [B, n_src, T]."""
        if self.backbone is None:
            raise RuntimeError("ScoreModelNCSNpp has no weights: call load_state_dict first")
        if xt.dim() != 3 or mix.dim() != 3 or xt.shape[1] != self.num_sources or mix.shape[1] != 1:
            raise ValueError(f"expected xt [B,{self.num_sources},T] and mix [B,1,T], got {tuple(xt.shape)}, "
                             f"{tuple(mix.shape)}")
        B, ns, T = xt.shape
        if mix.shape[0] != B or mix.shape[2] != T or time.shape != (B,):
            raise ValueError("xt, time and mix disagree on batch size or length")
        xt = xt.contiguous().float()
        mix = mix.contiguous().float()
        time = time.contiguous().float()
        bf = self._work(B, T)
        Fr, Wp = bf["Fr"], bf["Wp"]

        mix_key = (mix.data_ptr(), B, T)
        if self._mix_cache_on and self._mix_cache == mix_key:
            if self.use_cuda_graph:
                return self._replay(xt, time, bf, self._film_row())
        else:
            ops.stft_frames(mix, self.window, B, 1, T, Fr, bf["frames_mix"])
            self._dft(bf["frames_mix"], "fwd", bf["dft_mix"], B * Fr)
            ops.spec_pack(bf["dft_mix"], B, 1, Fr, Wp, ns, ns + 1, self.backbone.conv_in.cin_pad, self.spec_factor,
                          self.spec_abs_exponent,
                          bf["x_pyr"], bf["x_planes"])
            self._mix_cache = mix_key if self._mix_cache_on else None
        return self._evaluate(xt, time, bf, self._film_row())

    def _evaluate(self, xt, time, bf, film_row=None):
        """Everything that depends on xt / t: the launch sequence one CUDA graph captures."""
        B, ns, T = xt.shape
        Fr, Wp = bf["Fr"], bf["Wp"]
        ops.stft_frames(xt, self.window, B, ns, T, Fr, bf["frames"])
        self._dft(bf["frames"], "fwd", bf["dft"], B * ns * Fr)
        ops.spec_pack(bf["dft"], B, ns, Fr, Wp, 0, ns + 1, self.backbone.conv_in.cin_pad, self.spec_factor,
                      self.spec_abs_exponent,
                      bf["x_pyr"], bf["x_planes"])
        if film_row is not None:      # one row for the whole batch (uniform time); a no-op when replay already set it
            self.backbone.plan(B, Wp).film[0].copy_(film_row)
        pyr = self.backbone(bf["x_planes"], bf["x_pyr"], time, uniform=film_row is not None)
        ops.out_head(pyr, B, Wp, self.ch_in, ns, Fr, time, self.backbone.out_w, self.backbone.out_b,
                     self.spec_factor, self.spec_abs_exponent, bf["spec_out"])
        self._dft(bf["spec_out"], "inv", bf["frames_out"], B * ns * Fr)
        out = torch.empty(B, ns, T, device=self.dev, dtype=torch.float32)
        ops.istft_ola(bf["frames_out"], self.window, B, ns, Fr, T, out)
        return out

    def _replay(self, xt, time, bf, film_row=None):
        """Inside a sampling run (mixture spectrogram resident) the ~300 launches of one evaluation
        are replayed from a CUDA graph: static input buffers, one graph launch per evaluation.  With a uniform time
        the graph holds no time-embedding / FiLM launches: it reads the row staged here (one small copy)."""
        from . import _lib
        key = "graph_u" if film_row is not None else "graph"
        g = bf.get(key)
        if g is None:
            B, ns, T = xt.shape
            xs = torch.empty_like(xt)
            ts = torch.empty_like(time)
            row = torch.empty_like(film_row) if film_row is not None else None
            xs.copy_(xt); ts.copy_(time)
            if row is not None:
                row.copy_(film_row)
            self._evaluate(xs, ts, bf, row)            # warm-up outside capture: plans, attributes
            torch.cuda.current_stream().synchronize()
            graph = torch.cuda.CUDAGraph()
            n0 = _lib.N_CALLS
            with torch.cuda.graph(graph):
                out = self._evaluate(xs, ts, bf, row)
            g = bf[key] = (graph, xs, ts, row, out, _lib.N_CALLS - n0)
        graph, xs, ts, row, out, n_launches = g
        xs.copy_(xt)
        ts.copy_(time)
        if row is not None:
            row.copy_(film_row)
        graph.replay()
        _lib.N_CALLS += n_launches
        return out.clone()


# NCSNpp constructor defaults (reference models/ncsnpp.py:45-70): anything else is not on the hot path
_BACKBONE_DEFAULTS = dict(
    scale_by_sigma=True, nonlinearity="swish", ch_mult=(1, 1, 2, 2, 2, 2, 2), num_res_blocks=2,
    attn_resolutions=(16,), resamp_with_conv=True, conditional=True, fir=True, fir_kernel=[1, 3, 3, 1],
    skip_rescale=True, resblock_type="biggan", progressive="output_skip", progressive_input="input_skip",
    progressive_combine="sum", init_scale=0.0, fourier_scale=16, image_size=256, embedding_type="fourier",
    dropout=0.0, centered=False,
)
