"""Tensor-level wrappers over the C-ABI (``_lib``): each takes CUDA tensors the caller allocated,
checks layout on the host, and launches on torch's current stream.  PyTorch is only the
allocator / stream provider here; no torch operator computes anything on this path.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import SdeParams, call


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    if t is None:
        return None
    if not t.is_cuda:
        raise ValueError("dsep ops need CUDA tensors (there is no CPU path)")
    if not t.is_contiguous():
        raise ValueError("dsep ops need contiguous tensors")
    return t.data_ptr()


def use_device(device):
    """Makes ``device`` (``"cuda:1"``, ``1``, ``torch.device``) the current CUDA device: every kernel launches on
    ``torch.cuda.current_stream()`` of the CURRENT device, so a model living on cuda:1 while cuda:0 is current
    would launch on the wrong GPU (the reference CLI honours ``-d`` the same way through ``.to(device)``)."""
    if not torch.cuda.is_available():
        return
    dev = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
    if dev.type == "cuda" and dev.index is not None and dev.index != torch.cuda.current_device():
        torch.cuda.set_device(dev)


def require_device():
    if not torch.cuda.is_available():
        raise RuntimeError("diffsep_b200 needs a CUDA device (B200, sm_100); no CPU fallback exists")
    if not _lib.load().dsep_device_ok():
        raise RuntimeError("diffsep_b200 kernels are built for sm_100a (B200) only")


def _f32(t, name):
    if t is not None and t.dtype != torch.float32:
        raise ValueError(f"{name} must be float32, got {t.dtype}")
    return t


class Split:
    """A pair of fp16 planes (hi, lo) with hi + lo == x: a tensor-core operand."""

    __slots__ = ("hi", "lo")

    def __init__(self, hi, lo):
        self.hi, self.lo = hi, lo

    @staticmethod
    def empty(shape, device):
        return Split(torch.empty(shape, dtype=torch.float16, device=device),
                     torch.empty(shape, dtype=torch.float16, device=device))

    @staticmethod
    def zeros(shape, device):
        return Split(torch.zeros(shape, dtype=torch.float16, device=device),
                     torch.zeros(shape, dtype=torch.float16, device=device))

    @property
    def shape(self):
        return self.hi.shape


def split_f16(x, out: Split, prescale=1.0):
    _f32(x, "x")
    call("dsep_split_f16", ptr(x), x.numel(), prescale, ptr(out.hi), ptr(out.lo), stream())
    return out


def conv2d_tc(a: Split, B, H, W, Cin, w: Split, cout_pad, ksize, out, cout_store, bias=None, film=None,
              film_stride=0, residual=None, scale=1.0, acc_scale=1.0, passes=3, a2: Split = None, Cin2=0,
              w2: Split = None, stats=None, corr_rel=1.0, a8_exp=0):
    """passes = 2: ``a.lo`` / ``w.lo`` are the e4m3 correction planes (dsep_conv2d_fused8 fed with planes)."""
    _f32(out, "out"); _f32(bias, "bias"); _f32(film, "film"); _f32(residual, "residual")
    if stats is not None and stats.dtype != torch.float64:
        raise ValueError("stats must be float64 [B, C, 2]")
    if passes == 2:
        return conv2d_fused(B, H, W, Cin, w, cout_pad, ksize, out, cout_store, a=a, a2=a2, Cin2=Cin2, w2=w2, bias=bias,
                            film=film, film_stride=film_stride, residual=residual, scale=scale, acc_scale=acc_scale,
                            stats=stats, passes=2, corr_rel=corr_rel, a8_exp=a8_exp)
    call("dsep_conv2d_tc", ptr(a.hi), ptr(a.lo), B, H, W, Cin, ptr(w.hi), ptr(w.lo), cout_pad, ksize,
         ptr(a2.hi) if a2 else None, ptr(a2.lo) if a2 else None, Cin2, ptr(w2.hi) if w2 else None,
         ptr(w2.lo) if w2 else None, ptr(bias), film.data_ptr() if film is not None else None, film_stride,
         ptr(residual), scale, acc_scale, ptr(out), cout_store, ptr(stats), passes, stream())
    return out


def conv2d_fused(B, H, W, Cin, w: Split, cout_pad, ksize, out, cout_store, a: Split = None, x0=None, C0=0, x1=None,
                 C1=0, sc=None, sh=None, act=0, a2: Split = None, s0=None, S0=0, s1=None, S1=0, Cin2=0,
                 w2: Split = None, bias=None, film=None, film_stride=0, residual=None, scale=1.0, acc_scale=1.0,
                 stats=None, passes=3, corr_rel=1.0, a8_exp=0):
    """dsep_conv2d_fused: the convolution with GroupNorm-apply / SiLU / concat / split of its operands
    done in-kernel (x0[, x1] with sc, sh) instead of arriving as split planes (a).  passes = 2
    (experimental -DDSEP_FP8_CORR build only): dsep_conv2d_fused8, ``w.lo`` is the e4m3 correction plane."""
    _f32(out, "out"); _f32(bias, "bias"); _f32(film, "film"); _f32(residual, "residual")
    _f32(x0, "x0"); _f32(x1, "x1"); _f32(s0, "s0"); _f32(s1, "s1"); _f32(sc, "sc"); _f32(sh, "sh")
    g = _lib.ConvArgs(
        ptr(a.hi) if a else None, ptr(a.lo) if a else None, ptr(x0), ptr(x1), C0, C1, ptr(sc), ptr(sh), act,
        B, H, W, Cin, ptr(w.hi), ptr(w.lo), cout_pad, ksize, ptr(a2.hi) if a2 else None,
        ptr(a2.lo) if a2 else None, ptr(s0), ptr(s1), S0, S1, Cin2, ptr(w2.hi) if w2 else None,
        ptr(w2.lo) if w2 else None, ptr(bias), film.data_ptr() if film is not None else None, film_stride,
        ptr(residual), scale, acc_scale, ptr(out), cout_store, ptr(stats), passes)
    try:
        if passes == 2:
            call("dsep_conv2d_fused8", C.byref(g), corr_rel, a8_exp, stream())
        else:
            call("dsep_conv2d_fused", C.byref(g), stream())
    except NotImplementedError:      # a RuntimeError subclass: keep the reference's exception type
        raise
    except RuntimeError as e:
        desc = {k: getattr(g, k) for k, t in g._fields_ if t is C.c_int or t is C.c_float}
        desc.update({k: bool(getattr(g, k)) for k, t in g._fields_ if t is C.c_void_p})
        raise RuntimeError(f"{e} :: {desc}") from e
    return out


def gn_tables(st0, C0, st1, C1, B, P, groups, gamma, beta, eps, sc, sh):
    call("dsep_gn_tables", ptr(st0), C0, ptr(st1), C1, B, P, groups, ptr(gamma), ptr(beta), eps, ptr(sc), ptr(sh),
         stream())


def channel_stats(x, Cc, B, P, stats):
    if stats.dtype != torch.float64:
        raise ValueError("stats must be float64 [B, C, 2]")
    call("dsep_channel_stats", ptr(_f32(x, "x")), Cc, B, P, ptr(stats), stream())
    return stats


def zero(t):
    call("dsep_zero", ptr(t), t.numel() * t.element_size(), stream())
    return t


def gn_act_split(x0, C0, st0, x1, C1, st1, B, P, groups, gamma, beta, eps, act, a: Split = None, r: Split = None,
                 compute_mask=0):
    """compute_mask (small maps): bit 0 / 1 = compute the per-channel sums of x0 / x1 in the kernel and write them to
    st0 / st1 first (dsep_gn_stats_act_split) instead of reading them."""
    if compute_mask:
        call("dsep_gn_stats_act_split", ptr(x0), C0, ptr(st0), ptr(x1), C1, ptr(st1), B, P, groups, ptr(gamma),
             ptr(beta), eps, act, ptr(a.hi) if a else None, ptr(a.lo) if a else None, ptr(r.hi) if r else None,
             ptr(r.lo) if r else None, compute_mask, stream())
        return
    call("dsep_gn_act_split", ptr(x0), C0, ptr(st0), ptr(x1), C1, ptr(st1), B, P, groups, ptr(gamma), ptr(beta),
         eps, act, ptr(a.hi) if a else None, ptr(a.lo) if a else None, ptr(r.hi) if r else None,
         ptr(r.lo) if r else None, stream())


def fir_resample(x, B, H, W, Cc, mode, groups=0, stats=None, gamma=None, beta=None, eps=1e-6, a: Split = None,
                 r: Split = None, y=None, a8_exp=None):
    """a8_exp given: ``a.lo`` receives the e4m3 correction plane of the passes = 2 convolution instead of fp16 lo."""
    if a8_exp is not None:
        call("dsep_fir_resample8", ptr(_f32(x, "x")), B, H, W, Cc, mode, groups, ptr(stats), ptr(gamma), ptr(beta),
             eps, ptr(a.hi), ptr(a.lo), ptr(r.hi) if r else None, ptr(r.lo) if r else None, ptr(y), a8_exp, stream())
        return
    call("dsep_fir_resample", ptr(_f32(x, "x")), B, H, W, Cc, mode, groups, ptr(stats), ptr(gamma), ptr(beta),
         eps, ptr(a.hi) if a else None, ptr(a.lo) if a else None, ptr(r.hi) if r else None,
         ptr(r.lo) if r else None, ptr(y), stream())


def fir_resample_f32(x, B, H, W, Cc, mode, groups, stats, gamma, beta, eps, af, y=None):
    """af = FIR(SiLU(GN(x))) in fp32 (+ y = FIR(x)): dsep_fir_resample_f32"""
    call("dsep_fir_resample_f32", ptr(_f32(x, "x")), B, H, W, Cc, mode, groups, ptr(stats), ptr(gamma), ptr(beta),
         eps, ptr(_f32(af, "af")), ptr(y), stream())


def upfirdn2d_planes(x, planes, H, W, up, down, pad, out):
    call("dsep_upfirdn2d", ptr(_f32(x, "x")), planes, H, W, up, up, down, down, pad[0], pad[1], pad[0], pad[1],
         ptr(out), stream())
    return out


def combine(pyr, Cp, w, bias, h, out, B, P, Cc, stats=None):
    """``stats`` (fp64 [B, Cc, 2], zeroed by the caller): the per-channel sums of ``out`` are added to it"""
    if stats is not None and (stats.dtype != torch.float64 or stats.numel() != B * Cc * 2):
        raise ValueError("combine: stats must be float64 [B, C, 2]")
    call("dsep_combine", ptr(pyr), Cp, ptr(w), ptr(bias), ptr(h), ptr(out), B, P, Cc, ptr(stats), stream())
    return out


def im2col3x3(x, B, H, W, Cc, Cp, col):
    call("dsep_im2col3x3", ptr(_f32(x, "x")), B, H, W, Cc, Cp, ptr(_f32(col, "col")), stream())
    return col


def tap_gather3x3(z, B, H, W, ZC, CO, out, bias=None, residual=None):
    """out = bias + residual + sum over the 9 taps of z[neighbour][tap * CO + co] (dsep_tap_gather3x3)"""
    call("dsep_tap_gather3x3", ptr(_f32(z, "z")), B, H, W, ZC, CO, ptr(_f32(bias, "bias")) if bias is not None else None,
         ptr(_f32(residual, "residual")) if residual is not None else None, ptr(_f32(out, "out")), stream())
    return out


def add(a, b, y):
    call("dsep_add", ptr(a), ptr(b), ptr(y), a.numel(), stream())
    return y


def attention(qkv, B, S, Cc, scale, o: Split):
    call("dsep_attention", ptr(_f32(qkv, "qkv")), B, S, Cc, scale, ptr(o.hi), ptr(o.lo), stream())


def time_embedding(t, Wf, w1, b1, w2, b2, B, nf, out):
    call("dsep_time_embedding", ptr(_f32(t, "t")), ptr(Wf), ptr(w1), ptr(b1), ptr(w2), ptr(b2), B, nf,
         ptr(out), stream())
    return out


def film(temb_act, Wd, bd, B, D, R, out):
    call("dsep_film", ptr(temb_act), ptr(Wd), ptr(bd), B, D, R, ptr(out), stream())
    return out


def stft_frames(x, window, B, Cc, T, Fr, frames):
    call("dsep_stft_frames", ptr(_f32(x, "x")), ptr(window), B, Cc, T, Fr, ptr(frames), stream())
    return frames


def sgemm(A, lda, Bm, ldb, Cm, ldc, M, N, K):
    call("dsep_sgemm", ptr(A), lda, ptr(Bm), ldb, ptr(Cm), ldc, M, N, K, stream())
    return Cm


def spec_pack(dft, B, Cw, Fr, Wp, chan0, Ctot, Cpad, factor, exponent, x_f32, a: Split):
    call("dsep_spec_pack", ptr(dft), B, Cw, Fr, Wp, chan0, Ctot, Cpad, factor, exponent, ptr(x_f32),
         ptr(a.hi) if a else None, ptr(a.lo) if a else None, stream())


def out_head(pyr, B, Wp, Cp, nsrc, Fr, t, w, bias, factor, exponent, spec):
    call("dsep_out_head", ptr(pyr), B, Wp, Cp, nsrc, Fr, ptr(_f32(t, "t")), ptr(w), ptr(bias), factor, exponent,
         ptr(spec), stream())
    return spec


def istft_ola(frames_t, window, B, Cc, Fr, T, out):
    call("dsep_istft_ola", ptr(frames_t), ptr(window), B, Cc, Fr, T, ptr(out), stream())
    return out


def sde_params(d_lambda, sigma_min, sigma_max, T_end=1.0, ndim=2):
    return SdeParams(float(d_lambda), float(sigma_min), float(sigma_max), float(T_end), int(ndim))


def sde_prior(p, mix, sigma_mix, noise, seed, offset, B, T, x, mix_channels=1, mean_scale=0.5, sigma_channels=1):
    call("dsep_sde_prior", C.byref(p), ptr(_f32(mix, "mix")), mix_channels, mean_scale, ptr(sigma_mix),
         sigma_channels, ptr(_f32(noise, "noise")), seed, offset, B, T, ptr(x), stream())
    return x


def sde_corrector(p, x, score, t, sigma_mix, noise, seed, offset, snr, B, T, x_out, x_mean):
    call("dsep_sde_corrector", C.byref(p), ptr(_f32(x, "x")), ptr(_f32(score, "score")), ptr(_f32(t, "t")),
         ptr(sigma_mix), ptr(_f32(noise, "noise")), seed, offset, snr, B, T, ptr(x_out), ptr(x_mean), stream())


def sde_predictor(p, x, score, t, sigma_mix, noise, seed, offset, dt, B, T, x_out, x_mean, probability_flow=False):
    call("dsep_sde_predictor", C.byref(p), ptr(_f32(x, "x")), ptr(_f32(score, "score")), ptr(_f32(t, "t")),
         ptr(sigma_mix), ptr(_f32(noise, "noise")), seed, offset, dt, int(bool(probability_flow)), B, T, ptr(x_out),
         ptr(x_mean), stream())


def sde_perturb(p, x0, t, sigma_mix, noise, seed, offset, B, T, x_t, z_out=None):
    call("dsep_sde_perturb", C.byref(p), ptr(_f32(x0, "x0")), ptr(_f32(t, "t")), ptr(sigma_mix),
         ptr(_f32(noise, "noise")), seed, offset, B, T, ptr(x_t), ptr(z_out), stream())
    return x_t


def score_loss(p, score, z, t, sigma_mix, B, T, loss):
    if loss.dtype != torch.float64:
        raise ValueError("loss must be float64 [B]")
    call("dsep_score_loss", C.byref(p), ptr(_f32(score, "score")), ptr(_f32(z, "z")), ptr(_f32(t, "t")),
         ptr(sigma_mix), B, T, ptr(loss), stream())
    return loss


def sde_corrector_ald(p, x, score, t, noise, seed, offset, snr, B, T, x_out, x_mean):
    call("dsep_sde_corrector_ald", C.byref(p), ptr(_f32(x, "x")), ptr(_f32(score, "score")), ptr(_f32(t, "t")),
         ptr(_f32(noise, "noise")), seed, offset, snr, B, T, ptr(x_out), ptr(x_mean), stream())


def sde_corrector_langevin(x, score, noise, snr, B, n, norms, x_out, x_mean):
    call("dsep_sde_corrector_langevin", ptr(_f32(x, "x")), ptr(_f32(score, "score")), ptr(_f32(noise, "noise")), snr,
         B, n, ptr(norms), ptr(x_out), ptr(x_mean), stream())


def sigma_mix(mix, B, T, avg_len, out):
    call("dsep_sigma_mix", ptr(_f32(mix, "mix")), B, T, avg_len, ptr(out), stream())
    return out


def normalize(mix, B, n, out, mean=None, std=None):
    call("dsep_normalize", ptr(_f32(mix, "mix")), B, n, ptr(out), ptr(mean), ptr(std), stream())
    return out


def scale_output(mix, sep, B, nsrc, T, out):
    call("dsep_scale_output", ptr(_f32(mix, "mix")), ptr(_f32(sep, "sep")), B, nsrc, T, ptr(out), stream())
    return out


def randn(z, seed, offset):
    call("dsep_randn", ptr(_f32(z, "z")), z.numel(), seed, offset, stream())
    return z
