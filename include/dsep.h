/*
 * libdsep — C-ABI of the B200-native DiffSep reverse-diffusion hot path.
 *
 * Every entry point takes plain device pointers, sizes and a CUDA stream (as void*),
 * launches asynchronously on that stream, allocates nothing that outlives the call, never
 * synchronises, and returns 0 or a negative DSEP_ERR_* code (message via dsep_last_error()).
 * Outputs are caller-allocated.  No torch types cross this boundary.
 *
 * What each function replaces in the reference (fakufaku/diffusion-separation @ d1855e9) is
 * cited as file:line.  The reference's only native FFI on this path is the pybind11 op
 *   upfirdn2d(Tensor in, Tensor kernel, up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1)
 * (models/ncsnpp_utils/op/upfirdn2d.cpp:12-22); everything else it delegates to
 * cuDNN/cuBLAS/cuFFT through PyTorch.  Those library calls are what the remaining entry points
 * stand in for.  INTEGRATION.md shows the ctypes/pybind stubs a maintainer would add.
 *
 * Layouts: activations are channels-last, fp32 [B, H, W, C]; tensor-core operands are a pair of
 * fp16 planes (hi, lo) with hi + lo == x to 2^-22 ("split" tensors) in the same [B, H, W, C]
 * layout; waveforms are [B, C, T] fp32.
 */
#ifndef DSEP_H_
#define DSEP_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DSEP_OK 0
#define DSEP_ERR_INVALID (-1)     /* bad argument (the Python wrappers raise ValueError)      */
#define DSEP_ERR_CUDA (-2)        /* CUDA runtime/driver error (wrappers raise RuntimeError)  */
#define DSEP_ERR_UNSUPPORTED (-3) /* valid in the reference but outside the hot path's shapes */

#define DSEP_ABI_VERSION 8

typedef void* dsep_stream_t; /* cudaStream_t */

const char* dsep_last_error(void);
int dsep_abi_version(void);
/* hash of the sources + compiler flags this binary was built from (diffsep_b200/build.py: source_hash()) */
const char* dsep_source_hash(void);
/* 1 if the running device is sm_100 (B200); the product path refuses anything else. */
int dsep_device_ok(void);
/* channel granularity of dsep_conv2d_tc's operands (Cin, Cin2 must be multiples of it): 64 */
int dsep_conv_kblock(void);

/* ---- tensor-core convolution -------------------------------------------------------------
 * out[b,h,w,n] = scale * ( acc_scale * ( sum_{tap,c} A[b,h+dy,w+dx,c] * Wt[tap,n,c]
 *                                        + sum_c A2[b,h,w,c] * W2[n,c] )
 *                          + bias[n] + film[b,n] + residual[b,h,w,n] )
 * stats[b,n,0..1] += (sum, sum of squares) of out[b,:,:,n]                       (if stats != NULL)
 * Implicit GEMM on tcgen05 (TMA-fed, TMEM accumulators).  ksize 3 (pad 1) or 1.  A is a split
 * tensor [B,H,W,Cin] (Cin % dsep_conv_kblock() == 0); Wt a split tensor [ksize*ksize, Cout_pad, Cin] with
 * Cout_pad in {16} or a multiple of 64; only the first cout_store channels are written, with
 * row pitch cout_store.  A2 [B,H,W,Cin2] / W2 [Cout_pad, Cin2] (Cin2 = 0: none) is a fused 1x1
 * convolution accumulated into the same tile (the ResBlock shortcut Conv_2).
 * passes = 3: hi*hi + lo*hi + hi*lo (fp32-grade); passes = 1: hi*hi (11-bit operands, TF32-grade).
 * acc_scale undoes the power-of-two pre-scaling that keeps the fp16 weight planes in the normal
 * range (Wt and W2 hold w / acc_scale).  bias/film/residual may be NULL.  film is [B, film_stride]
 * (pointer already offset to this layer's first channel).  stats is double [B, cout_store, 2],
 * accumulated atomically (zero it first; needs Cout >= 64 and a map of at least 128 pixels).
 * Replaces nn.Conv2d -> cuDNN in ddpm_conv3x3/ddpm_conv1x1 (models/ncsnpp_utils/layers.py:112-156),
 * NIN (layers.py:678-689), Dense_0 bias add, Conv_2 shortcut and the (x+h)/sqrt(2) residual
 * (layerspp.py:311-323), and the reduction half of the following nn.GroupNorm. */
int dsep_conv2d_tc(const void* a_hi, const void* a_lo, int B, int H, int W, int Cin,
                   const void* w_hi, const void* w_lo, int Cout_pad, int ksize,
                   const void* a2_hi, const void* a2_lo, int Cin2, const void* w2_hi, const void* w2_lo,
                   const float* bias, const float* film, int film_stride, const float* residual,
                   float scale, float acc_scale, float* out, int cout_store, double* stats, int passes,
                   dsep_stream_t stream);

/* Same convolution with the operand-producing passes folded in (the north-star fusion: one ResBlock
 * = two launches).  Any field left NULL/0 selects the dsep_conv2d_tc behaviour for that operand.
 *   main operand:     x0 != NULL  =>  A = act(x * sc[b,c] + sh[b,c]) of the fp32, channel-concatenated
 *                     [x0 (C0 ch) | x1 (C1 ch)] (C0 + C1 == Cin), built in-kernel; sc/sh [B, Cin] come
 *                     from dsep_gn_tables (NULL: identity), act 0 none / 1 SiLU; zero outside the image
 *                     (the convolution pads the ACTIVATED tensor).  Replaces nn.GroupNorm + nn.SiLU +
 *                     torch.cat in front of Conv_0 / Conv_1 / NIN_0..2 (layerspp.py:292-309, 76-82).
 *   shortcut operand: s0 != NULL  =>  A2 = the raw fp32 [s0 (S0 ch) | s1 (S1 ch)] (S0 + S1 == Cin2).
 * Needs a map of at least 16 x 8 pixels; Cout_pad = 16 (the output-pyramid convs) is accepted without a fused
 * shortcut and without statistics (smaller maps: dsep_gn_act_split + dsep_conv2d_tc). */
typedef struct {
    const void *a_hi, *a_lo;            /* split planes [B,H,W,Cin], or NULL with x0 set            */
    const float *x0, *x1;               /* fp32 activations [B,H,W,C0], [B,H,W,C1]                  */
    int C0, C1;
    const float *sc, *sh;               /* [B, Cin] affine of the in-kernel prologue                */
    int act;
    int B, H, W, Cin;
    const void *w_hi, *w_lo;            /* [ksize*ksize, Cout_pad, Cin]                             */
    int Cout_pad, ksize;
    const void *a2_hi, *a2_lo;          /* shortcut planes [B,H,W,Cin2], or NULL with s0 set        */
    const float *s0, *s1;
    int S0, S1;
    int Cin2;
    const void *w2_hi, *w2_lo;          /* [Cout_pad, Cin2]                                         */
    const float *bias, *film;
    int film_stride;
    const float* residual;
    float scale, acc_scale;
    float* out;
    int cout_store;
    double* stats;
    int passes;
} dsep_conv_args;
int dsep_conv2d_fused(const dsep_conv_args* args, dsep_stream_t stream);
/* The product's default mode, passes = 2 (dsep_has_fp8_corr() == 1 unless built with -DDSEP_FP8_CORR=0, where this
 * entry point returns DSEP_ERR_UNSUPPORTED): the same convolution with per K = 16 step one fp16 product hi*hi plus
 * ONE e4m3 tensor-core product carrying both correction terms (2 tensor-core units per MAC instead of 3;
 * tools/numerics_study.py, profiles/parity_r02.md).  args->w_lo then points to the e4m3 weight plane
 * [taps, Cout_pad, 2*Cin] bytes: per 8 input channels the 16 bytes [W_hi8 x 8 | W_lo8 x 8], W_hi8 = e4m3(W_hi * 2^q),
 * W_lo8 = e4m3(W_lo * 2^(q+11)) on top of the fp16 planes' prescale; activations are prescaled by 2^a8_exp (hi) and
 * 2^(a8_exp+11) (lo) — in-kernel when x0 is given, otherwise a_lo must already be the e4m3 activation plane
 * [B,H,W,2*Cin] bytes, per 8 channels [A_lo8 x 8 | A_hi8 x 8] (dsep_fir_resample8 writes it); corr_rel =
 * 2^-(q + a8_exp + 11) weighs the correction accumulator.  A fused 1x1 shortcut keeps fp16 (hi, lo) planes.
 * Needs Cout >= 64 and a map of at least 16 x 8 (the halo kernel). */
int dsep_conv2d_fused8(const dsep_conv_args* args, float corr_rel, int a8_exp, dsep_stream_t stream);
int dsep_has_fp8_corr(void);
/* How many of the dsep_conv2d_fused8 calls of this process took the wide-tile kernel (conv_wide.cu: with corr_rel == 1,
 * x0 given, Cout_pad % 128 == 0, H % 32 == 0 and W % 8 == 0 the same convolution runs on 8 x 32 pixel tiles with the
 * pixels on the tensor core's N side).  A counter for tests and profiles, not part of the data path. */
int dsep_conv_wide_launches(void);
/* Per-(batch entry, channel) GroupNorm scale / shift from per-channel sums of a (concatenated) input:
 * sc = gamma * rstd[group], sh = beta - mean[group] * sc, so that GN(x) = x * sc + sh.  sc, sh: [B, C0+C1]. */
int dsep_gn_tables(const double* st0, int C0, const double* st1, int C1, int B, int P, int groups,
                   const float* gamma, const float* beta, float eps, float* sc, float* sh,
                   dsep_stream_t stream);

/* x * prescale (fp32, n elements) -> split fp16 planes. */
int dsep_split_f16(const float* x, int64_t n, float prescale, void* hi, void* lo,
                   dsep_stream_t stream);

/* ---- GroupNorm / SiLU / FIR resampling ------------------------------------------------------
 * GroupNorm statistics travel as per-channel sums: double [B, C, 2] = (sum, sum of squares) over the
 * P = H*W pixels, produced by dsep_conv2d_tc's epilogue or by dsep_channel_stats; consumers
 * combine them per group, so a channel-concatenated input [x0 (C0 ch) | x1 (C1 ch)] (x1 may be
 * NULL, C1 = 0) needs no pass of its own.
 * Replaces nn.GroupNorm(min(C//4,32), eps=1e-6) + nn.SiLU (layerspp.py:264-266,292; ncsnpp.py:253-258)
 * and torch.cat([h, hs.pop()]) (ncsnpp.py:411). */
int dsep_channel_stats(const float* x, int C, int B, int P, double* stats, dsep_stream_t stream);
/* memset(ptr, 0, bytes) on the stream (zeroing the statistics arena before an evaluation). */
int dsep_zero(void* ptr, int64_t bytes, dsep_stream_t stream);
/* a = act(GN(x)) as split planes (act: 0 none, 1 SiLU); optionally also the raw x as split
 * planes (r_hi/r_lo, for the 1x1 shortcut conv) — all in the concatenated channel layout.
 * st0/st1 NULL => no GroupNorm (plain split). */
int dsep_gn_act_split(const float* x0, int C0, const double* st0, const float* x1, int C1,
                      const double* st1, int B, int P, int groups, const float* gamma, const float* beta,
                      float eps, int act, void* a_hi, void* a_lo, void* r_hi, void* r_lo,
                      dsep_stream_t stream);
/* The same pass for SMALL maps (P <= 1024 pixels per batch entry: the 8x8 / 4x4 levels, whose conv tiles span several
 * batch entries and therefore get no fused statistics from their producer): one block per batch entry first computes
 * the per-channel (sum, sum of squares) of x0 (compute_mask & 1) and / or x1 (& 2) and WRITES them to st0 / st1
 * ([B,C,2] float64, as dsep_channel_stats would), then proceeds as dsep_gn_act_split.  Saves one launch per tensor. */
int dsep_gn_stats_act_split(const float* x0, int C0, double* st0, const float* x1, int C1, double* st1, int B, int P,
                            int groups, const float* gamma, const float* beta, float eps, int act, void* a_hi,
                            void* a_lo, void* r_hi, void* r_lo, int compute_mask, dsep_stream_t stream);
/* 2x FIR resampling with taps [1,3,3,1] of an fp32 [B,H,W,C] tensor (mode 1: up, 2: down).
 * Outputs (each nullable pair): a = FIR(act(GN(x))) split, r = FIR(x) split, y = FIR(x) fp32.
 * st/gamma/beta NULL => no GroupNorm branch.
 * Replaces upsample_2d/downsample_2d (models/ncsnpp_utils/up_or_down_sampling.py:206-273)
 * -> upfirdn2d (op/upfirdn2d.py:145-156, upfirdn2d_kernel.cu:107-369). */
int dsep_fir_resample(const float* x, int B, int H, int W, int C, int mode, int groups,
                      const double* st, const float* gamma, const float* beta, float eps,
                      void* a_hi, void* a_lo, void* r_hi, void* r_lo, float* y, dsep_stream_t stream);
/* Same, with the a operand written for dsep_conv2d_fused8: a_hi = fp16 hi plane, a_8 = the e4m3 correction plane
 * ([B,Ho,Wo,2*C] bytes: per 8 channels [A_lo8 x 8 | A_hi8 x 8], prescaled by 2^(a8_exp+11) / 2^a8_exp). */
int dsep_fir_resample8(const float* x, int B, int H, int W, int C, int mode, int groups,
                       const double* st, const float* gamma, const float* beta, float eps,
                       void* a_hi, void* a_8, void* r_hi, void* r_lo, float* y, int a8_exp,
                       dsep_stream_t stream);
/* Same resampling with the activated branch in fp32: af = FIR(SiLU(GN(x))) [B,Ho,Wo,C] (and optionally y = FIR(x)),
 * for a consumer that builds its operand planes itself — the up / down ResBlock's Conv_0 then takes af as x0 of
 * dsep_conv2d_fused8 (no GroupNorm tables, no activation) and runs on the wide-tile kernel like every other conv. */
int dsep_fir_resample_f32(const float* x, int B, int H, int W, int C, int mode, int groups, const double* st,
                          const float* gamma, const float* beta, float eps, float* af, float* y,
                          dsep_stream_t stream);
/* Drop-in for the reference's own FFI signature on its own layout: in [planes, H, W] fp32,
 * kernel fixed to outer([1,3,3,1])/16*up^2; supports exactly the two calls the model makes
 * (up=2,down=1,pad=(2,1)) and (up=1,down=2,pad=(1,1)); anything else -> DSEP_ERR_UNSUPPORTED.
 * Replaces upfirdn2d_op (upfirdn2d_kernel.cu:209-369) as bound in upfirdn2d.cpp:12-22. */
int dsep_upfirdn2d(const float* in, int planes, int H, int W, int up_x, int up_y, int down_x,
                   int down_y, int pad_x0, int pad_x1, int pad_y0, int pad_y1, float* out,
                   dsep_stream_t stream);

/* out = h + bias + conv1x1(pyr): Combine(method="sum") (layerspp.py:52-57). pyr fp32 [B,P,Cp],
 * w fp32 [C,Cp]; out may alias h.  stats (nullable): fp64 [B,C,2], the per-channel (sum, sum of squares) of `out` are
 * ADDED to it (zero it first) — the statistics of the GroupNorm that reads the combined tensor (layerspp.py:291-296),
 * taken here instead of by a dsep_channel_stats pass over `out`. */
int dsep_combine(const float* pyr, int Cp, const float* w, const float* bias, const float* h,
                 float* out, int B, int P, int C, double* stats, dsep_stream_t stream);
/* y = a + b (fp32, n elements): pyramid accumulation (ncsnpp.py:440). */
/* im2col rows of a 3x3 / pad 1 convolution with few input channels: x [B,H,W,C] fp32 -> col [B,H,W,Cp] fp32,
 * col[pix][tap * C + c] = x[pix + tap offset][c] (zero outside the image and for columns >= 9 * C; tap = ky * 3 + kx).
 * The network's input conv (ncsnpp.py:347-349, C = 6) then runs as a 1x1 convolution with ONE 64-channel K-block
 * (dsep_conv2d_fused8 with x0 = col, no GroupNorm tables) instead of nine taps of a 6 -> 64 padded operand. */
int dsep_im2col3x3(const float* x, int B, int H, int W, int C, int Cp, float* col, dsep_stream_t stream);
/* Second half of a narrow 3x3 / pad 1 convolution run as "1x1 conv to 9 * CO channels, then gather": z [B,H,W,ZC] fp32
 * with z[pix][tap * CO + co] = sum_ci W[co,ci,ky,kx] * a[pix][ci] (tap = ky * 3 + kx; produced by dsep_conv2d_fused8 with
 * ksize 1 and the 54 tap-major weight rows) ->
 *   out[b,h,w,co] = bias[co] + residual[b,h,w,co] + sum_tap z[b, h + ky - 1, w + kx - 1][tap * CO + co]
 * over the neighbours inside the image; out / residual fp32 [B,H,W,CO], bias / residual may be null.  The output
 * pyramid's conv3x3(C -> 6) (ncsnpp.py:419-440, layers.py:141-156) on maps where a halo patch per 6 channels does not pay. */
int dsep_tap_gather3x3(const float* z, int B, int H, int W, int ZC, int CO, const float* bias,
                       const float* residual, float* out, dsep_stream_t stream);
int dsep_add(const float* a, const float* b, float* y, int64_t n, dsep_stream_t stream);

/* ---- attention ---------------------------------------------------------------------------
 * qkv fp32 [B,S,3C] (q | k | v per token); o = softmax(q k^T * scale) v written as split planes
 * [B,S,C].  Replaces the two einsums + softmax of AttnBlockpp.forward (layerspp.py:83-88).
 * C a multiple of 64 up to 256 (every NCSN++ width on this path): flash-style on tcgen05 — Q K^T and P V as
 * three fp16 products each with TMEM accumulators, 128 queries per CTA, keys in tiles of 32, the S x S matrix
 * never stored anywhere (csrc/attention_tc.cu); other widths: an fp32 CUDA-core kernel. */
int dsep_attention(const float* qkv, int B, int S, int C, float scale, void* o_hi, void* o_lo,
                   dsep_stream_t stream);

/* ---- time embedding -------------------------------------------------------------------------
 * temb_act = SiLU(Linear2(SiLU(Linear1([sin,cos](log(t) * Wf * 2 * pi))))) -> [B, 4nf]
 * (ncsnpp.py:324-343, layerspp.py:39-41; the SiLU every block applies first is folded in). */
int dsep_time_embedding(const float* t, const float* Wf, const float* w1, const float* b1,
                        const float* w2, const float* b2, int B, int nf, float* temb_act,
                        dsep_stream_t stream);
/* film[b, r] = sum_d temb_act[b,d] * Wd[r,d] + bd[r] for all ResBlocks' stacked Dense_0 rows
 * (layerspp.py:311-313). */
int dsep_film(const float* temb_act, const float* Wd, const float* bd, int B, int D, int R,
              float* film, dsep_stream_t stream);

/* ---- STFT-510 / iSTFT ---------------------------------------------------------------------
 * Framing exactly as torch.stft(n_fft=510, hop=128, center=True, pad_mode="constant") on the
 * signal right-padded by 382 zeros (score_models.py:107-112): frame f covers samples
 * [128 f - 255, 128 f + 255), zero outside [0,T), times periodic Hann(510).
 * frames: fp32 [B*C*Fr, 512] (columns 510, 511 zero).  x: [B, C, T].  window: fp32 [510]
 * (torch.hann_window(510), the `stft.window` buffer of a checkpoint). */
int dsep_stft_frames(const float* x, const float* window, int B, int C, int T, int Fr, float* frames,
                     dsep_stream_t stream);
/* C[M,N] = A[M,K] * Bm[K,N], fp32 FFMA, row-major, leading dimensions in elements.
 * Used with the DFT-510 basis in place of cuFFT (torch.stft/istft). */
int dsep_sgemm(const float* A, int lda, const float* Bm, int ldb, float* C, int ldc, int M, int N,
               int K, dsep_stream_t stream);
/* dft [B*Cw*Fr, 512] (re,im interleaved per bin) -> network input, channels-last:
 * |S|^e e^{j arg S} * factor (score_models.py:41-48), channel order [re_0..re_{Cw-1}, im_0..]
 * (:72-76), frames zero-padded to Wp (:83-91), then 2x-1 (ncsnpp.py:347-349).
 * chan0/Ctot let the xt channels and the (hoisted, step-invariant) mix channel be written by
 * separate calls: source channel c of this call lands in real slot chan0+c, imag slot
 * Ctot+chan0+c.  x_f32: [B,256,Wp,2*Ctot]; a_hi/a_lo: [B,256,Wp,Cpad] split (Cpad>=2*Ctot,
 * extra channels must be pre-zeroed by the caller). */
int dsep_spec_pack(const float* dft, int B, int Cw, int Fr, int Wp, int chan0, int Ctot, int Cpad,
                   float factor, float exponent, float* x_f32, void* a_hi, void* a_lo,
                   dsep_stream_t stream);
/* pyramid [B,256,Wp,Cp] fp32 -> /t[b] -> conv1x1 Cp->2*nsrc (ncsnpp.py:472-477) -> complex
 * [re_s, im_s] (score_models.py:78-81) -> /factor, |S|^{1/e} e^{j arg S} (:59-64) -> spec
 * [B*nsrc*Fr, 512] (re,im interleaved, frames beyond Fr dropped). */
int dsep_out_head(const float* pyr, int B, int Wp, int Cp, int nsrc, int Fr, const float* t,
                  const float* w, const float* bias, float factor, float exponent, float* spec,
                  dsep_stream_t stream);
/* frames_t [B*C*Fr, 512] (time-domain frames from the inverse basis) -> window, overlap-add,
 * divide by the squared-window envelope, drop 255 samples each side, crop / zero-pad to T
 * (torch.istft; score_models.py:122-123, :99-105).  out [B,C,T]. */
int dsep_istft_ola(const float* frames_t, const float* window, int B, int C, int Fr, int T, float* out,
                   dsep_stream_t stream);

/* ---- SDE arithmetic -------------------------------------------------------------------------
 * All on x [B,ndim,T], ndim = 2 sources or 3 (params.ndim; 0 reads as 2).  t is a device array [B].  sigma_mix
 * [B,T] is NULL for MixSDE and PriorMixSDE._std_sigma_mix(mix) for PriorMixSDE.  noise may be NULL: then N(0,1)
 * samples are drawn in-kernel (Philox4x32-10, Box-Muller) from (seed, offset).
 * prior:      x = mean_scale mix + L(T) z                  (sdes/sdes.py:334-346, 564-587); mix is [B,1,T]
 *             (mix_channels 1: broadcast over the sources, mean_scale 0.5) or, the true_mean branch, [B,ndim,T]
 *             (mean_scale 0.5 for MixSDE — its broadcast_to(0.5 y) — and 1 for PriorMixSDE, whose sigma_mix is
 *             then per source: sigma_channels = ndim, L = (s1 A + s2 Pn) diag(sigma))
 * corrector:  xm = x + 2 snr^2 L L score ; x' = xm + 2 snr L z   (sdes/correctors.py:109-128)
 * predictor:  xm = x + lambda dt (x - mean_c x) + c G^2 score ; x' = xm + G z, G = g(t) sqrt(dt)
 *             (sdes/predictors.py:39-66, sdes/sdes.py:93-107,163-171,275-284); probability_flow: c = 1/2
 *             and no noise (sdes.py:143-152,167-170), else c = 1 */
typedef struct {
    float d_lambda, sigma_min, sigma_max, T_end;
    int ndim;
} dsep_sde_params;
int dsep_sde_prior(const dsep_sde_params* p, const float* mix, int mix_channels, float mean_scale,
                   const float* sigma_mix, int sigma_channels, const float* noise, uint64_t seed, uint64_t offset,
                   int B, int T, float* x, dsep_stream_t stream);
int dsep_sde_corrector(const dsep_sde_params* p, const float* x, const float* score, const float* t,
                       const float* sigma_mix, const float* noise, uint64_t seed, uint64_t offset,
                       float snr, int B, int T, float* x_out, float* x_mean, dsep_stream_t stream);
int dsep_sde_predictor(const dsep_sde_params* p, const float* x, const float* score, const float* t,
                       const float* sigma_mix, const float* noise, uint64_t seed, uint64_t offset,
                       float dt, int probability_flow, int B, int T, float* x_out, float* x_mean,
                       dsep_stream_t stream);
/* ald, the original annealed Langevin corrector (sdes/correctors.py:58-91; MixSDE only):
 * std = sqrt(first-row sum of the covariance) = sqrt(ev1(t)); step = 2 (snr std)^2;
 * xm = x + step score ; x' = xm + sqrt(2 step) z. */
int dsep_sde_corrector_ald(const dsep_sde_params* p, const float* x, const float* score, const float* t,
                           const float* noise, uint64_t seed, uint64_t offset, float snr, int B, int T,
                           float* x_out, float* x_mean, dsep_stream_t stream);
/* langevin (sdes/correctors.py:35-55): step = 2 (snr mean_b||z_b|| / mean_b||score_b||)^2 (batch means
 * of per-entry L2 norms); xm = x + step score ; x' = xm + sqrt(2 step) z.  x, score, noise: [B, n];
 * norms: float scratch [2, B].  noise is an explicit tensor (dsep_randn). */
int dsep_sde_corrector_langevin(const float* x, const float* score, const float* noise, float snr, int B,
                                int n, float* norms, float* x_out, float* x_mean, dsep_stream_t stream);
/* PriorMixSDE._std_sigma_mix (sdes/sdes.py:477-489): 0.5 sqrt(clamp(avgpool_k(mix^2), 1e-4)). */
int dsep_sigma_mix(const float* mix, int B, int T, int avg_len, float* sigma, dsep_stream_t stream);
/* normalize_batch (pl_model.py:81-88): per-utterance (x - mean) / clamp(std_unbiased, 1e-5). */
int dsep_normalize(const float* mix, int B, int n, float* out, float* mean, float* std,
                   dsep_stream_t stream);
/* scale_output (separate.py:73-78): sep * <mix,sep> / sum(sep^2 + 1e-10), per (b, source). */
int dsep_scale_output(const float* mix, const float* sep, int B, int nsrc, int T, float* out,
                      dsep_stream_t stream);
/* z ~ N(0,1), n elements, Philox4x32-10 keyed by (seed, offset). */
int dsep_randn(float* z, int64_t n, uint64_t seed, uint64_t offset, dsep_stream_t stream);

/* Training-side forward pieces (forward only, no gradients).
 * dsep_sde_perturb: DiffSepModel.sample_prior with the default init_hack (pl_model.py:179-188, 243-247) on
 * sde.marginal_prob (sdes/sdes.py:322-324 / 560-562): x_t = (A + e^{-lambda t} Pn) x0 + L(t) z with
 * L = (sqrt(ev1) A + sqrt(ev2) Pn) [* sigma_mix (PriorMixSDE; [B,T] from dsep_sigma_mix, else NULL)].  x0, x_t: [B,ndim,T];
 * t: [B]; noise: an injected [B,ndim,T] tensor or NULL for the Philox stream (seed, offset); z_out (nullable)
 * receives the noise that was used (compute_score_loss needs it).
 * dsep_score_loss: the MSE of compute_score_loss after the network call (pl_model.py:418-424):
 * loss[b] = mean over (channel, time) of ((L(t) score) + z)^2, float64 [B] (its mean over b is MSELoss's default). */
int dsep_sde_perturb(const dsep_sde_params* p, const float* x0, const float* t, const float* sigma_mix,
                     const float* noise, uint64_t seed, uint64_t offset, int B, int T, float* x_t, float* z_out,
                     dsep_stream_t stream);
int dsep_score_loss(const dsep_sde_params* p, const float* score, const float* z, const float* t,
                    const float* sigma_mix, int B, int T, double* loss, dsep_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DSEP_H_ */
