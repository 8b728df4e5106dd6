// HBM-bound passes between the tensor-core convolutions: GroupNorm statistics, GroupNorm+SiLU
// with fp32 -> (hi, lo) fp16 operand splitting and channel concatenation, 2x FIR resampling
// (StyleGAN2 upfirdn2d with taps [1,3,3,1]), the 6-channel Combine 1x1, and small helpers.
// All tensors are channels-last so every warp touches contiguous 128-byte lines.
#include <stdarg.h>

#include "common.cuh"

namespace dsep {

// ------------------------------------------------------------------------- error plumbing
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return DSEP_ERR_CUDA;
    }
    return DSEP_OK;
}

struct alignas(8) f16x4 {
    __half v[4];
};

__device__ __forceinline__ void split4(const float4 x, f16x4& hi, f16x4& lo) {
    split_f16(x.x, hi.v[0], lo.v[0]);
    split_f16(x.y, hi.v[1], lo.v[1]);
    split_f16(x.z, hi.v[2], lo.v[2]);
    split_f16(x.w, hi.v[3], lo.v[3]);
}

// (hi fp16, e4m3 correction) form of a quad for the convolution's passes = 2 mode: per 8 channels the second
// plane holds the 16 bytes [A_lo8 x 8 | A_hi8 x 8] (conv_tc.cu patch_store); this quad's channels are
// 4 * (q & 1) .. + 3 of its group of 8.  a8_hi / a8_lo are the power-of-two prescales of the two parts.
__device__ __forceinline__ uint32_t e4m3x4_pack(float a, float b, float c, float d) {
    uint16_t p0, p1;
    asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(p0) : "f"(b), "f"(a));
    asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(p1) : "f"(d), "f"(c));
    return static_cast<uint32_t>(p0) | (static_cast<uint32_t>(p1) << 16);
}
__device__ __forceinline__ void split4_e4m3(const float4 x, f16x4& hi, uint32_t& lo8, uint32_t& hi8, float a8_hi,
                                            float a8_lo) {
    f16x4 l;
    split4(x, hi, l);
    const float h0 = __half2float(hi.v[0]), h1 = __half2float(hi.v[1]), h2 = __half2float(hi.v[2]),
                h3 = __half2float(hi.v[3]);
    const float c = 60000.0f;      // split_f16 clamps before rounding; take the residual of the clamped value
    const float x0 = fminf(fmaxf(x.x, -c), c), x1 = fminf(fmaxf(x.y, -c), c), x2 = fminf(fmaxf(x.z, -c), c),
                x3 = fminf(fmaxf(x.w, -c), c);
    lo8 = e4m3x4_pack((x0 - h0) * a8_lo, (x1 - h1) * a8_lo, (x2 - h2) * a8_lo, (x3 - h3) * a8_lo);
    hi8 = e4m3x4_pack(h0 * a8_hi, h1 * a8_hi, h2 * a8_hi, h3 * a8_hi);
}

// ---------------------------------------------------------------------------- split
__global__ void split_kernel(const float* __restrict__ x, int64_t n4, float prescale,
                             f16x4* __restrict__ hi, f16x4* __restrict__ lo) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4;
         i += (int64_t)gridDim.x * blockDim.x) {
        f16x4 h, l;
        float4 v = reinterpret_cast<const float4*>(x)[i];
        v.x *= prescale; v.y *= prescale; v.z *= prescale; v.w *= prescale;
        split4(v, h, l);
        hi[i] = h;
        lo[i] = l;
    }
}

// ---------------------------------------------------------------------------- channel statistics
// stats[b, c, 0..1] += (sum, sum of squares) of x[b, :, c].  grid (chunks, B); thread t owns channel
// quad (t % Q) of pixels (t / Q) + k*ppi of its chunk; fp64 accumulation end to end.
__global__ void __launch_bounds__(256)
channel_stats_kernel(const float* __restrict__ x, int C, int P, int pix_per_block, double* __restrict__ stats) {
    extern __shared__ double s_acc[];   // [C][2]
    const int Q = C >> 2;
    const int b = blockIdx.y;
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) s_acc[i] = 0.0;
    __syncthreads();
    const int ppi = blockDim.x / Q;
    const int q = threadIdx.x % Q, j = threadIdx.x / Q;
    if (j < ppi) {
        const int p_begin = blockIdx.x * pix_per_block;
        const int p_end = min(P, p_begin + pix_per_block);
        double s[4] = {0.0, 0.0, 0.0, 0.0}, ss[4] = {0.0, 0.0, 0.0, 0.0};
        const float* src = x + static_cast<size_t>(b) * P * C + q * 4;
        for (int p = p_begin + j; p < p_end; p += ppi) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(src + static_cast<size_t>(p) * C));
            s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
            ss[0] += (double)v.x * v.x; ss[1] += (double)v.y * v.y;
            ss[2] += (double)v.z * v.z; ss[3] += (double)v.w * v.w;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            atomicAdd(&s_acc[(q * 4 + u) * 2 + 0], s[u]);
            atomicAdd(&s_acc[(q * 4 + u) * 2 + 1], ss[u]);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x)
        atomicAdd(&stats[static_cast<size_t>(b) * 2 * C + i], s_acc[i]);
}

// Per-channel GroupNorm scale / shift of batch entry b from per-channel sums of a (possibly
// concatenated) input: y = x * sc[c] + sh[c].  Block-wide; s_sc / s_sh have Ct entries.
__device__ __forceinline__ void gn_tables(const double* __restrict__ st0, int C0, const double* __restrict__ st1,
                                          int C1, int b, int groups, double count_per_channel,
                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                          float eps, float* s_sc, float* s_sh, float* s_mean, float* s_rstd) {
    const int Ct = C0 + C1, cpg = Ct / groups;
    if (st0 == nullptr) {
        for (int c = threadIdx.x; c < Ct; c += blockDim.x) { s_sc[c] = 1.0f; s_sh[c] = 0.0f; }
        __syncthreads();
        return;
    }
    for (int g = threadIdx.x; g < groups; g += blockDim.x) {
        double s = 0.0, ss = 0.0;
        for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
            const double* st = c < C0 ? st0 + (static_cast<size_t>(b) * C0 + c) * 2
                                      : st1 + (static_cast<size_t>(b) * C1 + (c - C0)) * 2;
            s += st[0];
            ss += st[1];
        }
        const double n = count_per_channel * cpg;
        const double m = s / n;
        double var = ss / n - m * m;
        if (var < 0.0) var = 0.0;
        s_mean[g] = static_cast<float>(m);
        s_rstd[g] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    }
    __syncthreads();
    for (int c = threadIdx.x; c < Ct; c += blockDim.x) {
        const int g = c / cpg;
        const float sc = gamma[c] * s_rstd[g];
        s_sc[c] = sc;
        s_sh[c] = beta[c] - s_mean[g] * sc;
    }
    __syncthreads();
}

// one block per batch entry: writes the tables to global memory for dsep_conv2d_fused's prologue
__global__ void __launch_bounds__(256)
gn_tables_kernel(const double* __restrict__ st0, int C0, const double* __restrict__ st1, int C1, int P, int groups,
                 const float* __restrict__ gamma, const float* __restrict__ beta, float eps, float* __restrict__ sc,
                 float* __restrict__ sh) {
    extern __shared__ float s_tab[];   // sc[Ct] | sh[Ct] | mean[64] | rstd[64]
    const int Ct = C0 + C1, b = blockIdx.x;
    gn_tables(st0, C0, st1, C1, b, groups, (double)P, gamma, beta, eps, s_tab, s_tab + Ct, s_tab + 2 * Ct,
              s_tab + 2 * Ct + 64);
    for (int c = threadIdx.x; c < Ct; c += blockDim.x) {
        sc[static_cast<size_t>(b) * Ct + c] = s_tab[c];
        sh[static_cast<size_t>(b) * Ct + c] = s_tab[Ct + c];
    }
}

__device__ __forceinline__ float silu_fast(float x) { return __fdividef(x, 1.0f + __expf(-x)); }

__device__ __forceinline__ float4 affine_act(const float4 v, const float4 sc, const float4 sh, bool act) {
    float4 y;
    y.x = fmaf(v.x, sc.x, sh.x); y.y = fmaf(v.y, sc.y, sh.y);
    y.z = fmaf(v.z, sc.z, sh.z); y.w = fmaf(v.w, sc.w, sh.w);
    if (act) { y.x = silu_fast(y.x); y.y = silu_fast(y.y); y.z = silu_fast(y.z); y.w = silu_fast(y.w); }
    return y;
}

// ------------------------------------------------------- GN + act + split (+ raw split)
// grid (chunks, B); each thread converts one channel-quad of one pixel per step, two steps in flight.
// One block's per-channel (sum, sum of squares) of ITS batch entry's [P, Cs] slab -> st[b, c, 2] (float64): the small
// maps (8x8, 4x4: 16 - 64 pixels) whose conv tiles span several batch entries get no fused statistics from their
// producer, and a separate channel_stats launch per tensor was 32 launches of pure latency per evaluation.
__device__ __forceinline__ void slab_stats(const float* __restrict__ xs, int Cs, int P, int b, double* __restrict__ st,
                                           float* s_red /* [256 * 8] */) {
    const int Qs = Cs >> 2, rows = 256 / Qs;                // host guarantees 256 % Qs == 0
    const int q = threadIdx.x % Qs, r = threadIdx.x / Qs;
    float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const float4* base = reinterpret_cast<const float4*>(xs + static_cast<size_t>(b) * P * Cs) + q;
    auto add = [&](const float4 v) {
        a[0] += v.x; a[1] += v.y; a[2] += v.z; a[3] += v.w;
        a[4] = fmaf(v.x, v.x, a[4]); a[5] = fmaf(v.y, v.y, a[5]); a[6] = fmaf(v.z, v.z, a[6]); a[7] = fmaf(v.w, v.w, a[7]);
    };
    int p = r;
    for (; p + 3 * rows < P; p += 4 * rows) {        // four loads in flight (the slab comes from L2)
        const float4 v0 = __ldg(base + static_cast<size_t>(p) * Qs), v1 = __ldg(base + static_cast<size_t>(p + rows) * Qs);
        const float4 v2 = __ldg(base + static_cast<size_t>(p + 2 * rows) * Qs);
        const float4 v3 = __ldg(base + static_cast<size_t>(p + 3 * rows) * Qs);
        add(v0); add(v1); add(v2); add(v3);
    }
    for (; p < P; p += rows) add(__ldg(base + static_cast<size_t>(p) * Qs));
#pragma unroll
    for (int i = 0; i < 8; ++i) s_red[threadIdx.x * 8 + i] = a[i];
    __syncthreads();
    if (threadIdx.x < Qs) {
        double t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int rr = 0; rr < rows; ++rr)
#pragma unroll
            for (int i = 0; i < 8; ++i) t[i] += static_cast<double>(s_red[(rr * Qs + threadIdx.x) * 8 + i]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double* o = st + (static_cast<size_t>(b) * Cs + threadIdx.x * 4 + j) * 2;
            o[0] = t[j];
            o[1] = t[4 + j];
        }
    }
    __syncthreads();      // the block reads these sums back (gn_tables): block-scope ordering of its own global writes
}

__global__ void __launch_bounds__(256)
gn_act_split_kernel(const float* __restrict__ x0, int C0, double* __restrict__ st0,
                    const float* __restrict__ x1, int C1, double* __restrict__ st1, int P, int groups,
                    const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int act,
                    f16x4* __restrict__ a_hi, f16x4* __restrict__ a_lo, f16x4* __restrict__ r_hi,
                    f16x4* __restrict__ r_lo, int pix_per_block, int compute_mask) {
    extern __shared__ float s_tab[];   // sc[Ct] | sh[Ct] | mean[64] | rstd[64]  (compute_mask: + [256 * 8] scratch)
    const int Ct = C0 + C1, Q = Ct >> 2;
    float* s_sc = s_tab;
    float* s_sh = s_tab + Ct;
    const int b = blockIdx.y;
    if (compute_mask) {       // one block per batch entry (grid.x == 1)
        float* s_red = s_tab + 2 * Ct + 128;
        if (compute_mask & 1) slab_stats(x0, C0, P, b, st0, s_red);
        if (compute_mask & 2) slab_stats(x1, C1, P, b, st1, s_red);
    }
    gn_tables(st0, C0, st1, C1, b, groups, (double)P, gamma, beta, eps, s_sc, s_sh, s_tab + 2 * Ct,
              s_tab + 2 * Ct + 64);
    const int64_t e_begin = (int64_t)blockIdx.x * pix_per_block * Q;
    const int64_t e_end = min((int64_t)P * Q, e_begin + (int64_t)pix_per_block * Q);
    const bool on = act != 0;
    auto load = [&](int64_t e, int& q) -> float4 {
        const int p = static_cast<int>(e / Q);
        q = static_cast<int>(e - (int64_t)p * Q);
        const int c = q * 4;
        return c < C0 ? __ldg(reinterpret_cast<const float4*>(x0 + (static_cast<size_t>(b) * P + p) * C0 + c))
                      : __ldg(reinterpret_cast<const float4*>(x1 + (static_cast<size_t>(b) * P + p) * C1 + (c - C0)));
    };
    auto emit = [&](int64_t e, int q, const float4 v) {
        const size_t o = static_cast<size_t>(b) * P * Q + e;
        f16x4 h, l;
        if (r_hi != nullptr) {
            split4(v, h, l);
            r_hi[o] = h;
            r_lo[o] = l;
        }
        if (a_hi != nullptr) {
            const float4 sc = *reinterpret_cast<const float4*>(s_sc + q * 4);
            const float4 sh = *reinterpret_cast<const float4*>(s_sh + q * 4);
            split4(affine_act(v, sc, sh, on), h, l);
            a_hi[o] = h;
            a_lo[o] = l;
        }
    };
    int64_t e = e_begin + threadIdx.x;
    for (; e + blockDim.x < e_end; e += 2 * blockDim.x) {
        int q0, q1;
        const float4 v0 = load(e, q0);
        const float4 v1 = load(e + blockDim.x, q1);
        emit(e, q0, v0);
        emit(e + blockDim.x, q1, v1);
    }
    if (e < e_end) {
        int q0;
        const float4 v0 = load(e, q0);
        emit(e, q0, v0);
    }
}

// --------------------------------------------------------------------------- FIR resample
// taps [1,3,3,1]; down: y[i] = (x[2i-1] + 3x[2i] + 3x[2i+1] + x[2i+2]) / 8 per axis;
// up (gain 2 per axis): y[2i] = (x[i-1] + 3x[i]) / 4, y[2i+1] = (3x[i] + x[i+1]) / 4; zeros outside.
//
// Tiled: a block stages the input patch of one (batch entry, spatial tile, 32-channel chunk) in shared
// memory ONCE — with GroupNorm+SiLU already applied, so the transcendental is evaluated once per
// input element instead of once per tap — and then forms every output of the tile from it.
//   MODE 2 (down): 4 x 8 output pixels  <- 10 x 18 input pixels
//   MODE 1 (up):  16 x 16 output pixels <- 10 x 10 input pixels (8 x 8 + halo)
template <int MODE>
struct FirTile {
    static constexpr int OH = MODE == 1 ? 16 : 4, OW = MODE == 1 ? 16 : 8;
    static constexpr int IH = MODE == 1 ? 10 : 10, IW = MODE == 1 ? 10 : 18;
    static constexpr int kSmemBytes = 2 * IH * IW * 32 * 4;
};

template <int MODE>
#ifndef DSEP_FIR_BLOCKS
#define DSEP_FIR_BLOCKS 4
#endif
__global__ void __launch_bounds__(256, DSEP_FIR_BLOCKS)   // 4 blocks of 46 KB per SM: the kernel is latency-bound (load -> sync -> stage -> sync)
fir_tile_kernel(const float* __restrict__ x, int H, int W, int C, int groups, const double* __restrict__ st,
                const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                f16x4* __restrict__ a_hi, f16x4* __restrict__ a_lo, f16x4* __restrict__ r_hi,
                f16x4* __restrict__ r_lo, float4* __restrict__ y, float4* __restrict__ af, int tiles_w, int tiles_h,
                float a8_hi, float a8_lo, int kt) {
    using FT = FirTile<MODE>;
    extern __shared__ __align__(16) float s_fir[];
    float4* s_act = reinterpret_cast<float4*>(s_fir);                  // [IH*IW][8 quads]
    float4* s_raw = s_act + FT::IH * FT::IW * 8;
    __shared__ float s_sc[32], s_sh[32];
    const int Ho = MODE == 1 ? H * 2 : H / 2, Wo = MODE == 1 ? W * 2 : W / 2;
    const int chunk = blockIdx.y, b = blockIdx.z;
    // a block walks kt consecutive tiles of one tile row: the next tile's patch loads are in flight while the current
    // tile is computed out of shared memory (one tile per block spent most of its life in load -> sync -> stage ->
    // sync), and the GroupNorm scale / shift of the chunk is computed once per block
    const int groups_w = (tiles_w + kt - 1) / kt;
    const int th_i = blockIdx.x / groups_w;
    const int tw_begin = (blockIdx.x % groups_w) * kt;
    const int tw_end = min(tiles_w, tw_begin + kt);
    const int oh0 = th_i * FT::OH;
    const int ih0 = MODE == 1 ? oh0 / 2 - 1 : oh0 * 2 - 1;
    const int c0 = chunk * 32;
    const bool want_act = a_hi != nullptr || af != nullptr;     // the GroupNorm + SiLU branch (planes and / or fp32)
    const bool xf = st != nullptr && want_act;
    const bool want_raw = r_hi != nullptr || y != nullptr;
    // the patch loads go out first and stay in flight while 32 threads turn the GroupNorm sums into this chunk's
    // scale / shift (both are one memory round trip; back to back they were ~2 of the ~8 us a block lives)
    constexpr int NI = FT::IH * FT::IW * 8;
    constexpr int PER = (NI + 255) / 256;
    float4 raw[PER];
    auto load_tile = [&](int tw_i) {
        const int ow0 = tw_i * FT::OW;
        const int iw0 = MODE == 1 ? ow0 / 2 - 1 : ow0 * 2 - 1;
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const int i = threadIdx.x + k * 256;
            raw[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < NI) {
                const int q = i & 7, pix = i >> 3;
                const int ih = ih0 + pix / FT::IW, iw = iw0 + pix % FT::IW;
                if (ih >= 0 && ih < H && iw >= 0 && iw < W)
                    raw[k] = __ldg(reinterpret_cast<const float4*>(x + ((static_cast<size_t>(b) * H + ih) * W + iw) * C + c0 + q * 4));
            }
        }
    };
    load_tile(tw_begin);
    if (threadIdx.x < 32) {
        float sc = 1.0f, sh = 0.0f;
        if (xf) {
            const int cpg = C / groups, c = c0 + threadIdx.x, g = c / cpg;
            double s = 0.0, ss = 0.0;
            for (int cc = g * cpg; cc < (g + 1) * cpg; ++cc) {
                s += st[(static_cast<size_t>(b) * C + cc) * 2 + 0];
                ss += st[(static_cast<size_t>(b) * C + cc) * 2 + 1];
            }
            const double n = (double)H * W * cpg, m = s / n;
            double var = ss / n - m * m;
            if (var < 0.0) var = 0.0;
            sc = gamma[c] * static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
            sh = beta[c] - static_cast<float>(m) * sc;
        }
        s_sc[threadIdx.x] = sc;
        s_sh[threadIdx.x] = sh;
    }
    __syncthreads();
    const int Q = C >> 2;
    for (int tw_i = tw_begin; tw_i < tw_end; ++tw_i) {
    const int ow0 = tw_i * FT::OW;
    const int iw0 = MODE == 1 ? ow0 / 2 - 1 : ow0 * 2 - 1;
    // stage the patch: out-of-image pixels are zero AFTER the activation (the FIR pads its input)
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        const int i = threadIdx.x + k * 256;
        if (i < NI) {
            const int q = i & 7, pix = i >> 3;
            const int ih = ih0 + pix / FT::IW, iw = iw0 + pix % FT::IW;
            float4 av = make_float4(0.f, 0.f, 0.f, 0.f);
            if (want_act && ih >= 0 && ih < H && iw >= 0 && iw < W)
                av = xf ? affine_act(raw[k], *reinterpret_cast<const float4*>(s_sc + q * 4),
                                     *reinterpret_cast<const float4*>(s_sh + q * 4), true)
                        : raw[k];
            s_act[i] = av;
            if (want_raw) s_raw[i] = raw[k];
        }
    }
    __syncthreads();
    if (tw_i + 1 < tw_end) load_tile(tw_i + 1);        // in flight while this tile is computed
    for (int i = threadIdx.x; i < FT::OH * FT::OW * 8; i += blockDim.x) {
        const int q = i & 7, opix = i >> 3;
        const int oy = opix / FT::OW, ox = opix % FT::OW;
        const int oh = oh0 + oy, ow = ow0 + ox;
        if (oh >= Ho || ow >= Wo) continue;
        float4 acc_a = make_float4(0.f, 0.f, 0.f, 0.f), acc_r = acc_a;
        if (MODE == 1) {
            // patch row of input pixel (oh/2) is oy/2 + 1; even outputs use (i-1, i), odd (i, i+1)
            const int py = (oy >> 1) + ((oy & 1) ? 1 : 0), px = (ox >> 1) + ((ox & 1) ? 1 : 0);
            const float wy0 = (oy & 1) ? 0.75f : 0.25f, wx0 = (ox & 1) ? 0.75f : 0.25f;
#pragma unroll
            for (int a = 0; a < 2; ++a) {
#pragma unroll
                for (int d = 0; d < 2; ++d) {
                    const float wgt = (a == 0 ? wy0 : 1.0f - wy0) * (d == 0 ? wx0 : 1.0f - wx0);
                    const int si = ((py + a) * FT::IW + (px + d)) * 8 + q;
                    if (want_act) {
                        const float4 t = s_act[si];
                        acc_a.x = fmaf(wgt, t.x, acc_a.x); acc_a.y = fmaf(wgt, t.y, acc_a.y);
                        acc_a.z = fmaf(wgt, t.z, acc_a.z); acc_a.w = fmaf(wgt, t.w, acc_a.w);
                    }
                    if (want_raw) {
                        const float4 t = s_raw[si];
                        acc_r.x = fmaf(wgt, t.x, acc_r.x); acc_r.y = fmaf(wgt, t.y, acc_r.y);
                        acc_r.z = fmaf(wgt, t.z, acc_r.z); acc_r.w = fmaf(wgt, t.w, acc_r.w);
                    }
                }
            }
        } else {
            const float wt[4] = {0.125f, 0.375f, 0.375f, 0.125f};
#pragma unroll
            for (int a = 0; a < 4; ++a) {
#pragma unroll
                for (int d = 0; d < 4; ++d) {
                    const float wgt = wt[a] * wt[d];
                    const int si = ((2 * oy + a) * FT::IW + (2 * ox + d)) * 8 + q;
                    if (want_act) {
                        const float4 t = s_act[si];
                        acc_a.x = fmaf(wgt, t.x, acc_a.x); acc_a.y = fmaf(wgt, t.y, acc_a.y);
                        acc_a.z = fmaf(wgt, t.z, acc_a.z); acc_a.w = fmaf(wgt, t.w, acc_a.w);
                    }
                    if (want_raw) {
                        const float4 t = s_raw[si];
                        acc_r.x = fmaf(wgt, t.x, acc_r.x); acc_r.y = fmaf(wgt, t.y, acc_r.y);
                        acc_r.z = fmaf(wgt, t.z, acc_r.z); acc_r.w = fmaf(wgt, t.w, acc_r.w);
                    }
                }
            }
        }
        const size_t o = ((static_cast<size_t>(b) * Ho + oh) * Wo + ow) * Q + chunk * 8 + q;
        f16x4 h, l;
        if (a_hi != nullptr) {
            if (a8_hi != 0.f) {      // second plane = e4m3 corrections: [A_lo8 x 8 | A_hi8 x 8] per 8 channels
                uint32_t lo8, hi8;
                split4_e4m3(acc_a, h, lo8, hi8, a8_hi, a8_lo);
                a_hi[o] = h;
                uint32_t* g8 = reinterpret_cast<uint32_t*>(a_lo + (o & ~static_cast<size_t>(1)));   // the group's 16 bytes
                g8[o & 1] = lo8;
                g8[2 + (o & 1)] = hi8;
            } else {
                split4(acc_a, h, l); a_hi[o] = h; a_lo[o] = l;
            }
        }
        if (r_hi != nullptr) { split4(acc_r, h, l); r_hi[o] = h; r_lo[o] = l; }
        if (y != nullptr) y[o] = acc_r;
        if (af != nullptr) af[o] = acc_a;
    }
    __syncthreads();       // every thread is done with this tile's shared-memory patch before the next one is staged
    }
}

// scalar variant for channel counts that are not a multiple of 4 (the 6-channel pyramids);
// also the kernel behind dsep_upfirdn2d (C = 1 planes).
template <int MODE>
__global__ void __launch_bounds__(256)
fir_scalar_kernel(const float* __restrict__ x, int64_t total_out, int H, int W, int C,
                  float* __restrict__ y) {
    const int Ho = MODE == 1 ? H * 2 : H / 2, Wo = MODE == 1 ? W * 2 : W / 2;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total_out;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int c = static_cast<int>(e % C);
        int64_t r = e / C;
        const int oj = static_cast<int>(r % Wo); r /= Wo;
        const int oi = static_cast<int>(r % Ho);
        const int64_t b = r / Ho;
        constexpr int NTAP = MODE == 1 ? 2 : 4;
        int i0, j0;
        float wi[NTAP], wj[NTAP];
        if (MODE == 1) {
            const int ii = oi >> 1, jj = oj >> 1;
            if (oi & 1) { i0 = ii; wi[0] = 0.75f; wi[1] = 0.25f; } else { i0 = ii - 1; wi[0] = 0.25f; wi[1] = 0.75f; }
            if (oj & 1) { j0 = jj; wj[0] = 0.75f; wj[1] = 0.25f; } else { j0 = jj - 1; wj[0] = 0.25f; wj[1] = 0.75f; }
        } else {
            i0 = 2 * oi - 1; j0 = 2 * oj - 1;
            wi[0] = 0.125f; wi[1] = 0.375f; wi[2] = 0.375f; wi[3] = 0.125f;
            wj[0] = 0.125f; wj[1] = 0.375f; wj[2] = 0.375f; wj[3] = 0.125f;
        }
        float acc = 0.f;
#pragma unroll
        for (int a = 0; a < NTAP; ++a) {
            const int i = i0 + a;
            if (i < 0 || i >= H) continue;
#pragma unroll
            for (int d = 0; d < NTAP; ++d) {
                const int j = j0 + d;
                if (j < 0 || j >= W) continue;
                acc += wi[a] * wj[d] * x[((b * H + i) * W + j) * C + c];
            }
        }
        y[e] = acc;
    }
}

// ------------------------------------------------------------------------------ Combine
// out[b,p,c] = h[b,p,c] + bias[c] + sum_k w[c,k] * pyr[b,p,k]
// weights are staged transposed, s_w[k][c]: a warp reads 32 consecutive channel quads of one k with one
// conflict-free LDS.128 per lane (the [c][k] order put the lanes' words 24 floats apart: 8-way bank conflicts,
// 1.5 TB/s; profiles/membound_r2a.md)
// grid (chunks, B): a block owns a run of pixels of ONE batch entry, so that the per-channel (sum, sum of squares) of
// its outputs — the statistics of the GroupNorm that consumes the combined tensor — leave the block once (fp64,
// accumulated like channel_stats_kernel does) instead of costing a channel_stats launch and a second read of `out`.
__global__ void __launch_bounds__(256)
combine_kernel(const float* __restrict__ pyr, int Cp, const float* __restrict__ w,
               const float* __restrict__ bias, const float* __restrict__ h, float* __restrict__ out,
               int P, int C, int pix_per_block, double* __restrict__ stats) {
    extern __shared__ __align__(16) float s_w[];   // [Cp][C] + [C] (+ [C][2] doubles when stats are taken)
    float* s_b = s_w + C * Cp;
    double* s_acc = reinterpret_cast<double*>(s_b + C);
    for (int i = threadIdx.x; i < C * Cp; i += blockDim.x) s_w[(i % Cp) * C + i / Cp] = w[i];
    for (int i = threadIdx.x; i < C; i += blockDim.x) s_b[i] = bias ? bias[i] : 0.f;
    if (stats != nullptr)
        for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) s_acc[i] = 0.0;
    __syncthreads();
    const int Q = C >> 2, b = blockIdx.y;
    const int ppi = blockDim.x / Q;                 // pixels per iteration (threads beyond ppi * Q idle)
    const int q = threadIdx.x % Q, j = threadIdx.x / Q;
    if (j < ppi) {
        const int c = q * 4;
        const int p_begin = blockIdx.x * pix_per_block;
        const int p_end = min(P, p_begin + pix_per_block);
        double s[4] = {0.0, 0.0, 0.0, 0.0}, ss[4] = {0.0, 0.0, 0.0, 0.0};
        for (int p = p_begin + j; p < p_end; p += ppi) {
            const int64_t pix = static_cast<int64_t>(b) * P + p;
            const int64_t e = pix * Q + q;
            float4 v = __ldg(reinterpret_cast<const float4*>(h) + e);
            float4 a = *reinterpret_cast<const float4*>(s_b + c);
            for (int k = 0; k < Cp; ++k) {
                const float pk = __ldg(pyr + pix * Cp + k);
                const float4 wk = *reinterpret_cast<const float4*>(s_w + k * C + c);
                a.x = fmaf(wk.x, pk, a.x); a.y = fmaf(wk.y, pk, a.y); a.z = fmaf(wk.z, pk, a.z); a.w = fmaf(wk.w, pk, a.w);
            }
            v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
            reinterpret_cast<float4*>(out)[e] = v;
            s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
            ss[0] += (double)v.x * v.x; ss[1] += (double)v.y * v.y;
            ss[2] += (double)v.z * v.z; ss[3] += (double)v.w * v.w;
        }
        if (stats != nullptr) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                atomicAdd(&s_acc[(c + u) * 2 + 0], s[u]);
                atomicAdd(&s_acc[(c + u) * 2 + 1], ss[u]);
            }
        }
    }
    if (stats != nullptr) {
        __syncthreads();
        for (int i = threadIdx.x; i < 2 * C; i += blockDim.x)
            atomicAdd(&stats[static_cast<size_t>(b) * 2 * C + i], s_acc[i]);
    }
}

__global__ void add_kernel(const float* __restrict__ a, const float* __restrict__ b,
                           float* __restrict__ y, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        y[i] = a[i] + b[i];
}

static int grid_for(int64_t work_items, int threads = 256, int max_blocks = 148 * 16) {
    int64_t g = (work_items + threads - 1) / threads;
    if (g > max_blocks) g = max_blocks;
    if (g < 1) g = 1;
    return static_cast<int>(g);
}

// pixels handled per block so that the grid is ~8 blocks per SM per batch entry at most
static int pix_per_block_for(int P, int B) {
    int target_blocks = (148 * 8 + B - 1) / B;
    if (target_blocks < 1) target_blocks = 1;
    int ppb = (P + target_blocks - 1) / target_blocks;
    if (ppb < 8) ppb = 8;
    return ppb;
}

}  // namespace dsep

using namespace dsep;

extern "C" const char* dsep_last_error(void) { return dsep::g_err; }
extern "C" int dsep_abi_version(void) { return DSEP_ABI_VERSION; }
#ifndef DSEP_SOURCE_HASH
#define DSEP_SOURCE_HASH "unknown"
#endif
// "dsep-source-hash=<hash>" as one literal: build.py finds the tag by scanning the file, without dlopen()ing a
// possibly stale library into the process that is about to replace it
static const char kSourceHashTag[] = "dsep-source-hash=" DSEP_SOURCE_HASH;
extern "C" const char* dsep_source_hash(void) { return kSourceHashTag + 17; }
extern "C" int dsep_device_ok(void) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
    return major == 10 ? 1 : 0;
}

extern "C" int dsep_split_f16(const float* x, int64_t n, float prescale, void* hi, void* lo,
                              dsep_stream_t stream) {
    DSEP_REQUIRE(x && hi && lo, "split_f16: null pointer");
    DSEP_REQUIRE(n >= 0 && n % 4 == 0, "split_f16: n must be a multiple of 4");
    if (n == 0) return DSEP_OK;
    split_kernel<<<grid_for(n / 4), 256, 0, (cudaStream_t)stream>>>(x, n / 4, prescale, (f16x4*)hi, (f16x4*)lo);
    return check_launch("split_kernel");
}

static int check_gn_shape(const char* who, int C0, int C1, int groups) {
    const int Ct = C0 + C1;
    DSEP_REQUIRE(C0 > 0 && C1 >= 0 && C0 % 4 == 0 && C1 % 4 == 0, "%s: channel counts must be multiples of 4", who);
    DSEP_REQUIRE(groups > 0 && groups <= 64 && Ct % groups == 0, "%s: unsupported groups=%d for %d channels", who,
                 groups, Ct);
    DSEP_REQUIRE(Ct <= 1024, "%s: at most 1024 channels", who);
    return DSEP_OK;
}

extern "C" int dsep_zero(void* ptr, int64_t bytes, dsep_stream_t stream) {
    DSEP_REQUIRE(ptr && bytes >= 0, "zero: bad arguments");
    if (bytes == 0) return DSEP_OK;
    cudaError_t e = cudaMemsetAsync(ptr, 0, static_cast<size_t>(bytes), (cudaStream_t)stream);
    if (e != cudaSuccess) {
        set_error("zero: %s", cudaGetErrorString(e));
        return DSEP_ERR_CUDA;
    }
    return DSEP_OK;
}

extern "C" int dsep_channel_stats(const float* x, int C, int B, int P, double* stats, dsep_stream_t stream) {
    DSEP_REQUIRE(x && stats, "channel_stats: null pointer");
    DSEP_REQUIRE(B > 0 && P > 0 && B <= 65535, "channel_stats: empty tensor");
    DSEP_REQUIRE(C > 0 && C % 4 == 0 && C <= 1024, "channel_stats: C must be a multiple of 4, at most 1024");
    cudaStream_t s = (cudaStream_t)stream;
    cudaMemsetAsync(stats, 0, sizeof(double) * 2 * B * C, s);
    const int ppb = pix_per_block_for(P, B);
    dim3 grid(ceil_div(P, ppb), B);
    channel_stats_kernel<<<grid, 256, sizeof(double) * 2 * C, s>>>(x, C, P, ppb, stats);
    return check_launch("channel_stats_kernel");
}

extern "C" int dsep_gn_tables(const double* st0, int C0, const double* st1, int C1, int B, int P, int groups,
                              const float* gamma, const float* beta, float eps, float* sc, float* sh,
                              dsep_stream_t stream) {
    DSEP_REQUIRE(st0 && gamma && beta && sc && sh && (C1 == 0 || st1), "gn_tables: null pointer");
    DSEP_REQUIRE(B > 0 && P > 0 && B <= 65535, "gn_tables: empty tensor");
    int rc = check_gn_shape("gn_tables", C0, C1, groups);
    if (rc) return rc;
    const size_t smem = sizeof(float) * (2 * (C0 + C1) + 128);
    gn_tables_kernel<<<B, 256, smem, (cudaStream_t)stream>>>(st0, C0, st1, C1, P, groups, gamma, beta, eps, sc, sh);
    return check_launch("gn_tables_kernel");
}

static int gn_act_split_impl(const float* x0, int C0, double* st0, const float* x1, int C1, double* st1, int B, int P,
                             int groups, const float* gamma, const float* beta, float eps, int act, void* a_hi,
                             void* a_lo, void* r_hi, void* r_lo, int compute_mask, dsep_stream_t stream) {
    DSEP_REQUIRE(x0 && (C1 == 0 || x1), "gn_act_split: null input");
    DSEP_REQUIRE((a_hi && a_lo) || (r_hi && r_lo), "gn_act_split: no output requested");
    DSEP_REQUIRE(st0 == nullptr || (gamma && beta && (C1 == 0 || st1)),
                 "gn_act_split: statistics need gamma, beta and (for a concatenated input) both halves");
    DSEP_REQUIRE(act == 0 || act == 1, "gn_act_split: act must be 0 or 1");
    DSEP_REQUIRE(B > 0 && P > 0 && B <= 65535, "gn_act_split: empty tensor");
    if (st0 == nullptr) groups = 1;
    int rc = check_gn_shape("gn_act_split", C0, C1, groups);
    if (rc) return rc;
    int ppb = pix_per_block_for(P, B);
    size_t smem = sizeof(float) * (2 * (C0 + C1) + 128);
    if (compute_mask) {
        DSEP_REQUIRE((compute_mask & ~3) == 0 && st0 && (!(compute_mask & 2) || (C1 > 0 && st1)),
                     "gn_stats_act_split: bad compute mask %d", compute_mask);
        DSEP_REQUIRE(P <= 1024 && (!(compute_mask & 1) || (C0 <= 1024 && 256 % (C0 / 4) == 0)) &&
                         (!(compute_mask & 2) || (C1 <= 1024 && 256 % (C1 / 4) == 0)),
                     "gn_stats_act_split: in-kernel statistics are for small maps (P <= 1024) and channel counts "
                     "whose quads divide 256 (got P=%d, C0=%d, C1=%d)", P, C0, C1);
        ppb = P;                                  // one block per batch entry
        smem += sizeof(float) * 256 * 8;
    }
    dim3 grid(ceil_div(P, ppb), B);
    gn_act_split_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(
        x0, C0, st0, x1, C1, st1, P, groups, gamma, beta, eps, act, (f16x4*)a_hi, (f16x4*)a_lo, (f16x4*)r_hi,
        (f16x4*)r_lo, ppb, compute_mask);
    return check_launch("gn_act_split_kernel");
}

extern "C" int dsep_gn_act_split(const float* x0, int C0, const double* st0, const float* x1, int C1,
                                 const double* st1, int B, int P, int groups, const float* gamma,
                                 const float* beta, float eps, int act, void* a_hi, void* a_lo, void* r_hi,
                                 void* r_lo, dsep_stream_t stream) {
    return gn_act_split_impl(x0, C0, const_cast<double*>(st0), x1, C1, const_cast<double*>(st1), B, P, groups, gamma,
                             beta, eps, act, a_hi, a_lo, r_hi, r_lo, 0, stream);
}

// Same, computing the per-channel sums of x0 (compute_mask & 1) and / or x1 (& 2) in the kernel first and WRITING them
// to st0 / st1 (later consumers of the tensor's statistics read them from there).
extern "C" int dsep_gn_stats_act_split(const float* x0, int C0, double* st0, const float* x1, int C1, double* st1,
                                       int B, int P, int groups, const float* gamma, const float* beta, float eps,
                                       int act, void* a_hi, void* a_lo, void* r_hi, void* r_lo, int compute_mask,
                                       dsep_stream_t stream) {
    DSEP_REQUIRE(compute_mask != 0, "gn_stats_act_split: nothing to compute (use dsep_gn_act_split)");
    return gn_act_split_impl(x0, C0, st0, x1, C1, st1, B, P, groups, gamma, beta, eps, act, a_hi, a_lo, r_hi, r_lo,
                             compute_mask, stream);
}

template <int MODE>
static int launch_fir_tile(const float* x, int B, int H, int W, int C, int groups, const double* st,
                           const float* gamma, const float* beta, float eps, void* a_hi, void* a_lo,
                           void* r_hi, void* r_lo, float* y, float* af, float a8_hi, float a8_lo, cudaStream_t s) {
    using FT = FirTile<MODE>;
    static PerDeviceAttr attr;
    const cudaError_t e = set_max_smem_once(attr, fir_tile_kernel<MODE>, FT::kSmemBytes);
    if (e != cudaSuccess) {
        set_error("fir_resample: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        return DSEP_ERR_CUDA;
    }
    const int Ho = MODE == 1 ? H * 2 : H / 2, Wo = MODE == 1 ? W * 2 : W / 2;
    const int tiles_w = ceil_div(Wo, FT::OW), tiles_h = ceil_div(Ho, FT::OH);
    // tiles per block along w: enough blocks to fill the GPU a few times over, at most 8 tiles each
    static const int kt_env = getenv("DSEP_FIR_KT") ? atoi(getenv("DSEP_FIR_KT")) : 0;
    int kt = kt_env > 0 ? kt_env : 4;
    while (kt > 1 && (int64_t)tiles_h * ceil_div(tiles_w, kt) * (C / 32) * B < 148 * 8) kt >>= 1;
    dim3 grid(tiles_h * ceil_div(tiles_w, kt), C / 32, B);
    fir_tile_kernel<MODE><<<grid, 256, FT::kSmemBytes, s>>>(x, H, W, C, groups, st, gamma, beta, eps, (f16x4*)a_hi,
                                                            (f16x4*)a_lo, (f16x4*)r_hi, (f16x4*)r_lo, (float4*)y,
                                                            (float4*)af, tiles_w, tiles_h, a8_hi, a8_lo, kt);
    return check_launch("fir_tile_kernel");
}

static int fir_resample_impl(const float* x, int B, int H, int W, int C, int mode, int groups,
                             const double* st, const float* gamma, const float* beta, float eps,
                             void* a_hi, void* a_lo, void* r_hi, void* r_lo, float* y, float a8_hi, float a8_lo,
                             dsep_stream_t stream, float* af = nullptr) {
    DSEP_REQUIRE(x, "fir_resample: null input");
    DSEP_REQUIRE(mode == 1 || mode == 2, "fir_resample: mode must be 1 (up) or 2 (down)");
    DSEP_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && B <= 65535, "fir_resample: empty tensor");
    DSEP_REQUIRE(mode == 1 || (H % 2 == 0 && W % 2 == 0), "fir_resample: down needs even H, W");
    DSEP_REQUIRE(a_hi || r_hi || y || af, "fir_resample: no output requested");
    cudaStream_t s = (cudaStream_t)stream;
    const int Ho = mode == 1 ? H * 2 : H / 2, Wo = mode == 1 ? W * 2 : W / 2;
    if (C % 32 != 0) {
        DSEP_REQUIRE(!a_hi && !r_hi && !af && y, "fir_resample: C %% 32 != 0 supports the fp32 output only");
        const int64_t total = (int64_t)B * Ho * Wo * C;
        if (mode == 1) fir_scalar_kernel<1><<<grid_for(total), 256, 0, s>>>(x, total, H, W, C, y);
        else fir_scalar_kernel<2><<<grid_for(total), 256, 0, s>>>(x, total, H, W, C, y);
        return check_launch("fir_scalar_kernel");
    }
    DSEP_REQUIRE((a_hi == nullptr) == (a_lo == nullptr) && (r_hi == nullptr) == (r_lo == nullptr),
                 "fir_resample: hi/lo planes must come in pairs");
    if (st) {
        DSEP_REQUIRE(gamma && beta && (a_hi || af), "fir_resample: GroupNorm branch needs gamma, beta and an output for it");
        int rc = check_gn_shape("fir_resample", C, 0, groups);
        if (rc) return rc;
    }
    return mode == 1 ? launch_fir_tile<1>(x, B, H, W, C, groups, st, gamma, beta, eps, a_hi, a_lo, r_hi, r_lo, y, af,
                                          a8_hi, a8_lo, s)
                     : launch_fir_tile<2>(x, B, H, W, C, groups, st, gamma, beta, eps, a_hi, a_lo, r_hi, r_lo, y, af,
                                          a8_hi, a8_lo, s);
}

extern "C" int dsep_fir_resample(const float* x, int B, int H, int W, int C, int mode, int groups,
                                 const double* st, const float* gamma, const float* beta, float eps,
                                 void* a_hi, void* a_lo, void* r_hi, void* r_lo, float* y,
                                 dsep_stream_t stream) {
    return fir_resample_impl(x, B, H, W, C, mode, groups, st, gamma, beta, eps, a_hi, a_lo, r_hi, r_lo, y, 0.f, 0.f,
                             stream);
}

extern "C" int dsep_fir_resample8(const float* x, int B, int H, int W, int C, int mode, int groups,
                                  const double* st, const float* gamma, const float* beta, float eps,
                                  void* a_hi, void* a_8, void* r_hi, void* r_lo, float* y, int a8_exp,
                                  dsep_stream_t stream) {
    DSEP_REQUIRE(a_hi && a_8 && C % 32 == 0, "fir_resample8: needs the a planes and C %% 32 == 0");
    DSEP_REQUIRE(a8_exp >= -8 && a8_exp <= 8, "fir_resample8: a8_exp out of range");
    return fir_resample_impl(x, B, H, W, C, mode, groups, st, gamma, beta, eps, a_hi, a_8, r_hi, r_lo, y,
                             ldexpf(1.0f, a8_exp), ldexpf(1.0f, a8_exp + 11), stream);
}

// The activated branch FIR(SiLU(GN(x))) as fp32 (for a consumer that builds its operand planes itself: the fused
// convolution with x0 = af and no GroupNorm tables) and / or FIR(x) in y.
extern "C" int dsep_fir_resample_f32(const float* x, int B, int H, int W, int C, int mode, int groups,
                                     const double* st, const float* gamma, const float* beta, float eps, float* af,
                                     float* y, dsep_stream_t stream) {
    DSEP_REQUIRE(af && st && C % 32 == 0, "fir_resample_f32: needs af, the GroupNorm sums and C %% 32 == 0");
    return fir_resample_impl(x, B, H, W, C, mode, groups, st, gamma, beta, eps, nullptr, nullptr, nullptr, nullptr, y,
                             0.f, 0.f, stream, af);
}

extern "C" int dsep_upfirdn2d(const float* in, int planes, int H, int W, int up_x, int up_y, int down_x,
                              int down_y, int pad_x0, int pad_x1, int pad_y0, int pad_y1, float* out,
                              dsep_stream_t stream) {
    DSEP_REQUIRE(in && out, "upfirdn2d: null pointer");
    DSEP_REQUIRE(planes > 0 && H > 0 && W > 0, "upfirdn2d: empty tensor");
    const bool is_up = up_x == 2 && up_y == 2 && down_x == 1 && down_y == 1 && pad_x0 == 2 && pad_x1 == 1 &&
                       pad_y0 == 2 && pad_y1 == 1;
    const bool is_down = up_x == 1 && up_y == 1 && down_x == 2 && down_y == 2 && pad_x0 == 1 && pad_x1 == 1 &&
                         pad_y0 == 1 && pad_y1 == 1 && H % 2 == 0 && W % 2 == 0;
    if (!is_up && !is_down) {
        set_error("upfirdn2d: only the model's two FIR modes are implemented "
                  "(up=2,pad=(2,1)) / (down=2,pad=(1,1)); got up=(%d,%d) down=(%d,%d) pad=(%d,%d,%d,%d)",
                  up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1);
        return DSEP_ERR_UNSUPPORTED;
    }
    cudaStream_t s = (cudaStream_t)stream;
    if (is_up) {
        const int64_t total = (int64_t)planes * H * 2 * W * 2;
        fir_scalar_kernel<1><<<grid_for(total), 256, 0, s>>>(in, total, H, W, 1, out);
    } else {
        const int64_t total = (int64_t)planes * (H / 2) * (W / 2);
        fir_scalar_kernel<2><<<grid_for(total), 256, 0, s>>>(in, total, H, W, 1, out);
    }
    return check_launch("fir_scalar_kernel");
}

extern "C" int dsep_combine(const float* pyr, int Cp, const float* w, const float* bias, const float* h,
                            float* out, int B, int P, int C, double* stats, dsep_stream_t stream) {
    DSEP_REQUIRE(pyr && w && h && out, "combine: null pointer");
    DSEP_REQUIRE(B > 0 && B <= 65535 && P > 0 && C > 0 && C % 4 == 0 && C <= 1024 && Cp > 0 && Cp <= 16,
                 "combine: bad shape");
    const int ppb = pix_per_block_for(P, B);
    dim3 grid(ceil_div(P, ppb), B);
    const size_t smem = sizeof(float) * (C * Cp + C) + (stats ? sizeof(double) * 2 * C : 0);
    combine_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(pyr, Cp, w, bias, h, out, P, C, ppb, stats);
    return check_launch("combine_kernel");
}

// ------------------------------------------------------------------ im2col of the network's input convolution
// The first layer (ncsnpp.py:347-349, conv3x3 of the 2 * (nsrc + 1) = 6 input planes) has K = 54: as a 3x3 conv on the
// tensor core its channel dimension pads 6 -> 64 and nine taps run (0.98 ms at [32,256,256]); as a 1x1 conv over the
// im2col rows col[pix][tap * C + c] (54 -> 64) it is ONE 64-channel K-block.  One float4 of a row per thread.
template <int CT, int CPT>      // compile-time (C, Cp) for the network's shape (6, 64); (0, 0): run-time values
__global__ void __launch_bounds__(256)
im2col3x3_kernel(const float* __restrict__ x, int H, int W, int C_, int Cp_, int64_t total, float4* __restrict__ col) {
    const int C = CT ? CT : C_, Cp = CPT ? CPT : Cp_;
    const int Q = Cp >> 2;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int q = static_cast<int>(e % Q);
        const uint32_t pix = static_cast<uint32_t>(e / Q);          // B * H * W < 2^32
        const uint32_t row = pix / static_cast<uint32_t>(W);        // b * H + h
        const int w = static_cast<int>(pix - row * static_cast<uint32_t>(W));
        const int h = static_cast<int>(row % static_cast<uint32_t>(H));
        float v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int k = q * 4 + i;
            v[i] = 0.f;
            if (k < 9 * C) {
                const int tap = k / C, c = k - tap * C;
                const int dy = tap / 3 - 1, dx = tap - (tap / 3) * 3 - 1;
                const int hh = h + dy, ww = w + dx;
                if (hh >= 0 && hh < H && ww >= 0 && ww < W)
                    v[i] = __ldg(x + (static_cast<int64_t>(pix) + dy * W + dx) * C + c);
            }
        }
        col[e] = make_float4(v[0], v[1], v[2], v[3]);
    }
}

extern "C" int dsep_im2col3x3(const float* x, int B, int H, int W, int C, int Cp, float* col, dsep_stream_t stream) {
    DSEP_REQUIRE(x && col, "im2col3x3: null pointer");
    DSEP_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && Cp % 4 == 0 && 9 * C <= Cp, "im2col3x3: bad shape (9 * C <= Cp, Cp %% 4 == 0)");
    const int64_t total = (int64_t)B * H * W * (Cp / 4);
    DSEP_REQUIRE((int64_t)B * H * W < (int64_t(1) << 32), "im2col3x3: tensor too large");
    const int blocks = grid_for(total, 256, 148 * 64);
    if (C == 6 && Cp == 64)
        im2col3x3_kernel<6, 64><<<blocks, 256, 0, (cudaStream_t)stream>>>(x, H, W, C, Cp, total, (float4*)col);
    else
        im2col3x3_kernel<0, 0><<<blocks, 256, 0, (cudaStream_t)stream>>>(x, H, W, C, Cp, total, (float4*)col);
    return check_launch("im2col3x3_kernel");
}

// ------------------------------------------------------------------ narrow 3x3 convolution, second half
// A 3x3 convolution with a handful of output channels (the output pyramid's 128 / 256 -> 6, ncsnpp.py:419-440) wastes
// the tensor core as a 3x3 problem: the full halo patch is built and nine taps run for 6 of the MMA's 16 ... 128 output
// rows (0.95 ms at [32,256,256,128] against a 0.16 ms HBM floor).  It is run instead as
//     z[pix][tap * CO + co] = sum_ci W[co, ci, tap] * act(x)[pix][ci]          a 1x1 conv, 9 * CO = 54 output channels,
//     out[h, w][co] = bias[co] + residual + sum_tap z[h + ky - 1, w + kx - 1][tap * CO + co]     (this kernel)
// i.e. the taps move from the operand side (a 9x larger K loop over a halo patch) to the output side (a 9-term gather
// of the 1x1 result, each z element read exactly once; out-of-image neighbours are skipped: the conv zero-pads its
// ACTIVATED input, so they contribute nothing).
template <int COT>          // compile-time CO (6: three float2 per tap), 0 = run-time, one thread per (pixel, channel)
__global__ void __launch_bounds__(256)
tap_gather_kernel(const float* __restrict__ z, int H, int W, int ZC, int CO_, const float* __restrict__ bias,
                  const float* __restrict__ residual, float* __restrict__ out, int64_t total) {
    const int CO = COT ? COT : CO_;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t pix = static_cast<uint32_t>(COT ? e : e / CO);          // B * H * W < 2^32
        const int c0 = COT ? 0 : static_cast<int>(e - static_cast<int64_t>(pix) * CO);
        const uint32_t row = pix / static_cast<uint32_t>(W);                    // b * H + h
        const int w = static_cast<int>(pix - row * static_cast<uint32_t>(W));
        const int h = static_cast<int>(row % static_cast<uint32_t>(H));
        constexpr int NV = COT ? COT : 1;
        float acc[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) acc[i] = bias != nullptr ? __ldg(bias + c0 + i) : 0.f;
        if (residual != nullptr) {
#pragma unroll
            for (int i = 0; i < NV; ++i) acc[i] += __ldg(residual + static_cast<int64_t>(pix) * CO + c0 + i);
        }
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
            const int dy = tap / 3 - 1, dx = tap % 3 - 1;
            const int hh = h + dy, ww = w + dx;
            if (hh < 0 || hh >= H || ww < 0 || ww >= W) continue;
            const float* q = z + (static_cast<int64_t>(pix) + dy * W + dx) * ZC + tap * CO + c0;
            if (COT != 0 && COT % 2 == 0) {
#pragma unroll
                for (int i = 0; i < NV / 2; ++i) {
                    const float2 v = __ldg(reinterpret_cast<const float2*>(q) + i);
                    acc[2 * i] += v.x;
                    acc[2 * i + 1] += v.y;
                }
            } else {
#pragma unroll
                for (int i = 0; i < NV; ++i) acc[i] += __ldg(q + i);
            }
        }
        float* o = out + static_cast<int64_t>(pix) * CO + c0;
        if (COT != 0 && COT % 2 == 0) {
#pragma unroll
            for (int i = 0; i < NV / 2; ++i) reinterpret_cast<float2*>(o)[i] = make_float2(acc[2 * i], acc[2 * i + 1]);
        } else {
#pragma unroll
            for (int i = 0; i < NV; ++i) o[i] = acc[i];
        }
    }
}

extern "C" int dsep_tap_gather3x3(const float* z, int B, int H, int W, int ZC, int CO, const float* bias,
                                  const float* residual, float* out, dsep_stream_t stream) {
    DSEP_REQUIRE(z && out, "tap_gather3x3: null pointer");
    DSEP_REQUIRE(B > 0 && H > 0 && W > 0 && CO > 0 && 9 * CO <= ZC, "tap_gather3x3: bad shape (9 * CO <= ZC)");
    DSEP_REQUIRE((int64_t)B * H * W < (int64_t(1) << 32), "tap_gather3x3: tensor too large");
    const int64_t pixels = (int64_t)B * H * W;
    if (CO == 6 && ZC % 2 == 0) {
        tap_gather_kernel<6><<<grid_for(pixels, 256, 148 * 64), 256, 0, (cudaStream_t)stream>>>(
            z, H, W, ZC, CO, bias, residual, out, pixels);
    } else {
        const int64_t total = pixels * CO;
        tap_gather_kernel<0><<<grid_for(total, 256, 148 * 64), 256, 0, (cudaStream_t)stream>>>(
            z, H, W, ZC, CO, bias, residual, out, total);
    }
    return check_launch("tap_gather_kernel");
}

extern "C" int dsep_add(const float* a, const float* b, float* y, int64_t n, dsep_stream_t stream) {
    DSEP_REQUIRE(a && b && y && n >= 0, "add: bad arguments");
    if (n == 0) return DSEP_OK;
    add_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(a, b, y, n);
    return check_launch("add_kernel");
}
