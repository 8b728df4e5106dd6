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

// ---------------------------------------------------------------------------- GN statistics
// grid (chunks, B); thread t owns channel-quad (t % Q) of pixels (t / Q) + k*ppi of its chunk.
__global__ void __launch_bounds__(256)
gn_stats_kernel(const float* __restrict__ x0, int C0, const float* __restrict__ x1, int C1, int P,
                int groups, int pix_per_block, double* __restrict__ stats) {
    __shared__ double s_sum[64], s_sq[64];
    const int Ct = C0 + C1, Q = Ct >> 2, cpg = Ct / groups;
    const int b = blockIdx.y;
    if (threadIdx.x < groups) { s_sum[threadIdx.x] = 0.0; s_sq[threadIdx.x] = 0.0; }
    __syncthreads();
    const int ppi = blockDim.x / Q;
    const int q = threadIdx.x % Q, j = threadIdx.x / Q;
    if (j < ppi) {
        const int c = q * 4;
        const float* src;
        int cs, cl;
        if (c < C0) { src = x0; cs = C0; cl = c; } else { src = x1; cs = C1; cl = c - C0; }
        const int p_begin = blockIdx.x * pix_per_block;
        const int p_end = min(P, p_begin + pix_per_block);
        // the quad may straddle two groups when a group has only 2 channels (nf = 64)
        double s01 = 0.0, ss01 = 0.0, s23 = 0.0, ss23 = 0.0;
        for (int p = p_begin + j; p < p_end; p += ppi) {
            const float4 v = *reinterpret_cast<const float4*>(src + (static_cast<size_t>(b) * P + p) * cs + cl);
            s01 += (double)v.x + (double)v.y;
            s23 += (double)v.z + (double)v.w;
            ss01 += (double)v.x * v.x + (double)v.y * v.y;
            ss23 += (double)v.z * v.z + (double)v.w * v.w;
        }
        const int g0 = c / cpg, g1 = (c + 2) / cpg;
        if (g0 == g1) {
            atomicAdd(&s_sum[g0], s01 + s23);
            atomicAdd(&s_sq[g0], ss01 + ss23);
        } else {
            atomicAdd(&s_sum[g0], s01); atomicAdd(&s_sq[g0], ss01);
            atomicAdd(&s_sum[g1], s23); atomicAdd(&s_sq[g1], ss23);
        }
    }
    __syncthreads();
    if (threadIdx.x < groups) {
        atomicAdd(&stats[(static_cast<size_t>(b) * groups + threadIdx.x) * 2 + 0], s_sum[threadIdx.x]);
        atomicAdd(&stats[(static_cast<size_t>(b) * groups + threadIdx.x) * 2 + 1], s_sq[threadIdx.x]);
    }
}

__device__ __forceinline__ void gn_finalize(const double* stats, int b, int groups, int g, double count,
                                            float eps, float& mean, float& rstd) {
    const double s = stats[(static_cast<size_t>(b) * groups + g) * 2 + 0];
    const double ss = stats[(static_cast<size_t>(b) * groups + g) * 2 + 1];
    const double m = s / count;
    double var = ss / count - m * m;
    if (var < 0.0) var = 0.0;
    mean = static_cast<float>(m);
    rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
}

// ------------------------------------------------------- GN + act + split (+ raw split)
// grid (chunks, B); each thread converts one channel-quad of one pixel per iteration.
__global__ void __launch_bounds__(256)
gn_act_split_kernel(const float* __restrict__ x0, int C0, const float* __restrict__ x1, int C1, int P,
                    int groups, const double* __restrict__ stats, const float* __restrict__ gamma,
                    const float* __restrict__ beta, float eps, int act, f16x4* __restrict__ a_hi,
                    f16x4* __restrict__ a_lo, f16x4* __restrict__ r_hi, f16x4* __restrict__ r_lo,
                    int pix_per_block) {
    __shared__ float s_mean[64], s_rstd[64];
    const int Ct = C0 + C1, Q = Ct >> 2, cpg = Ct / groups;
    const int b = blockIdx.y;
    if (stats != nullptr && threadIdx.x < groups)
        gn_finalize(stats, b, groups, threadIdx.x, (double)P * cpg, eps, s_mean[threadIdx.x],
                    s_rstd[threadIdx.x]);
    __syncthreads();
    const int64_t e_begin = (int64_t)blockIdx.x * pix_per_block * Q;
    const int64_t e_end = min((int64_t)P * Q, e_begin + (int64_t)pix_per_block * Q);
    for (int64_t e = e_begin + threadIdx.x; e < e_end; e += blockDim.x) {
        const int p = static_cast<int>(e / Q), q = static_cast<int>(e % Q);
        const int c = q * 4;
        const float4 v = c < C0
            ? *reinterpret_cast<const float4*>(x0 + (static_cast<size_t>(b) * P + p) * C0 + c)
            : *reinterpret_cast<const float4*>(x1 + (static_cast<size_t>(b) * P + p) * C1 + (c - C0));
        const size_t o = (static_cast<size_t>(b) * P + p) * Q + q;
        f16x4 h, l;
        if (r_hi != nullptr) {
            split4(v, h, l);
            r_hi[o] = h;
            r_lo[o] = l;
        }
        if (a_hi != nullptr) {
            float4 y = v;
            if (stats != nullptr) {
                const int g0 = c / cpg, g1 = (c + 2) / cpg;
                const float m0 = s_mean[g0], rs0 = s_rstd[g0], m1 = s_mean[g1], rs1 = s_rstd[g1];
                const float4 ga = *reinterpret_cast<const float4*>(gamma + c);
                const float4 be = *reinterpret_cast<const float4*>(beta + c);
                y.x = (v.x - m0) * rs0 * ga.x + be.x;
                y.y = (v.y - m0) * rs0 * ga.y + be.y;
                y.z = (v.z - m1) * rs1 * ga.z + be.z;
                y.w = (v.w - m1) * rs1 * ga.w + be.w;
            }
            if (act == 1) { y.x = silu_f(y.x); y.y = silu_f(y.y); y.z = silu_f(y.z); y.w = silu_f(y.w); }
            split4(y, h, l);
            a_hi[o] = h;
            a_lo[o] = l;
        }
    }
}

// --------------------------------------------------------------------------- FIR resample
// taps [1,3,3,1]; down: y[i] = (x[2i-1] + 3x[2i] + 3x[2i+1] + x[2i+2]) / 8 per axis;
// up (gain 2 per axis): y[2i] = (x[i-1] + 3x[i]) / 4, y[2i+1] = (3x[i] + x[i+1]) / 4; zeros outside.
template <int MODE>   // 1 up, 2 down
__global__ void __launch_bounds__(256)
fir_quad_kernel(const float* __restrict__ x, int H, int W, int C, int groups,
                const double* __restrict__ stats, const float* __restrict__ gamma,
                const float* __restrict__ beta, float eps, f16x4* __restrict__ a_hi,
                f16x4* __restrict__ a_lo, f16x4* __restrict__ r_hi, f16x4* __restrict__ r_lo,
                float4* __restrict__ y, int pix_per_block) {
    __shared__ float s_mean[64], s_rstd[64];
    const int Q = C >> 2;
    const int b = blockIdx.y;
    const bool xf = stats != nullptr && a_hi != nullptr;
    const int cpg = xf ? C / groups : 1;
    if (xf && threadIdx.x < groups)
        gn_finalize(stats, b, groups, threadIdx.x, (double)H * W * cpg, eps, s_mean[threadIdx.x],
                    s_rstd[threadIdx.x]);
    __syncthreads();
    const int Ho = MODE == 1 ? H * 2 : H / 2, Wo = MODE == 1 ? W * 2 : W / 2;
    const int64_t e_begin = (int64_t)blockIdx.x * pix_per_block * Q;
    const int64_t e_end = min((int64_t)Ho * Wo * Q, e_begin + (int64_t)pix_per_block * Q);
    const bool want_raw = (r_hi != nullptr) || (y != nullptr);
    for (int64_t e = e_begin + threadIdx.x; e < e_end; e += blockDim.x) {
        const int q = static_cast<int>(e % Q);
        const int po = static_cast<int>(e / Q);
        const int oi = po / Wo, oj = po % Wo;
        const int c = q * 4;
        float m0 = 0.f, rs0 = 1.f, m1 = 0.f, rs1 = 1.f;
        float4 ga = make_float4(1.f, 1.f, 1.f, 1.f), be = make_float4(0.f, 0.f, 0.f, 0.f);
        if (xf) {
            const int g0 = c / cpg, g1 = (c + 2) / cpg;
            m0 = s_mean[g0]; rs0 = s_rstd[g0]; m1 = s_mean[g1]; rs1 = s_rstd[g1];
            ga = *reinterpret_cast<const float4*>(gamma + c);
            be = *reinterpret_cast<const float4*>(beta + c);
        }
        float4 acc_r = make_float4(0.f, 0.f, 0.f, 0.f), acc_a = acc_r;
        constexpr int NTAP = MODE == 1 ? 2 : 4;
        int i0, j0;
        float wi[NTAP], wj[NTAP];
        if (MODE == 1) {
            // even output 2i: taps (i-1: 1/4, i: 3/4); odd 2i+1: (i: 3/4, i+1: 1/4)
            const int ii = oi >> 1, jj = oj >> 1;
            if (oi & 1) { i0 = ii; wi[0] = 0.75f; wi[1] = 0.25f; } else { i0 = ii - 1; wi[0] = 0.25f; wi[1] = 0.75f; }
            if (oj & 1) { j0 = jj; wj[0] = 0.75f; wj[1] = 0.25f; } else { j0 = jj - 1; wj[0] = 0.25f; wj[1] = 0.75f; }
        } else {
            i0 = 2 * oi - 1; j0 = 2 * oj - 1;
            wi[0] = 0.125f; wi[1] = 0.375f; wi[2] = 0.375f; wi[3] = 0.125f;
            wj[0] = 0.125f; wj[1] = 0.375f; wj[2] = 0.375f; wj[3] = 0.125f;
        }
#pragma unroll
        for (int a = 0; a < NTAP; ++a) {
            const int i = i0 + a;
            if (i < 0 || i >= H) continue;
#pragma unroll
            for (int d = 0; d < NTAP; ++d) {
                const int j = j0 + d;
                if (j < 0 || j >= W) continue;
                const float wgt = wi[a] * wj[d];
                const float4 v = *reinterpret_cast<const float4*>(
                    x + ((static_cast<size_t>(b) * H + i) * W + j) * C + c);
                if (want_raw) {
                    acc_r.x += wgt * v.x; acc_r.y += wgt * v.y; acc_r.z += wgt * v.z; acc_r.w += wgt * v.w;
                }
                if (a_hi != nullptr) {
                    float4 t = v;
                    if (xf) {
                        t.x = (v.x - m0) * rs0 * ga.x + be.x; t.y = (v.y - m0) * rs0 * ga.y + be.y;
                        t.z = (v.z - m1) * rs1 * ga.z + be.z; t.w = (v.w - m1) * rs1 * ga.w + be.w;
                        t.x = silu_f(t.x); t.y = silu_f(t.y); t.z = silu_f(t.z); t.w = silu_f(t.w);
                    }
                    acc_a.x += wgt * t.x; acc_a.y += wgt * t.y; acc_a.z += wgt * t.z; acc_a.w += wgt * t.w;
                }
            }
        }
        const size_t o = (static_cast<size_t>(b) * Ho * Wo + po) * Q + q;
        f16x4 h, l;
        if (a_hi != nullptr) { split4(acc_a, h, l); a_hi[o] = h; a_lo[o] = l; }
        if (r_hi != nullptr) { split4(acc_r, h, l); r_hi[o] = h; r_lo[o] = l; }
        if (y != nullptr) y[o] = acc_r;
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
__global__ void __launch_bounds__(256)
combine_kernel(const float* __restrict__ pyr, int Cp, const float* __restrict__ w,
               const float* __restrict__ bias, const float* __restrict__ h, float* __restrict__ out,
               int64_t total_quads, int C) {
    extern __shared__ float s_w[];   // [C * Cp] + [C]
    float* s_b = s_w + C * Cp;
    for (int i = threadIdx.x; i < C * Cp; i += blockDim.x) s_w[i] = w[i];
    for (int i = threadIdx.x; i < C; i += blockDim.x) s_b[i] = bias ? bias[i] : 0.f;
    __syncthreads();
    const int Q = C >> 2;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total_quads;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int q = static_cast<int>(e % Q);
        const int64_t pix = e / Q;
        const int c = q * 4;
        float4 v = reinterpret_cast<const float4*>(h)[e];
        float a[4] = {s_b[c], s_b[c + 1], s_b[c + 2], s_b[c + 3]};
        for (int k = 0; k < Cp; ++k) {
            const float pk = pyr[pix * Cp + k];
#pragma unroll
            for (int u = 0; u < 4; ++u) a[u] += s_w[(c + u) * Cp + k] * pk;
        }
        v.x += a[0]; v.y += a[1]; v.z += a[2]; v.w += a[3];
        reinterpret_cast<float4*>(out)[e] = v;
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
    DSEP_REQUIRE(groups > 0 && groups <= 64 && Ct % groups == 0 &&
                     (Ct / groups) % 2 == 0,
                 "%s: unsupported groups=%d for %d channels", who, groups, Ct);
    DSEP_REQUIRE(Ct / 4 <= 256, "%s: at most 1024 channels", who);
    return DSEP_OK;
}

extern "C" int dsep_gn_stats(const float* x0, int C0, const float* x1, int C1, int B, int P, int groups,
                             double* stats, dsep_stream_t stream) {
    DSEP_REQUIRE(x0 && stats && (C1 == 0 || x1), "gn_stats: null pointer");
    DSEP_REQUIRE(B > 0 && P > 0, "gn_stats: empty tensor");
    int rc = check_gn_shape("gn_stats", C0, C1, groups);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    cudaMemsetAsync(stats, 0, sizeof(double) * 2 * B * groups, s);
    const int ppb = pix_per_block_for(P, B);
    dim3 grid(ceil_div(P, ppb), B);
    gn_stats_kernel<<<grid, 256, 0, s>>>(x0, C0, x1, C1, P, groups, ppb, stats);
    return check_launch("gn_stats_kernel");
}

extern "C" int dsep_gn_act_split(const float* x0, int C0, const float* x1, int C1, int B, int P,
                                 int groups, const double* stats, const float* gamma, const float* beta,
                                 float eps, int act, void* a_hi, void* a_lo, void* r_hi, void* r_lo,
                                 dsep_stream_t stream) {
    DSEP_REQUIRE(x0 && (C1 == 0 || x1), "gn_act_split: null input");
    DSEP_REQUIRE((a_hi && a_lo) || (r_hi && r_lo), "gn_act_split: no output requested");
    DSEP_REQUIRE(stats == nullptr || (gamma && beta), "gn_act_split: stats without gamma/beta");
    DSEP_REQUIRE(act == 0 || act == 1, "gn_act_split: act must be 0 or 1");
    DSEP_REQUIRE(B > 0 && P > 0, "gn_act_split: empty tensor");
    DSEP_REQUIRE(C0 > 0 && C1 >= 0 && C0 % 4 == 0 && C1 % 4 == 0,
                 "gn_act_split: channel counts must be multiples of 4");
    if (stats) {
        int rc = check_gn_shape("gn_act_split", C0, C1, groups);
        if (rc) return rc;
    }
    const int ppb = pix_per_block_for(P, B);
    dim3 grid(ceil_div(P, ppb), B);
    gn_act_split_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
        x0, C0, x1, C1, P, stats ? groups : 1, stats, gamma, beta, eps, act, (f16x4*)a_hi, (f16x4*)a_lo,
        (f16x4*)r_hi, (f16x4*)r_lo, ppb);
    return check_launch("gn_act_split_kernel");
}

extern "C" int dsep_fir_resample(const float* x, int B, int H, int W, int C, int mode, int groups,
                                 const double* stats, const float* gamma, const float* beta, float eps,
                                 void* a_hi, void* a_lo, void* r_hi, void* r_lo, float* y,
                                 dsep_stream_t stream) {
    DSEP_REQUIRE(x, "fir_resample: null input");
    DSEP_REQUIRE(mode == 1 || mode == 2, "fir_resample: mode must be 1 (up) or 2 (down)");
    DSEP_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0, "fir_resample: empty tensor");
    DSEP_REQUIRE(mode == 1 || (H % 2 == 0 && W % 2 == 0), "fir_resample: down needs even H, W");
    DSEP_REQUIRE(a_hi || r_hi || y, "fir_resample: no output requested");
    cudaStream_t s = (cudaStream_t)stream;
    const int Ho = mode == 1 ? H * 2 : H / 2, Wo = mode == 1 ? W * 2 : W / 2;
    if (C % 4 != 0) {
        DSEP_REQUIRE(!a_hi && !r_hi && y, "fir_resample: C %% 4 != 0 supports the fp32 output only");
        const int64_t total = (int64_t)B * Ho * Wo * C;
        if (mode == 1) fir_scalar_kernel<1><<<grid_for(total), 256, 0, s>>>(x, total, H, W, C, y);
        else fir_scalar_kernel<2><<<grid_for(total), 256, 0, s>>>(x, total, H, W, C, y);
        return check_launch("fir_scalar_kernel");
    }
    DSEP_REQUIRE((a_hi == nullptr) == (a_lo == nullptr) && (r_hi == nullptr) == (r_lo == nullptr),
                 "fir_resample: hi/lo planes must come in pairs");
    if (stats) {
        DSEP_REQUIRE(gamma && beta && a_hi, "fir_resample: GroupNorm branch needs gamma, beta and a_hi/a_lo");
        int rc = check_gn_shape("fir_resample", C, 0, groups);
        if (rc) return rc;
    }
    const int ppb = pix_per_block_for(Ho * Wo, B);
    dim3 grid(ceil_div(Ho * Wo, ppb), B);
    if (mode == 1)
        fir_quad_kernel<1><<<grid, 256, 0, s>>>(x, H, W, C, groups, stats, gamma, beta, eps, (f16x4*)a_hi,
                                                (f16x4*)a_lo, (f16x4*)r_hi, (f16x4*)r_lo, (float4*)y, ppb);
    else
        fir_quad_kernel<2><<<grid, 256, 0, s>>>(x, H, W, C, groups, stats, gamma, beta, eps, (f16x4*)a_hi,
                                                (f16x4*)a_lo, (f16x4*)r_hi, (f16x4*)r_lo, (float4*)y, ppb);
    return check_launch("fir_quad_kernel");
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
                            float* out, int B, int P, int C, dsep_stream_t stream) {
    DSEP_REQUIRE(pyr && w && h && out, "combine: null pointer");
    DSEP_REQUIRE(B > 0 && P > 0 && C > 0 && C % 4 == 0 && Cp > 0 && Cp <= 16, "combine: bad shape");
    const int64_t total = (int64_t)B * P * (C / 4);
    const size_t smem = sizeof(float) * (C * Cp + C);
    combine_kernel<<<grid_for(total), 256, smem, (cudaStream_t)stream>>>(pyr, Cp, w, bias, h, out, total, C);
    return check_launch("combine_kernel");
}

extern "C" int dsep_add(const float* a, const float* b, float* y, int64_t n, dsep_stream_t stream) {
    DSEP_REQUIRE(a && b && y && n >= 0, "add: bad arguments");
    if (n == 0) return DSEP_OK;
    add_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(a, b, y, n);
    return check_launch("add_kernel");
}
