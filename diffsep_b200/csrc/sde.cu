// SDE arithmetic of the reverse-diffusion loop, one fused element-wise kernel per update.
//
// Restates (reference file:line)
//   MixSDE / PriorMixSDE  _cov_eigval, _std, mult_std, sde, prior_sampling   sdes/sdes.py:275-346, 451-587
//   SDE.discretize / RSDE.discretize                                          sdes/sdes.py:93-107, 163-171
//   ReverseDiffusionPredictor.update_fn                                       sdes/predictors.py:60-66
//   AnnealedLangevinDynamics2.update_fn                                       sdes/correctors.py:109-128
//   normalize_batch                                                           pl_model.py:81-88
//   scale_output                                                              separate.py:73-78
//
// With A the channel-averaging matrix and Pn = I - A (sdes.py:242-248) every matrix the reference
// builds is a A + b Pn, so  (a A + b Pn) v = a vbar + b (v - vbar)  with vbar the channel mean:
// the [B,n,n] / [B,n,n,T] einsums collapse to two FMAs per element.  x is [B, ndim, T] (ndim = 2, or 3 for the
// 3-speaker models); one thread owns VEC consecutive samples of ALL channels, so the channel mean never leaves registers and
// every access is a coalesced 4*VEC-byte vector.  The reference launches ~8-15 kernels per update
// (pow, exp, sqrt, einsum, randn_like, axpy ...); here it is one, HBM-bound: 4-5 arrays read,
// 2 written.
#include "common.cuh"

namespace dsep {

// ---------------------------------------------------------------- Philox4x32-10 + Box-Muller
struct Philox {
    uint32_t c[4];
};
__device__ __forceinline__ Philox philox4x32_10(uint64_t counter, uint64_t offset, uint64_t seed) {
    uint32_t c0 = static_cast<uint32_t>(counter), c1 = static_cast<uint32_t>(counter >> 32);
    uint32_t c2 = static_cast<uint32_t>(offset), c3 = static_cast<uint32_t>(offset >> 32);
    uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    Philox p;
    p.c[0] = c0; p.c[1] = c1; p.c[2] = c2; p.c[3] = c3;
    return p;
}
// 4 standard normals for the element quad `quad` of a stream identified by (seed, offset).
__device__ __forceinline__ void normal4(uint64_t quad, uint64_t offset, uint64_t seed, float (&z)[4]) {
    const Philox p = philox4x32_10(quad, offset, seed);
    const float k = 2.3283064365386963e-10f;   // 2^-32
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float u1 = (static_cast<float>(p.c[2 * i]) + 0.5f) * k;       // (0, 1]
        const float u2 = (static_cast<float>(p.c[2 * i + 1]) + 0.5f) * k;
        const float r = sqrtf(-2.0f * logf(fminf(u1, 1.0f)));
        float s, c;
        sincospif(2.0f * u2, &s, &c);
        z[2 * i] = r * c;
        z[2 * i + 1] = r * s;
    }
}

template <int VEC>
struct Vec {
    float v[VEC];
};
template <int VEC>
__device__ __forceinline__ Vec<VEC> ldv(const float* p) {
    Vec<VEC> r;
    if constexpr (VEC == 4) {
        const float4 t = *reinterpret_cast<const float4*>(p);
        r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
    } else {
        r.v[0] = *p;
    }
    return r;
}
template <int VEC>
__device__ __forceinline__ void stv(float* p, const Vec<VEC>& r) {
    if constexpr (VEC == 4) *reinterpret_cast<float4*>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
    else *p = r.v[0];
}
// noise for elements [e, e+VEC) of a [B,2,T] array: injected tensor or Philox stream
template <int VEC>
__device__ __forceinline__ Vec<VEC> noise_at(const float* noise, int64_t e, uint64_t seed, uint64_t offset) {
    if (noise != nullptr) return ldv<VEC>(noise + e);
    Vec<VEC> r;
    float z[4];
    normal4(static_cast<uint64_t>(e) >> 2, offset, seed, z);
    if constexpr (VEC == 4) {
#pragma unroll
        for (int i = 0; i < 4; ++i) r.v[i] = z[i];
    } else {
        r.v[0] = z[e & 3];
    }
    return r;
}

struct SdeScalars {
    float s1, s2;   // sqrt(ev1), sqrt(ev2) of the marginal covariance at t
    float g;        // diffusion coefficient g(t)
};
// sdes.py:296-310 (eigenvalues) and :282-283 (diffusion), evaluated in fp32 like the reference.
__device__ __forceinline__ SdeScalars sde_scalars(const dsep_sde_params& p, float t) {
    const float ratiosig = p.sigma_max / p.sigma_min;
    const float logsig = logf(ratiosig);
    const float mult = p.sigma_min * p.sigma_min;
    const float srp = powf(ratiosig, 2.0f * t);
    const float ev1 = mult * (srp - 1.0f);
    const float ev2 = mult * (srp - expf(-2.0f * p.d_lambda * t)) / (1.0f + p.d_lambda / logsig);
    SdeScalars s;
    s.s1 = sqrtf(fmaxf(ev1, 0.0f));
    s.s2 = sqrtf(fmaxf(ev2, 0.0f));
    s.g = p.sigma_min * powf(ratiosig, t) * sqrtf(2.0f * logsig);
    return s;
}

// (a A + b Pn) applied to the NC channel values v[c] of one sample: a vbar + b (v - vbar)
template <int NC>
__device__ __forceinline__ float chan_mean(const float (&v)[NC]) {
    if (NC == 2) return 0.5f * (v[0] + v[1]);
    float s = 0.0f;
#pragma unroll
    for (int c = 0; c < NC; ++c) s += v[c];
    return s * (1.0f / NC);
}
template <int NC>
__device__ __forceinline__ void apply_std(float a, float b, const float (&v)[NC], float (&o)[NC]) {
    const float m = chan_mean<NC>(v);
#pragma unroll
    for (int c = 0; c < NC; ++c) o[c] = a * m + b * (v[c] - m);
}

// MODE 0: prior, 1: ald2 corrector, 2: predictor, 3: ald corrector.   grid (ceil(T/VEC/256), B).  NC = ndim sources
// (2, or 3 for the 3-speaker models: every matrix is still a A + b Pn with A = ones(NC, NC) / NC, sdes.py:242-248).
// MODE 0 only: `flag` = channels of mix (1: broadcast, NC: the true_mean branch of prior_sampling, sdes.py:571-583),
// coef = factor on the mean (0.5 for a 1-channel mixture and always for MixSDE, :344; 1 for PriorMixSDE's NC-channel
// branch), sig_ch = channels of sigma_mix (NC when it was computed from an NC-channel input: L = (...)[c,d] sigma[d]).
template <int MODE, int VEC, int NC>
__global__ void __launch_bounds__(256)
sde_update_kernel(const dsep_sde_params p, const float* __restrict__ x, const float* __restrict__ score,
                  const float* __restrict__ mix, const float* __restrict__ tvec,
                  const float* __restrict__ sigma_mix, const float* __restrict__ noise, uint64_t seed,
                  uint64_t offset, float coef, int flag, int sig_ch, int T, float* __restrict__ x_out,
                  float* __restrict__ x_mean) {
    const int b = blockIdx.y;
    const int t0 = (blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    if (t0 >= T) return;
    const SdeScalars sc = sde_scalars(p, MODE == 0 ? p.T_end : tvec[b]);
    int64_t e[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) e[c] = (static_cast<int64_t>(b) * NC + c) * T + t0;
    Vec<VEC> sm[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) sm[c].v[i] = 1.0f;
    }
    if (sigma_mix != nullptr) {
        if (MODE == 0 && sig_ch == NC) {
#pragma unroll
            for (int c = 0; c < NC; ++c) sm[c] = ldv<VEC>(sigma_mix + e[c]);
        } else {
            sm[0] = ldv<VEC>(sigma_mix + static_cast<int64_t>(b) * T + t0);
#pragma unroll
            for (int c = 1; c < NC; ++c) sm[c] = sm[0];
        }
    }
    Vec<VEC> z[NC], o[NC], m[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) z[c] = noise_at<VEC>(noise, e[c], seed, offset);
    if (MODE == 0) {
        Vec<VEC> mx[NC];
        if (flag == NC) {
#pragma unroll
            for (int c = 0; c < NC; ++c) mx[c] = ldv<VEC>(mix + e[c]);
        } else {
            mx[0] = ldv<VEC>(mix + static_cast<int64_t>(b) * T + t0);
#pragma unroll
            for (int c = 1; c < NC; ++c) mx[c] = mx[0];
        }
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            float w[NC], l[NC];
#pragma unroll
            for (int c = 0; c < NC; ++c) w[c] = z[c].v[i] * sm[c].v[i];      // L = (s1 A + s2 Pn) diag(sigma)
            apply_std<NC>(sc.s1, sc.s2, w, l);
#pragma unroll
            for (int c = 0; c < NC; ++c) o[c].v[i] = coef * mx[c].v[i] + l[c];
        }
#pragma unroll
        for (int c = 0; c < NC; ++c) stv<VEC>(x_out + e[c], o[c]);
        return;
    }
    Vec<VEC> xv[NC], sv[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) { xv[c] = ldv<VEC>(x + e[c]); sv[c] = ldv<VEC>(score + e[c]); }
    if (MODE == 1) {
        // ald2: x_mean = x + 2 snr^2 L (L s);  x' = x_mean + 2 snr L z        (coef = snr)
        const float c2 = 2.0f * coef * coef, c1 = 2.0f * coef;
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            const float sg = sm[0].v[i];
            float a[NC], l[NC], g[NC], n[NC];
#pragma unroll
            for (int c = 0; c < NC; ++c) a[c] = sv[c].v[i];
            apply_std<NC>(sc.s1, sc.s2, a, l);
#pragma unroll
            for (int c = 0; c < NC; ++c) l[c] *= sg;
            apply_std<NC>(sc.s1, sc.s2, l, g);
#pragma unroll
            for (int c = 0; c < NC; ++c) a[c] = z[c].v[i];
            apply_std<NC>(sc.s1, sc.s2, a, n);
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                m[c].v[i] = xv[c].v[i] + c2 * (g[c] * sg);
                o[c].v[i] = m[c].v[i] + c1 * (n[c] * sg);
            }
        }
    } else if (MODE == 3) {
        // ald (original annealed Langevin, correctors.py:58-91): std = sqrt of the first-row sum of the
        // covariance = sqrt(ev1); step = 2 (snr std)^2; x_mean = x + step s; x' = x_mean + sqrt(2 step) z
        const float step = 2.0f * (coef * sc.s1) * (coef * sc.s1), nz = sqrtf(2.0f * step);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                m[c].v[i] = xv[c].v[i] + step * sv[c].v[i];
                o[c].v[i] = m[c].v[i] + nz * z[c].v[i];
            }
        }
    } else {
        // reverse diffusion: f = -lambda (x - xbar) dt, G = g sqrt(dt);  x_mean = x - (f - c G^2 s);
        // x' = x_mean + G z   (coef = dt).  probability flow (flag): c = 1/2 and no noise
        // (sdes.py:143-152,167-170); otherwise c = 1.
        const float dt = coef, sq = sqrtf(dt);
        const float cs = flag ? 0.5f : 1.0f, cz = flag ? 0.0f : 1.0f;
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            float a[NC];
#pragma unroll
            for (int c = 0; c < NC; ++c) a[c] = xv[c].v[i];
            const float xb = chan_mean<NC>(a);
            const float G = sc.g * sm[0].v[i] * sq;
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                const float f = -p.d_lambda * (a[c] - xb) * dt;
                m[c].v[i] = a[c] - (f - cs * G * G * sv[c].v[i]);
                o[c].v[i] = m[c].v[i] + cz * G * z[c].v[i];
            }
        }
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) stv<VEC>(x_out + e[c], o[c]);
    if (x_mean != nullptr) {
#pragma unroll
        for (int c = 0; c < NC; ++c) stv<VEC>(x_mean + e[c], m[c]);
    }
}

// ---------------------------------------------------------------- sigma_mix (PriorMixSDE)
// out[b,t] = 0.5 sqrt(max(sum_{j=t-k/2}^{t-k/2+k-1} mix[b,j]^2 / k, 1e-4)), zeros outside [0,T)
// (avg_pool1d counts the padding; for even k the extra last sample is dropped, sdes.py:480-487).
constexpr int kSigTile = 1024;
__global__ void __launch_bounds__(256)
sigma_mix_kernel(const float* __restrict__ mix, int T, int k, float* __restrict__ sigma) {
    extern __shared__ float s_sq[];   // kSigTile + k
    const int b = blockIdx.y;
    const int t_begin = blockIdx.x * kSigTile;
    const int half = k / 2;
    const float* src = mix + static_cast<int64_t>(b) * T;
    for (int i = threadIdx.x; i < kSigTile + k; i += blockDim.x) {
        const int j = t_begin - half + i;
        const float v = (j >= 0 && j < T) ? src[j] : 0.0f;
        s_sq[i] = v * v;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kSigTile; i += blockDim.x) {
        const int t = t_begin + i;
        if (t >= T) break;
        float acc = 0.0f;
        for (int j = 0; j < k; ++j) acc += s_sq[i + j];
        sigma[static_cast<int64_t>(b) * T + t] = 0.5f * sqrtf(fmaxf(acc / static_cast<float>(k), 1e-4f));
    }
}

// ---------------------------------------------------------------- block reductions (double)
__device__ __forceinline__ double block_sum(double v, double* s_red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    double t = 0.0;
    for (int i = 0; i < (blockDim.x >> 5); ++i) t += s_red[i];
    return t;
}

// one block per utterance: mean / unbiased std over n samples, out = (x - mean) / max(std, 1e-5)
__global__ void __launch_bounds__(1024)
normalize_kernel(const float* __restrict__ x, int n, float* __restrict__ out, float* __restrict__ mean_o,
                 float* __restrict__ std_o) {
    __shared__ double s_red[32];
    const float* src = x + static_cast<int64_t>(blockIdx.x) * n;
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += src[i];
    const double mean = block_sum(s, s_red) / n;
    double ss = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double d = src[i] - mean;
        ss += d * d;
    }
    const double var = block_sum(ss, s_red) / (n > 1 ? n - 1 : 1);
    const float mean_f = static_cast<float>(mean);
    const float std_f = fmaxf(static_cast<float>(sqrt(var)), 1e-5f);
    float* dst = out + static_cast<int64_t>(blockIdx.x) * n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = (src[i] - mean_f) / std_f;
    if (threadIdx.x == 0) {
        if (mean_o) mean_o[blockIdx.x] = mean_f;
        if (std_o) std_o[blockIdx.x] = std_f;
    }
}

// one block per (utterance, source): alpha = <mix, sep> / sum(sep^2 + 1e-10); out = alpha sep
__global__ void __launch_bounds__(1024)
scale_output_kernel(const float* __restrict__ mix, const float* __restrict__ sep, int nsrc, int T,
                    float* __restrict__ out) {
    __shared__ double s_red[32];
    const int b = blockIdx.x / nsrc;
    const float* m = mix + static_cast<int64_t>(b) * T;
    const float* s = sep + static_cast<int64_t>(blockIdx.x) * T;
    double num = 0.0, den = 0.0;
    for (int i = threadIdx.x; i < T; i += blockDim.x) {
        const float v = s[i];
        num += static_cast<double>(m[i]) * v;
        den += static_cast<double>(v) * v + 1e-10;
    }
    num = block_sum(num, s_red);
    den = block_sum(den, s_red);
    const float alpha = static_cast<float>(num / den);
    float* dst = out + static_cast<int64_t>(blockIdx.x) * T;
    for (int i = threadIdx.x; i < T; i += blockDim.x) dst[i] = alpha * s[i];
}

// ---------------------------------------------------------------- LangevinCorrector (correctors.py:35-55)
// norms[0][b] = ||score_b||, norms[1][b] = ||noise_b||; step = 2 (snr mean_b||noise_b|| / mean_b||score_b||)^2
// couples the batch entries through the two batch means, exactly like the reference.
__global__ void __launch_bounds__(1024)
item_norms_kernel(const float* __restrict__ score, const float* __restrict__ noise, int n, float* __restrict__ norms) {
    __shared__ double s_red[32];
    const float* src = (blockIdx.y == 0 ? score : noise) + static_cast<int64_t>(blockIdx.x) * n;
    double ss = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) ss += static_cast<double>(src[i]) * src[i];
    ss = block_sum(ss, s_red);
    if (threadIdx.x == 0) norms[blockIdx.y * gridDim.x + blockIdx.x] = static_cast<float>(sqrt(ss));
}

__global__ void __launch_bounds__(256)
langevin_kernel(const float* __restrict__ x, const float* __restrict__ score, const float* __restrict__ noise,
                const float* __restrict__ norms, float snr, int B, int64_t total, float* __restrict__ x_out,
                float* __restrict__ x_mean) {
    __shared__ float s_step;
    if (threadIdx.x == 0) {
        float g = 0.f, z = 0.f;
        for (int b = 0; b < B; ++b) { g += norms[b]; z += norms[B + b]; }
        const float r = snr * (z / B) / (g / B);
        s_step = r * r * 2.0f;
    }
    __syncthreads();
    const float step = s_step, nz = sqrtf(step * 2.0f);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const float m = x[i] + step * score[i];
        x_mean[i] = m;
        x_out[i] = m + nz * noise[i];
    }
}

__global__ void randn_kernel(float* __restrict__ z, int64_t n, uint64_t seed, uint64_t offset) {
    const int64_t quads = (n + 3) / 4;
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < quads;
         q += (int64_t)gridDim.x * blockDim.x) {
        float v[4];
        normal4(static_cast<uint64_t>(q), offset, seed, v);
        for (int i = 0; i < 4; ++i)
            if (q * 4 + i < n) z[q * 4 + i] = v[i];
    }
}

template <int MODE, int NC>
static int launch_update_nc(const dsep_sde_params* p, const float* x, const float* score, const float* mix,
                            const float* t, const float* sigma_mix, const float* noise, uint64_t seed,
                            uint64_t offset, float coef, int flag, int sig_ch, int B, int T, float* x_out,
                            float* x_mean, cudaStream_t s) {
    auto aligned = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
    const bool vec = (T % 4 == 0) && aligned(x) && aligned(score) && aligned(mix) && aligned(sigma_mix) &&
                     aligned(noise) && aligned(x_out) && aligned(x_mean);
    if (vec) {
        dim3 grid(ceil_div(T / 4, 256), B);
        sde_update_kernel<MODE, 4, NC><<<grid, 256, 0, s>>>(*p, x, score, mix, t, sigma_mix, noise, seed, offset,
                                                            coef, flag, sig_ch, T, x_out, x_mean);
    } else {
        dim3 grid(ceil_div(T, 256), B);
        sde_update_kernel<MODE, 1, NC><<<grid, 256, 0, s>>>(*p, x, score, mix, t, sigma_mix, noise, seed, offset,
                                                            coef, flag, sig_ch, T, x_out, x_mean);
    }
    return check_launch("sde_update_kernel");
}

// ------------------------------------------------------------------ training-side forward pieces
// DiffSepModel.sample_prior with the default init_hack (pl_model.py:179-188, 243-247) and the arithmetic of
// compute_score_loss around the network call (:418-424), on the marginal of the forward SDE (sdes.py:286-294,
// 315-328 / 472-494, 515-532, 560-562):
//   x_t = (A + e^{-lambda t} Pn) x0 + L(t) z,   L(t) = (s1 A + s2 Pn) [sigma_mix],    loss_b = mean_{c,t} ((L score) + z)^2
// One pass each (the reference: marginal_prob's two matrix builds, two einsums / matmuls, randn_like, MSELoss).
template <int VEC, int NC>
__global__ void __launch_bounds__(256)
sde_perturb_kernel(const dsep_sde_params p, const float* __restrict__ x0, const float* __restrict__ tvec,
                   const float* __restrict__ sigma_mix, const float* __restrict__ noise, uint64_t seed, uint64_t offset,
                   int T, float* __restrict__ x_t, float* __restrict__ z_out) {
    const int b = blockIdx.y;
    const int t0 = (blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    if (t0 >= T) return;
    const float t = tvec[b];
    const SdeScalars sc = sde_scalars(p, t);
    const float decay = expf(-t * p.d_lambda);
    int64_t e[NC];
    Vec<VEC> xv[NC], z[NC], o[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        e[c] = (static_cast<int64_t>(b) * NC + c) * T + t0;
        xv[c] = ldv<VEC>(x0 + e[c]);
        z[c] = noise_at<VEC>(noise, e[c], seed, offset);
    }
    Vec<VEC> sm;
#pragma unroll
    for (int i = 0; i < VEC; ++i) sm.v[i] = 1.0f;
    if (sigma_mix != nullptr) sm = ldv<VEC>(sigma_mix + static_cast<int64_t>(b) * T + t0);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        float a[NC], m[NC], w[NC], l[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) { a[c] = xv[c].v[i]; w[c] = z[c].v[i]; }
        apply_std<NC>(1.0f, decay, a, m);                  // mean = xbar + e^{-lambda t} (x0 - xbar)
        apply_std<NC>(sc.s1, sc.s2, w, l);
#pragma unroll
        for (int c = 0; c < NC; ++c) o[c].v[i] = m[c] + l[c] * sm.v[i];
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        stv<VEC>(x_t + e[c], o[c]);
        if (z_out != nullptr) stv<VEC>(z_out + e[c], z[c]);
    }
}

// loss[b] += sum over this block's samples of ((L score)[c] + z[c])^2 / (NC T)   (loss zeroed by the entry point)
template <int NC>
__global__ void __launch_bounds__(256)
score_loss_kernel(const dsep_sde_params p, const float* __restrict__ score, const float* __restrict__ z,
                  const float* __restrict__ tvec, const float* __restrict__ sigma_mix, int T,
                  double* __restrict__ loss) {
    const int b = blockIdx.y;
    const SdeScalars sc = sde_scalars(p, tvec[b]);
    double acc = 0.0;
    for (int t0 = blockIdx.x * blockDim.x + threadIdx.x; t0 < T; t0 += gridDim.x * blockDim.x) {
        float a[NC], l[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) a[c] = score[(static_cast<int64_t>(b) * NC + c) * T + t0];
        apply_std<NC>(sc.s1, sc.s2, a, l);
        const float sg = sigma_mix != nullptr ? sigma_mix[static_cast<int64_t>(b) * T + t0] : 1.0f;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const float d = l[c] * sg + z[(static_cast<int64_t>(b) * NC + c) * T + t0];
            acc += static_cast<double>(d) * d;
        }
    }
    __shared__ double s_acc[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s_acc[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int w = 0; w < 8; ++w) tot += s_acc[w];
        atomicAdd(loss + b, tot / (static_cast<double>(NC) * T));
    }
}


template <int MODE>
static int launch_update(const dsep_sde_params* p, const float* x, const float* score, const float* mix,
                         const float* t, const float* sigma_mix, const float* noise, uint64_t seed,
                         uint64_t offset, float coef, int flag, int sig_ch, int B, int T, float* x_out, float* x_mean,
                         cudaStream_t s) {
    if (p->ndim == 3)
        return launch_update_nc<MODE, 3>(p, x, score, mix, t, sigma_mix, noise, seed, offset, coef, flag, sig_ch, B, T,
                                         x_out, x_mean, s);
    return launch_update_nc<MODE, 2>(p, x, score, mix, t, sigma_mix, noise, seed, offset, coef, flag, sig_ch, B, T,
                                     x_out, x_mean, s);
}

static int check_sde(const char* who, const dsep_sde_params* p, int B, int T) {
    DSEP_REQUIRE(p != nullptr, "%s: null parameters", who);
    DSEP_REQUIRE(p->sigma_min > 0.f && p->sigma_max > p->sigma_min, "%s: need 0 < sigma_min < sigma_max", who);
    DSEP_REQUIRE(p->ndim == 0 || p->ndim == 2 || p->ndim == 3, "%s: ndim must be 2 or 3 (got %d)", who, p->ndim);
    DSEP_REQUIRE(B > 0 && T > 0 && B <= 65535, "%s: bad shape B=%d T=%d", who, B, T);
    return DSEP_OK;
}

}  // namespace dsep

using namespace dsep;

extern "C" int dsep_sde_prior(const dsep_sde_params* p, const float* mix, int mix_channels, float mean_scale,
                              const float* sigma_mix, int sigma_channels, const float* noise, uint64_t seed,
                              uint64_t offset, int B, int T, float* x, dsep_stream_t stream) {
    int rc = check_sde("sde_prior", p, B, T);
    if (rc) return rc;
    DSEP_REQUIRE(mix && x, "sde_prior: null pointer");
    const int nc = p->ndim == 3 ? 3 : 2;
    DSEP_REQUIRE(mix_channels == 1 || mix_channels == nc,
                 "The input provided to prior_sampling should have 1 channel, or the same as the number of speakers. "
                 "Found %d channels instead.", mix_channels);
    DSEP_REQUIRE(sigma_mix == nullptr || sigma_channels == 1 || sigma_channels == nc, "sde_prior: bad sigma_channels");
    return launch_update<0>(p, nullptr, nullptr, mix, nullptr, sigma_mix, noise, seed, offset, mean_scale, mix_channels,
                            sigma_channels, B, T, x, nullptr, (cudaStream_t)stream);
}

extern "C" int dsep_sde_corrector(const dsep_sde_params* p, const float* x, const float* score,
                                  const float* t, const float* sigma_mix, const float* noise, uint64_t seed,
                                  uint64_t offset, float snr, int B, int T, float* x_out, float* x_mean,
                                  dsep_stream_t stream) {
    int rc = check_sde("sde_corrector", p, B, T);
    if (rc) return rc;
    DSEP_REQUIRE(x && score && t && x_out, "sde_corrector: null pointer");
    return launch_update<1>(p, x, score, nullptr, t, sigma_mix, noise, seed, offset, snr, 0, 1, B, T, x_out, x_mean,
                            (cudaStream_t)stream);
}

extern "C" int dsep_sde_predictor(const dsep_sde_params* p, const float* x, const float* score,
                                  const float* t, const float* sigma_mix, const float* noise, uint64_t seed,
                                  uint64_t offset, float dt, int probability_flow, int B, int T, float* x_out,
                                  float* x_mean, dsep_stream_t stream) {
    int rc = check_sde("sde_predictor", p, B, T);
    if (rc) return rc;
    DSEP_REQUIRE(x && score && t && x_out, "sde_predictor: null pointer");
    DSEP_REQUIRE(dt > 0.f, "sde_predictor: dt must be positive");
    return launch_update<2>(p, x, score, nullptr, t, sigma_mix, noise, seed, offset, dt, probability_flow ? 1 : 0, 1, B,
                            T, x_out, x_mean, (cudaStream_t)stream);
}

extern "C" int dsep_sde_perturb(const dsep_sde_params* p, const float* x0, const float* t, const float* sigma_mix,
                                const float* noise, uint64_t seed, uint64_t offset, int B, int T, float* x_t,
                                float* z_out, dsep_stream_t stream) {
    int rc = check_sde("sde_perturb", p, B, T);
    if (rc) return rc;
    DSEP_REQUIRE(x0 && t && x_t, "sde_perturb: null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    auto aligned = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
    const bool vec = T % 4 == 0 && aligned(x0) && aligned(x_t) && (!sigma_mix || aligned(sigma_mix)) &&
                     (!noise || aligned(noise)) && (!z_out || aligned(z_out));
    const bool three = p->ndim == 3;
    if (vec) {
        dim3 grid(ceil_div(T / 4, 256), B);
        if (three) sde_perturb_kernel<4, 3><<<grid, 256, 0, s>>>(*p, x0, t, sigma_mix, noise, seed, offset, T, x_t, z_out);
        else sde_perturb_kernel<4, 2><<<grid, 256, 0, s>>>(*p, x0, t, sigma_mix, noise, seed, offset, T, x_t, z_out);
    } else {
        dim3 grid(ceil_div(T, 256), B);
        if (three) sde_perturb_kernel<1, 3><<<grid, 256, 0, s>>>(*p, x0, t, sigma_mix, noise, seed, offset, T, x_t, z_out);
        else sde_perturb_kernel<1, 2><<<grid, 256, 0, s>>>(*p, x0, t, sigma_mix, noise, seed, offset, T, x_t, z_out);
    }
    return check_launch("sde_perturb_kernel");
}

extern "C" int dsep_score_loss(const dsep_sde_params* p, const float* score, const float* z, const float* t,
                               const float* sigma_mix, int B, int T, double* loss, dsep_stream_t stream) {
    int rc = check_sde("score_loss", p, B, T);
    if (rc) return rc;
    DSEP_REQUIRE(score && z && t && loss, "score_loss: null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    if (cudaMemsetAsync(loss, 0, sizeof(double) * B, s) != cudaSuccess) return check_launch("score_loss: memset");
    int gx = ceil_div(T, 256 * 8);
    if (gx > 64) gx = 64;
    dim3 grid(gx, B);
    if (p->ndim == 3) score_loss_kernel<3><<<grid, 256, 0, s>>>(*p, score, z, t, sigma_mix, T, loss);
    else score_loss_kernel<2><<<grid, 256, 0, s>>>(*p, score, z, t, sigma_mix, T, loss);
    return check_launch("score_loss_kernel");
}

extern "C" int dsep_sde_corrector_ald(const dsep_sde_params* p, const float* x, const float* score,
                                      const float* t, const float* noise, uint64_t seed, uint64_t offset,
                                      float snr, int B, int T, float* x_out, float* x_mean,
                                      dsep_stream_t stream) {
    int rc = check_sde("sde_corrector_ald", p, B, T);
    if (rc) return rc;
    DSEP_REQUIRE(x && score && t && x_out, "sde_corrector_ald: null pointer");
    return launch_update<3>(p, x, score, nullptr, t, nullptr, noise, seed, offset, snr, 0, 1, B, T, x_out, x_mean,
                            (cudaStream_t)stream);
}

extern "C" int dsep_sde_corrector_langevin(const float* x, const float* score, const float* noise, float snr,
                                           int B, int n, float* norms, float* x_out, float* x_mean,
                                           dsep_stream_t stream) {
    DSEP_REQUIRE(x && score && noise && norms && x_out && x_mean, "sde_corrector_langevin: null pointer");
    DSEP_REQUIRE(B > 0 && B <= 4096 && n > 0, "sde_corrector_langevin: bad shape");
    cudaStream_t s = (cudaStream_t)stream;
    item_norms_kernel<<<dim3(B, 2), 1024, 0, s>>>(score, noise, n, norms);
    const int64_t total = (int64_t)B * n;
    int64_t blocks = (total + 1023) / 1024;
    if (blocks > 148 * 16) blocks = 148 * 16;
    langevin_kernel<<<(int)blocks, 256, 0, s>>>(x, score, noise, norms, snr, B, total, x_out, x_mean);
    return check_launch("langevin_kernel");
}

extern "C" int dsep_sigma_mix(const float* mix, int B, int T, int avg_len, float* sigma, dsep_stream_t stream) {
    DSEP_REQUIRE(mix && sigma, "sigma_mix: null pointer");
    DSEP_REQUIRE(B > 0 && T > 0 && B <= 65535, "sigma_mix: bad shape");
    DSEP_REQUIRE(avg_len > 0 && avg_len <= 4096, "sigma_mix: avg_len out of range");
    dim3 grid(ceil_div(T, kSigTile), B);
    const size_t smem = sizeof(float) * (kSigTile + avg_len);
    sigma_mix_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(mix, T, avg_len, sigma);
    return check_launch("sigma_mix_kernel");
}

extern "C" int dsep_normalize(const float* mix, int B, int n, float* out, float* mean, float* std,
                              dsep_stream_t stream) {
    DSEP_REQUIRE(mix && out, "normalize: null pointer");
    DSEP_REQUIRE(B > 0 && n > 0, "normalize: bad shape");
    normalize_kernel<<<B, 1024, 0, (cudaStream_t)stream>>>(mix, n, out, mean, std);
    return check_launch("normalize_kernel");
}

extern "C" int dsep_scale_output(const float* mix, const float* sep, int B, int nsrc, int T, float* out,
                                 dsep_stream_t stream) {
    DSEP_REQUIRE(mix && sep && out, "scale_output: null pointer");
    DSEP_REQUIRE(B > 0 && nsrc > 0 && T > 0, "scale_output: bad shape");
    scale_output_kernel<<<B * nsrc, 1024, 0, (cudaStream_t)stream>>>(mix, sep, nsrc, T, out);
    return check_launch("scale_output_kernel");
}

extern "C" int dsep_randn(float* z, int64_t n, uint64_t seed, uint64_t offset, dsep_stream_t stream) {
    DSEP_REQUIRE(z && n >= 0, "randn: bad arguments");
    if (n == 0) return DSEP_OK;
    int64_t blocks = (n / 4 + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    randn_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(z, n, seed, offset);
    return check_launch("randn_kernel");
}
