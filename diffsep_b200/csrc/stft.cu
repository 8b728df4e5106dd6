// STFT-510 / iSTFT around the score network, without cuFFT.
//
// n_fft = 510 = 2*3*5*17 is not a power of two (reference config/model/default.yaml:18-22), so
// the transform is a dense real DFT: frames [M, 512] x basis [512, 512] in fp32 FFMA
// (0.4 GFLOP per network evaluation, <2 % of the step).  The surrounding framing, magnitude
// compression, channel packing, 2x-1 input affine, output 1x1 conv, decompression and
// windowed overlap-add are fused into four element-wise kernels.
// Restates ScoreModelNCSNpp.pre_process/post_process (models/score_models.py:107-124).
#include "common.cuh"

namespace dsep {

constexpr int kNfft = 510;
constexpr int kHop = 128;
constexpr int kBins = 256;    // n_fft / 2 + 1
constexpr int kLd = 512;      // padded row length of frame / spectrum matrices

// ---------------------------------------------------------------------------- framing
__global__ void __launch_bounds__(256)
stft_frames_kernel(const float* __restrict__ x, const float* __restrict__ window, int T, int Fr,
                   int64_t rows, float* __restrict__ frames) {
    // one block per frame row; row = (b*C + c)*Fr + f
    for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
        const int f = static_cast<int>(row % Fr);
        const int64_t bc = row / Fr;
        const float* src = x + bc * T;
        const int start = f * kHop - kNfft / 2;
        for (int n = threadIdx.x; n < kLd; n += blockDim.x) {
            float v = 0.f;
            const int t = start + n;
            if (n < kNfft && t >= 0 && t < T) v = window[n] * src[t];
            frames[row * kLd + n] = v;
        }
    }
}

// ---------------------------------------------------------------------------- SGEMM
// C[M,N] = A[M,K] * B[K,N]; 128x128x8 tiles, 256 threads, 8x8 register micro-tiles,
// double-buffered shared memory.  N % 128 == 0, K % 8 == 0; M guarded.
__global__ void __launch_bounds__(256)
sgemm_kernel(const float* __restrict__ A, int lda, const float* __restrict__ Bm, int ldb,
             float* __restrict__ C, int ldc, int M, int N, int K) {
    __shared__ __align__(16) float As[2][8][128 + 4];
    __shared__ __align__(16) float Bs[2][8][128];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * 128, n0 = blockIdx.x * 128;
    const int a_row = tid >> 1, a_k = (tid & 1) * 4;         // A tile: 128 rows x 8 k
    const int b_k = tid >> 5, b_n = (tid & 31) * 4;          // B tile: 8 k x 128 n
    const int ty = tid >> 4, tx = tid & 15;                  // 16 x 16 threads, 8x8 each

    const bool a_ok = (m0 + a_row) < M;
    const float* a_ptr = A + static_cast<size_t>(a_ok ? m0 + a_row : 0) * lda + a_k;
    const float* b_ptr = Bm + static_cast<size_t>(b_k) * ldb + n0 + b_n;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    float4 a_reg = a_ok ? *reinterpret_cast<const float4*>(a_ptr) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 b_reg = *reinterpret_cast<const float4*>(b_ptr);
    As[0][a_k + 0][a_row] = a_reg.x; As[0][a_k + 1][a_row] = a_reg.y;
    As[0][a_k + 2][a_row] = a_reg.z; As[0][a_k + 3][a_row] = a_reg.w;
    *reinterpret_cast<float4*>(&Bs[0][b_k][b_n]) = b_reg;
    __syncthreads();

    const int nk = K / 8;
    for (int kt = 0; kt < nk; ++kt) {
        const int cur = kt & 1;
        if (kt + 1 < nk) {
            a_reg = a_ok ? *reinterpret_cast<const float4*>(a_ptr + (kt + 1) * 8) : make_float4(0.f, 0.f, 0.f, 0.f);
            b_reg = *reinterpret_cast<const float4*>(b_ptr + static_cast<size_t>(kt + 1) * 8 * ldb);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float a[8], b[8];
            *reinterpret_cast<float4*>(&a[0]) = *reinterpret_cast<const float4*>(&As[cur][k][ty * 4]);
            *reinterpret_cast<float4*>(&a[4]) = *reinterpret_cast<const float4*>(&As[cur][k][64 + ty * 4]);
            *reinterpret_cast<float4*>(&b[0]) = *reinterpret_cast<const float4*>(&Bs[cur][k][tx * 4]);
            *reinterpret_cast<float4*>(&b[4]) = *reinterpret_cast<const float4*>(&Bs[cur][k][64 + tx * 4]);
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < nk) {
            const int nxt = cur ^ 1;
            As[nxt][a_k + 0][a_row] = a_reg.x; As[nxt][a_k + 1][a_row] = a_reg.y;
            As[nxt][a_k + 2][a_row] = a_reg.z; As[nxt][a_k + 3][a_row] = a_reg.w;
            *reinterpret_cast<float4*>(&Bs[nxt][b_k][b_n]) = b_reg;
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int row = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (row >= M) continue;
        float* crow = C + static_cast<size_t>(row) * ldc + n0;
        *reinterpret_cast<float4*>(crow + tx * 4) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        *reinterpret_cast<float4*>(crow + 64 + tx * 4) = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
    }
}

// ---------------------------------------------------------------------------- spectrum -> network input
// thread per (b, bin k, frame f in [0, Wp)); writes all Cw channels of this call.
__global__ void __launch_bounds__(256)
spec_pack_kernel(const float* __restrict__ dft, int B, int Cw, int Fr, int Wp, int chan0, int Ctot,
                 int Cpad, float factor, float exponent, float* __restrict__ x_f32,
                 __half* __restrict__ a_hi, __half* __restrict__ a_lo) {
    const int64_t total = (int64_t)B * kBins * Wp;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int f = static_cast<int>(e % Wp);
        const int k = static_cast<int>((e / Wp) % kBins);
        const int b = static_cast<int>(e / ((int64_t)Wp * kBins));
        for (int c = 0; c < Cw; ++c) {
            float re = 0.f, im = 0.f;
            if (f < Fr) {
                const float2 z = *reinterpret_cast<const float2*>(
                    dft + ((static_cast<int64_t>(b) * Cw + c) * Fr + f) * kLd + 2 * k);
                const float mag = sqrtf(z.x * z.x + z.y * z.y);
                float s = 0.f;
                if (mag > 0.f) {
                    if (exponent == 0.5f) s = factor / sqrtf(mag);
                    else if (exponent == 1.0f) s = factor;
                    else s = factor * powf(mag, exponent - 1.0f);
                }
                re = z.x * s;
                im = z.y * s;
            }
            re = 2.0f * re - 1.0f;   // NCSNpp.forward: x = 2x - 1 (centered=False), padded frames -> -1
            im = 2.0f * im - 1.0f;
            const int64_t pix = e;   // (b*256 + k)*Wp + f
            const int cr = chan0 + c, ci = Ctot + chan0 + c;
            if (x_f32 != nullptr) {
                x_f32[pix * (2 * Ctot) + cr] = re;
                x_f32[pix * (2 * Ctot) + ci] = im;
            }
            if (a_hi != nullptr) {
                __half h, l;
                split_f16(re, h, l);
                a_hi[pix * Cpad + cr] = h; a_lo[pix * Cpad + cr] = l;
                split_f16(im, h, l);
                a_hi[pix * Cpad + ci] = h; a_lo[pix * Cpad + ci] = l;
            }
        }
    }
}

// ---------------------------------------------------------------------------- network output -> spectrum
// thread per (b, bin k, frame f < Fr): /t, 1x1 conv Cp -> 2*nsrc, complex pack, decompress.
__global__ void __launch_bounds__(256)
out_head_kernel(const float* __restrict__ pyr, int B, int Wp, int Cp, int nsrc, int Fr,
                const float* __restrict__ t, const float* __restrict__ w, const float* __restrict__ bias,
                float factor, float exponent, float* __restrict__ spec) {
    const int64_t total = (int64_t)B * kBins * Fr;
    const float inv_e = 1.0f / exponent;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int f = static_cast<int>(e % Fr);
        const int k = static_cast<int>((e / Fr) % kBins);
        const int b = static_cast<int>(e / ((int64_t)Fr * kBins));
        const float tb = t[b];
        float hvec[16];
        const float* src = pyr + ((static_cast<int64_t>(b) * kBins + k) * Wp + f) * Cp;
        for (int c = 0; c < Cp; ++c) hvec[c] = src[c] / tb;
        for (int s = 0; s < nsrc; ++s) {
            float re = bias ? bias[s] : 0.f, im = bias ? bias[nsrc + s] : 0.f;
            for (int c = 0; c < Cp; ++c) {
                re = fmaf(w[s * Cp + c], hvec[c], re);
                im = fmaf(w[(nsrc + s) * Cp + c], hvec[c], im);
            }
            re /= factor;
            im /= factor;
            if (exponent != 1.0f) {
                const float mag = sqrtf(re * re + im * im);
                const float g = exponent == 0.5f ? mag : (mag > 0.f ? powf(mag, inv_e - 1.0f) : 0.f);
                re *= g;
                im *= g;
            }
            float2* dst = reinterpret_cast<float2*>(
                spec + ((static_cast<int64_t>(b) * nsrc + s) * Fr + f) * kLd + 2 * k);
            *dst = make_float2(re, im);
        }
    }
}

// ---------------------------------------------------------------------------- overlap-add
__global__ void __launch_bounds__(256)
istft_ola_kernel(const float* __restrict__ frames_t, const float* __restrict__ window, int Fr, int T,
                 int64_t total, float* __restrict__ out) {
    const int len = kHop * (Fr - 1);   // torch.istft output length with center=True
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int t = static_cast<int>(e % T);
        const int64_t bc = e / T;
        float y = 0.f;
        if (t < len) {
            const int u = t + kNfft / 2;
            int f_hi = u / kHop;
            if (f_hi > Fr - 1) f_hi = Fr - 1;
            int f_lo = (u - (kNfft - 1) + kHop - 1) / kHop;
            if (f_lo < 0) f_lo = 0;
            float num = 0.f, den = 0.f;
            for (int f = f_lo; f <= f_hi; ++f) {
                const int n = u - f * kHop;
                const float wn = window[n];
                num = fmaf(wn, frames_t[(bc * Fr + f) * kLd + n], num);
                den = fmaf(wn, wn, den);
            }
            y = den > 1e-11f ? num / den : 0.f;
        }
        out[e] = y;
    }
}

static int grid1d(int64_t items, int max_blocks = 148 * 16) {
    int64_t g = (items + 255) / 256;
    if (g > max_blocks) g = max_blocks;
    if (g < 1) g = 1;
    return static_cast<int>(g);
}

}  // namespace dsep

using namespace dsep;

extern "C" int dsep_stft_frames(const float* x, const float* window, int B, int C, int T, int Fr,
                                float* frames, dsep_stream_t stream) {
    DSEP_REQUIRE(x && window && frames, "stft_frames: null pointer");
    DSEP_REQUIRE(B > 0 && C > 0 && T > 0, "stft_frames: empty signal");
    DSEP_REQUIRE(Fr == 1 + (T + (kNfft - kHop)) / kHop, "stft_frames: Fr must be 1 + (T + 382) / 128 (got %d for T=%d)", Fr, T);
    const int64_t rows = (int64_t)B * C * Fr;
    const int grid = rows < 148 * 32 ? (int)rows : 148 * 32;
    stft_frames_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, window, T, Fr, rows, frames);
    return check_launch("stft_frames_kernel");
}

extern "C" int dsep_sgemm(const float* A, int lda, const float* Bm, int ldb, float* C, int ldc, int M,
                          int N, int K, dsep_stream_t stream) {
    DSEP_REQUIRE(A && Bm && C, "sgemm: null pointer");
    DSEP_REQUIRE(M > 0 && N > 0 && K > 0, "sgemm: empty problem");
    DSEP_REQUIRE(N % 128 == 0 && K % 8 == 0, "sgemm: needs N %% 128 == 0 and K %% 8 == 0 (got N=%d K=%d)", N, K);
    DSEP_REQUIRE(lda % 4 == 0 && ldb % 4 == 0 && ldc % 4 == 0 && lda >= K && ldb >= N && ldc >= N,
                 "sgemm: bad leading dimensions");
    dim3 grid(N / 128, ceil_div(M, 128));
    sgemm_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(A, lda, Bm, ldb, C, ldc, M, N, K);
    return check_launch("sgemm_kernel");
}

extern "C" int dsep_spec_pack(const float* dft, int B, int Cw, int Fr, int Wp, int chan0, int Ctot,
                              int Cpad, float factor, float exponent, float* x_f32, void* a_hi,
                              void* a_lo, dsep_stream_t stream) {
    DSEP_REQUIRE(dft && (x_f32 || a_hi), "spec_pack: null pointer");
    DSEP_REQUIRE((a_hi == nullptr) == (a_lo == nullptr), "spec_pack: hi/lo planes must come in pairs");
    DSEP_REQUIRE(B > 0 && Cw > 0 && Fr > 0 && Wp >= Fr, "spec_pack: bad shape");
    DSEP_REQUIRE(chan0 >= 0 && chan0 + Cw <= Ctot && Cpad >= 2 * Ctot, "spec_pack: bad channel placement");
    DSEP_REQUIRE(exponent > 0.f && factor != 0.f, "spec_pack: bad transform parameters");
    const int64_t total = (int64_t)B * kBins * Wp;
    spec_pack_kernel<<<grid1d(total), 256, 0, (cudaStream_t)stream>>>(
        dft, B, Cw, Fr, Wp, chan0, Ctot, Cpad, factor, exponent, x_f32, (__half*)a_hi,
        (__half*)a_lo);
    return check_launch("spec_pack_kernel");
}

extern "C" int dsep_out_head(const float* pyr, int B, int Wp, int Cp, int nsrc, int Fr, const float* t,
                             const float* w, const float* bias, float factor, float exponent,
                             float* spec, dsep_stream_t stream) {
    DSEP_REQUIRE(pyr && t && w && spec, "out_head: null pointer");
    DSEP_REQUIRE(B > 0 && Fr > 0 && Wp >= Fr && Cp > 0 && Cp <= 16 && nsrc > 0, "out_head: bad shape");
    DSEP_REQUIRE(exponent > 0.f && factor != 0.f, "out_head: bad transform parameters");
    const int64_t total = (int64_t)B * kBins * Fr;
    out_head_kernel<<<grid1d(total), 256, 0, (cudaStream_t)stream>>>(pyr, B, Wp, Cp, nsrc, Fr, t, w, bias,
                                                                      fabsf(factor), exponent, spec);
    return check_launch("out_head_kernel");
}

extern "C" int dsep_istft_ola(const float* frames_t, const float* window, int B, int C, int Fr, int T,
                              float* out, dsep_stream_t stream) {
    DSEP_REQUIRE(frames_t && window && out, "istft_ola: null pointer");
    DSEP_REQUIRE(B > 0 && C > 0 && Fr > 0 && T > 0, "istft_ola: bad shape");
    const int64_t total = (int64_t)B * C * T;
    istft_ola_kernel<<<grid1d(total), 256, 0, (cudaStream_t)stream>>>(frames_t, window, Fr, T, total, out);
    return check_launch("istft_ola_kernel");
}
