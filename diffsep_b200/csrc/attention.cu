// Single-head self-attention over the 16 x W/16 map of the NCSN++ bottleneck and the time
// embedding MLP.  Attention is 0.04 % of the network's FLOPs, so this is a
// plain fp32 kernel: scores for a block of queries live in shared memory (no S x S matrix in
// HBM, unlike the reference's materialised einsum, layerspp.py:83-87), softmax in fp32,
// output written directly as (hi, lo) fp16 planes for the NIN_3 tensor-core projection.
#include <stdlib.h>

#include "common.cuh"

namespace dsep {

constexpr int kQT = 8;   // queries per block

// grid (ceil(S/kQT), B), 256 threads, dynamic smem: q[kQT][C] + scores[kQT][S]
__global__ void __launch_bounds__(256)
attention_kernel(const float* __restrict__ qkv, int S, int C, float scale,
                 __half* __restrict__ o_hi, __half* __restrict__ o_lo) {
    extern __shared__ __align__(16) float sm[];
    float* s_q = sm;                 // [kQT][C]
    float* s_p = sm + kQT * C;       // [kQT][S]
    __shared__ float s_inv[kQT];
    const int b = blockIdx.y;
    const int q0 = blockIdx.x * kQT;
    const int nq = min(kQT, S - q0);
    const float* base = qkv + static_cast<size_t>(b) * S * 3 * C;

    for (int i = threadIdx.x; i < kQT * C; i += blockDim.x) {
        const int qi = i / C, c = i % C;
        s_q[i] = qi < nq ? base[static_cast<size_t>(q0 + qi) * 3 * C + c] : 0.f;
    }
    __syncthreads();

    // scores: thread <-> key
    for (int j = threadIdx.x; j < S; j += blockDim.x) {
        const float4* krow = reinterpret_cast<const float4*>(base + static_cast<size_t>(j) * 3 * C + C);
        float acc[kQT];
#pragma unroll
        for (int qi = 0; qi < kQT; ++qi) acc[qi] = 0.f;
        for (int c4 = 0; c4 < C / 4; ++c4) {
            const float4 kv = krow[c4];
#pragma unroll
            for (int qi = 0; qi < kQT; ++qi) {
                const float4 qv = *reinterpret_cast<const float4*>(s_q + qi * C + c4 * 4);
                acc[qi] = fmaf(qv.x, kv.x, acc[qi]);
                acc[qi] = fmaf(qv.y, kv.y, acc[qi]);
                acc[qi] = fmaf(qv.z, kv.z, acc[qi]);
                acc[qi] = fmaf(qv.w, kv.w, acc[qi]);
            }
        }
#pragma unroll
        for (int qi = 0; qi < kQT; ++qi) s_p[qi * S + j] = acc[qi] * scale;
    }
    __syncthreads();

    // softmax: warp <-> query
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp < kQT) {
        float* row = s_p + warp * S;
        float mx = -INFINITY;
        for (int j = lane; j < S; j += 32) mx = fmaxf(mx, row[j]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float sum = 0.f;
        for (int j = lane; j < S; j += 32) {
            const float e = expf(row[j] - mx);
            row[j] = e;
            sum += e;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (lane == 0) s_inv[warp] = 1.0f / sum;
    }
    __syncthreads();

    // output: thread <-> channel
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float acc[kQT];
#pragma unroll
        for (int qi = 0; qi < kQT; ++qi) acc[qi] = 0.f;
        const float* vcol = base + 2 * C + c;
        for (int j = 0; j < S; ++j) {
            const float v = vcol[static_cast<size_t>(j) * 3 * C];
#pragma unroll
            for (int qi = 0; qi < kQT; ++qi) acc[qi] = fmaf(s_p[qi * S + j], v, acc[qi]);
        }
        for (int qi = 0; qi < nq; ++qi) {
            const float o = acc[qi] * s_inv[qi];
            __half h, l;
            split_f16(o, h, l);
            const size_t off = (static_cast<size_t>(b) * S + q0 + qi) * C + c;
            o_hi[off] = h;
            o_lo[off] = l;
        }
    }
}

// ------------------------------------------------------------------------ time embedding
// one block per batch entry; D = 4 nf.
__global__ void __launch_bounds__(256)
time_embedding_kernel(const float* __restrict__ t, const float* __restrict__ Wf,
                      const float* __restrict__ w1, const float* __restrict__ b1,
                      const float* __restrict__ w2, const float* __restrict__ b2, int nf,
                      float* __restrict__ temb_act) {
    extern __shared__ float sm[];
    float* s_emb = sm;              // [2 nf]
    float* s_h = sm + 2 * nf;       // [4 nf]
    const int b = blockIdx.x, D = 4 * nf, E = 2 * nf;
    // x_proj = ((log t * W) * 2) * pi in fp32, the reference's evaluation order (layerspp.py:40)
    const float lt = static_cast<float>(log(static_cast<double>(t[b])));
    const float pi_f = 3.14159265358979323846f;
    for (int i = threadIdx.x; i < nf; i += blockDim.x) {
        float xp = lt * Wf[i];
        xp = xp * 2.0f;
        xp = xp * pi_f;
        // sin/cos of a ~1e3 rad phase: evaluate in double so that only the fp32 phase matters
        s_emb[i] = static_cast<float>(sin(static_cast<double>(xp)));
        s_emb[nf + i] = static_cast<float>(cos(static_cast<double>(xp)));
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int r = warp; r < D; r += nw) {
        float acc = 0.f;
        for (int k = lane; k < E; k += 32) acc = fmaf(w1[static_cast<size_t>(r) * E + k], s_emb[k], acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) s_h[r] = silu_f(acc + b1[r]);
    }
    __syncthreads();
    for (int r = warp; r < D; r += nw) {
        float acc = 0.f;
        for (int k = lane; k < D; k += 32) acc = fmaf(w2[static_cast<size_t>(r) * D + k], s_h[k], acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) temb_act[static_cast<size_t>(b) * D + r] = silu_f(acc + b2[r]);
    }
}

// film[b, r] = temb_act[b,:] . Wd[r,:] + bd[r]; warp per output row, all batch entries at once
// in chunks of 8 so the weight row is read once per chunk.
__global__ void __launch_bounds__(256)
film_kernel(const float* __restrict__ temb_act, const float* __restrict__ Wd,
            const float* __restrict__ bd, int B, int D, int R, float* __restrict__ film) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= R) return;
    const float* wrow = Wd + static_cast<size_t>(warp) * D;
    for (int b0 = 0; b0 < B; b0 += 8) {
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.f;
        for (int k = lane; k < D; k += 32) {
            const float w = wrow[k];
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (b0 + i < B) acc[i] = fmaf(w, temb_act[static_cast<size_t>(b0 + i) * D + k], acc[i]);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
        }
        if (lane == 0)
            for (int i = 0; i < 8 && b0 + i < B; ++i)
                film[static_cast<size_t>(b0 + i) * R + warp] = acc[i] + bd[warp];
    }
}

}  // namespace dsep

using namespace dsep;

// attention_tc.cu: the tcgen05 flash-style kernel (C a multiple of 64, at most 256)
int dsep_attention_tc_launch(const float* qkv, int B, int S, int C, float scale, void* o_hi, void* o_lo,
                             cudaStream_t stream);

extern "C" int dsep_attention(const float* qkv, int B, int S, int C, float scale, void* o_hi, void* o_lo,
                              dsep_stream_t stream) {
    DSEP_REQUIRE(qkv && o_hi && o_lo, "attention: null pointer");
    DSEP_REQUIRE(B > 0 && S > 0 && C > 0 && C % 4 == 0 && B <= 65535, "attention: bad shape");
    // DSEP_ATTN_TC=0: the fp32 CUDA-core kernel below (kept for A/B checks and for channel counts the tensor-core
    // kernel does not tile)
    static const int tc_env = getenv("DSEP_ATTN_TC") ? atoi(getenv("DSEP_ATTN_TC")) : 1;
    if (tc_env != 0 && C % 64 == 0 && C <= 256)
        return dsep_attention_tc_launch(qkv, B, S, C, scale, o_hi, o_lo, (cudaStream_t)stream);
    const size_t smem = sizeof(float) * (static_cast<size_t>(kQT) * C + static_cast<size_t>(kQT) * S);
    DSEP_REQUIRE(smem <= 200 * 1024, "attention: S=%d too long for the single-pass score buffer", S);
    if (smem > 48 * 1024) {
        static PerDeviceAttr attr;
        const cudaError_t e = set_max_smem_once(attr, attention_kernel, 200 * 1024);
        if (e != cudaSuccess) {
            set_error("attention: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return DSEP_ERR_CUDA;
        }
    }
    dim3 grid(ceil_div(S, kQT), B);
    attention_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(qkv, S, C, scale, (__half*)o_hi,
                                                                (__half*)o_lo);
    return check_launch("attention_kernel");
}

extern "C" int dsep_time_embedding(const float* t, const float* Wf, const float* w1, const float* b1,
                                   const float* w2, const float* b2, int B, int nf, float* temb_act,
                                   dsep_stream_t stream) {
    DSEP_REQUIRE(t && Wf && w1 && b1 && w2 && b2 && temb_act, "time_embedding: null pointer");
    DSEP_REQUIRE(B > 0 && nf > 0 && nf <= 1024, "time_embedding: bad shape");
    const size_t smem = sizeof(float) * 6 * nf;
    time_embedding_kernel<<<B, 256, smem, (cudaStream_t)stream>>>(t, Wf, w1, b1, w2, b2, nf, temb_act);
    return check_launch("time_embedding_kernel");
}

extern "C" int dsep_film(const float* temb_act, const float* Wd, const float* bd, int B, int D, int R,
                         float* film, dsep_stream_t stream) {
    DSEP_REQUIRE(temb_act && Wd && bd && film, "film: null pointer");
    DSEP_REQUIRE(B > 0 && D > 0 && R > 0, "film: bad shape");
    const int blocks = ceil_div(R * 32, 256);
    film_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(temb_act, Wd, bd, B, D, R, film);
    return check_launch("film_kernel");
}
