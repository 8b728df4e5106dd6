// Implicit-GEMM 3x3 / 1x1 convolution on tcgen05 tensor cores (sm_100a).
//
//   out[b,h,w,n] = scale * ( acc_scale * sum_{tap,c} A[b,h+dy,w+dx,c] * Wt[tap,n,c] + bias[n]
//                            + film[b,n] + residual[b,h,w,n] )
//
// Replaces the cuDNN convolutions behind ddpm_conv3x3 / ddpm_conv1x1 / NIN of the reference
// (models/ncsnpp_utils/layers.py:112-156, 678-689) together with the Dense_0 bias add and the
// (x + h)/sqrt(2) residual of ResnetBlockBigGANpp.forward (layerspp.py:311-323).
//
// Design (B200-first):
//   * activations are channels-last, so an output tile of 128 pixels x 64 input channels of one
//     filter tap is ONE 4-D TMA box [64 ch, tw, th, tb] at coordinates shifted by (dx, dy);
//     out-of-image taps are zero-filled by TMA — no im2col, no padding copies;
//   * operands are (hi, lo) fp16 planes (weights pre-scaled by a power of two, undone by
//     acc_scale); three tcgen05.mma passes hi*hi + lo*hi + hi*lo accumulate in fp32 in TMEM,
//     which reproduces the fp32 convolution to ~1e-6 (passes = 1 gives TF32-grade 11-bit
//     operands, what cuDNN runs for the reference by default on a GPU);
//   * persistent CTAs (one per SM), warp-specialised: warp 0 TMA producer, warp 1 MMA issuer,
//     warp 2 TMEM allocator, warps 4-7 epilogue; smem ring of NSTAGES K-blocks, TMEM
//     accumulator double-buffered so the epilogue of tile i overlaps the MMAs of tile i+1;
//   * epilogue: tcgen05.ld -> per-warp smem transpose -> bias/FiLM/residual/scale fused ->
//     128-byte coalesced fp32 stores.
#include <mutex>

#include "common.cuh"

namespace dsep {

struct ConvParams {
    int B, H, W, Cin, Cout_pad, cout_store;
    int taps;              // 1 or 9
    int tw_log2, th_log2;  // pixel tile: tw x th x tb = 128
    int tiles_w, tiles_h, tiles_b, tiles_n, total_tiles;
    int kblocks;           // Cin / 64
    int passes;            // 1 or 3
    const float* bias;
    const float* film;
    int film_stride;
    const float* residual;
    float scale, acc_scale;
    float* out;
};

template <int NT>
struct ConvCfg {
    static constexpr int kStageBytes = 2 * 16384 + 2 * NT * 128;
    static constexpr int kStages = NT >= 128 ? 3 : (NT >= 64 ? 4 : 5);
    static constexpr int kStagingBytes = NT >= 32 ? 4 * 32 * 36 * 4 : 0;
    static constexpr int kTmemCols = 2 * NT < 32 ? 32 : 2 * NT;
    static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + 256 + 1024;
};

template <int NT>
__global__ void __launch_bounds__(256, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
               const __grid_constant__ CUtensorMap tm_w_hi, const __grid_constant__ CUtensorMap tm_w_lo,
               const ConvParams p) {
    using Cfg = ConvCfg<NT>;
    constexpr int NS = Cfg::kStages;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* stage_base = smem;
    float* staging = reinterpret_cast<float*>(smem + NS * Cfg::kStageBytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NS * Cfg::kStageBytes + Cfg::kStagingBytes);
    uint64_t* full = bars;             // [NS]   TMA -> MMA
    uint64_t* empty = bars + NS;       // [NS]   MMA -> TMA
    uint64_t* tfull = bars + 2 * NS;   // [2]    MMA -> epilogue
    uint64_t* tempty = tfull + 2;      // [2]    epilogue -> MMA
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_a_hi);
        tma_prefetch_desc(&tm_w_hi);
        if (p.passes == 3) {
            tma_prefetch_desc(&tm_a_lo);
            tma_prefetch_desc(&tm_w_lo);
        }
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < NS; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull[i], 1);
            mbar_init(&tempty[i], 4);
        }
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc<Cfg::kTmemCols>(tmem_ptr);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    const int tb_log2 = 7 - p.tw_log2 - p.th_log2;
    const int kiters = p.taps * p.kblocks;
    const uint32_t stage_tx = (p.passes == 3 ? 2u : 1u) * (16384u + NT * 128u);

    if (warp == 0 && lane == 0) {
        // ------------------------------------------------------------------ TMA producer
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
            int r = tile;
            const int nt = r % p.tiles_n; r /= p.tiles_n;
            const int wt = r % p.tiles_w; r /= p.tiles_w;
            const int ht = r % p.tiles_h; r /= p.tiles_h;
            const int w0 = wt << p.tw_log2, h0 = ht << p.th_log2, b0 = r << tb_log2;
            const int n0 = nt * NT;
            for (int tap = 0; tap < p.taps; ++tap) {
                const int dy = p.taps == 9 ? tap / 3 - 1 : 0;
                const int dx = p.taps == 9 ? tap % 3 - 1 : 0;
                for (int kb = 0; kb < p.kblocks; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1u);
                    uint8_t* s = stage_base + stage * Cfg::kStageBytes;
                    mbar_arrive_expect_tx(&full[stage], stage_tx);
                    tma_load_4d(s, &tm_a_hi, &full[stage], kb * 64, w0 + dx, h0 + dy, b0);
                    tma_load_2d(s + 32768, &tm_w_hi, &full[stage], kb * 64, tap * p.Cout_pad + n0);
                    if (p.passes == 3) {
                        tma_load_4d(s + 16384, &tm_a_lo, &full[stage], kb * 64, w0 + dx, h0 + dy, b0);
                        tma_load_2d(s + 32768 + NT * 128, &tm_w_lo, &full[stage], kb * 64,
                                    tap * p.Cout_pad + n0);
                    }
                    if (++stage == NS) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1 && lane == 0) {
        // ------------------------------------------------------------------ MMA issuer
        constexpr uint32_t idesc = umma_idesc_f16(128, NT);
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
            const int as = it & 1;
            mbar_wait(&tempty[as], ((it >> 1) & 1) ^ 1u);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + as * NT;
            for (int ki = 0; ki < kiters; ++ki) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                const uint32_t s = smem_u32(stage_base + stage * Cfg::kStageBytes);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint64_t a_hi = umma_desc_sw128(s + k * 32);
                    const uint64_t b_hi = umma_desc_sw128(s + 32768 + k * 32);
                    umma_f16(d_tmem, a_hi, b_hi, idesc, (ki | k) != 0);
                    if (p.passes == 3) {
                        const uint64_t a_lo = umma_desc_sw128(s + 16384 + k * 32);
                        const uint64_t b_lo = umma_desc_sw128(s + 32768 + NT * 128 + k * 32);
                        umma_f16(d_tmem, a_lo, b_hi, idesc, 1);
                        umma_f16(d_tmem, a_hi, b_lo, idesc, 1);
                    }
                }
                umma_commit(&empty[stage]);   // frees the smem slot once these MMAs retire
                if (++stage == NS) { stage = 0; phase ^= 1u; }
            }
            umma_commit(&tfull[as]);          // accumulator complete -> epilogue
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------------ epilogue
        const int wq = warp - 4;              // TMEM lane quarter == warp_id % 4
        const int tw_mask = (1 << p.tw_log2) - 1, th_mask = (1 << p.th_log2) - 1;
        int it = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
            const int as = it & 1;
            int r = tile;
            const int nt = r % p.tiles_n; r /= p.tiles_n;
            const int wt = r % p.tiles_w; r /= p.tiles_w;
            const int ht = r % p.tiles_h; r /= p.tiles_h;
            const int w0 = wt << p.tw_log2, h0 = ht << p.th_log2, b0 = r << tb_log2;
            const int n0 = nt * NT;
            mbar_wait(&tfull[as], (it >> 1) & 1);
            tc_fence_after();
            const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + as * NT;

            if constexpr (NT >= 32) {
                float* stg = staging + wq * (32 * 36);
#pragma unroll 1
                for (int c = 0; c < NT / 32; ++c) {
                    uint32_t v[32];
                    tmem_ld_32x32(t_addr + c * 32, v);
                    tmem_ld_wait();
                    if (c == NT / 32 - 1) {   // TMEM fully drained: hand the buffer back early
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tempty[as]);
                    }
                    float4* dst = reinterpret_cast<float4*>(stg + lane * 36);
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        dst[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                             __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
                    __syncwarp();
                    const int q = lane & 7;
                    const int n = n0 + c * 32 + q * 4;
                    float4 bz = make_float4(0.f, 0.f, 0.f, 0.f);
                    const bool n_ok = n < p.cout_store;
                    if (p.bias != nullptr && n_ok) bz = *reinterpret_cast<const float4*>(p.bias + n);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int row = i * 4 + (lane >> 3);
                        const int m = wq * 32 + row;
                        const int w = w0 + (m & tw_mask);
                        const int h = h0 + ((m >> p.tw_log2) & th_mask);
                        const int b = b0 + (m >> (p.tw_log2 + p.th_log2));
                        if (n_ok && b < p.B && h < p.H && w < p.W) {
                            float4 a = *reinterpret_cast<const float4*>(stg + row * 36 + q * 4);
                            a.x = fmaf(a.x, p.acc_scale, bz.x); a.y = fmaf(a.y, p.acc_scale, bz.y);
                            a.z = fmaf(a.z, p.acc_scale, bz.z); a.w = fmaf(a.w, p.acc_scale, bz.w);
                            if (p.film != nullptr) {
                                const float4 f = *reinterpret_cast<const float4*>(
                                    p.film + static_cast<size_t>(b) * p.film_stride + n);
                                a.x += f.x; a.y += f.y; a.z += f.z; a.w += f.w;
                            }
                            const size_t off =
                                ((static_cast<size_t>(b) * p.H + h) * p.W + w) * p.cout_store + n;
                            if (p.residual != nullptr) {
                                const float4 rr = *reinterpret_cast<const float4*>(p.residual + off);
                                a.x += rr.x; a.y += rr.y; a.z += rr.z; a.w += rr.w;
                            }
                            a.x *= p.scale; a.y *= p.scale; a.z *= p.scale; a.w *= p.scale;
                            *reinterpret_cast<float4*>(p.out + off) = a;
                        }
                    }
                    __syncwarp();
                }
            } else {
                // narrow output (pyramid convs, Cout = 6 padded to 16): thread == pixel
                uint32_t v[16];
                tmem_ld_32x16(t_addr, v);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[as]);
                const int m = wq * 32 + lane;
                const int w = w0 + (m & tw_mask);
                const int h = h0 + ((m >> p.tw_log2) & th_mask);
                const int b = b0 + (m >> (p.tw_log2 + p.th_log2));
                if (b < p.B && h < p.H && w < p.W) {
                    const size_t off = ((static_cast<size_t>(b) * p.H + h) * p.W + w) * p.cout_store;
#pragma unroll
                    for (int n = 0; n < 16; ++n) {
                        if (n < p.cout_store) {
                            float a = __uint_as_float(v[n]) * p.acc_scale;
                            if (p.bias != nullptr) a += p.bias[n];
                            if (p.film != nullptr) a += p.film[static_cast<size_t>(b) * p.film_stride + n];
                            if (p.residual != nullptr) a += p.residual[off + n];
                            p.out[off + n] = a * p.scale;
                        }
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc<Cfg::kTmemCols>(tmem_base);
}

// ------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    });
    return fn;
}

static int make_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims,
                    const cuuint64_t* strides_bytes, const cuuint32_t* box) {
    EncodeTiledFn enc = get_encode_tiled();
    if (enc == nullptr) {
        set_error("cuTensorMapEncodeTiled not available from the CUDA driver");
        return DSEP_ERR_CUDA;
    }
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, const_cast<void*>(base), dims,
                     strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
        return DSEP_ERR_CUDA;
    }
    return DSEP_OK;
}

static int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

template <int NT>
static int launch_conv(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& w_hi,
                       const CUtensorMap& w_lo, const ConvParams& p, cudaStream_t stream) {
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [] {
        attr_err = cudaFuncSetAttribute(conv_tc_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        ConvCfg<NT>::kSmemBytes);
    });
    if (attr_err != cudaSuccess) {
        set_error("cudaFuncSetAttribute(conv_tc_kernel<%d>): %s", NT, cudaGetErrorString(attr_err));
        return DSEP_ERR_CUDA;
    }
    const int grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
    conv_tc_kernel<NT><<<grid, 256, ConvCfg<NT>::kSmemBytes, stream>>>(a_hi, a_lo, w_hi, w_lo, p);
    return check_launch("conv_tc_kernel");
}

static int ilog2(int v) {
    int l = 0;
    while ((1 << l) < v) ++l;
    return l;
}

}  // namespace dsep

extern "C" int dsep_conv2d_tc(const void* a_hi, const void* a_lo, int B, int H, int W, int Cin,
                              const void* w_hi, const void* w_lo, int Cout_pad, int ksize,
                              const float* bias, const float* film, int film_stride,
                              const float* residual, float scale, float acc_scale, float* out,
                              int cout_store, int passes, dsep_stream_t stream) {
    using namespace dsep;
    DSEP_REQUIRE(a_hi && w_hi && out, "conv2d_tc: null operand");
    DSEP_REQUIRE(passes == 1 || passes == 3, "conv2d_tc: passes must be 1 or 3 (got %d)", passes);
    DSEP_REQUIRE(passes == 1 || (a_lo && w_lo), "conv2d_tc: passes=3 needs the lo planes");
    DSEP_REQUIRE(ksize == 1 || ksize == 3, "conv2d_tc: ksize must be 1 or 3 (got %d)", ksize);
    DSEP_REQUIRE(B > 0 && H > 0 && W > 0, "conv2d_tc: empty tensor");
    DSEP_REQUIRE(Cin > 0 && Cin % 64 == 0, "conv2d_tc: Cin must be a multiple of 64 (got %d)", Cin);
    DSEP_REQUIRE(Cout_pad == 16 || (Cout_pad > 0 && Cout_pad % 64 == 0),
                 "conv2d_tc: Cout_pad must be 16 or a multiple of 64 (got %d)", Cout_pad);
    DSEP_REQUIRE(cout_store > 0 && cout_store <= Cout_pad, "conv2d_tc: bad cout_store %d", cout_store);
    DSEP_REQUIRE(Cout_pad == 16 || cout_store % 4 == 0, "conv2d_tc: cout_store must be a multiple of 4");
    const int NT = Cout_pad == 16 ? 16 : (Cout_pad % 128 == 0 ? 128 : 64);

    ConvParams p{};
    p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.Cout_pad = Cout_pad; p.cout_store = cout_store;
    p.taps = ksize * ksize;
    int tw = 1 << ilog2(W); if (tw > 16) tw = 16;
    int th = 1 << ilog2(H); if (th > 128 / tw) th = 128 / tw;
    const int tb = 128 / (tw * th);
    p.tw_log2 = ilog2(tw); p.th_log2 = ilog2(th);
    p.tiles_w = ceil_div(W, tw); p.tiles_h = ceil_div(H, th); p.tiles_b = ceil_div(B, tb);
    p.tiles_n = Cout_pad / NT;
    p.total_tiles = p.tiles_w * p.tiles_h * p.tiles_b * p.tiles_n;
    p.kblocks = Cin / 64;
    p.passes = passes;
    p.bias = bias; p.film = film; p.film_stride = film_stride; p.residual = residual;
    p.scale = scale; p.acc_scale = acc_scale; p.out = out;

    CUtensorMap ma_hi, ma_lo, mw_hi, mw_lo;
    const cuuint64_t adims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    const cuuint64_t astr[3] = {(cuuint64_t)Cin * 2, (cuuint64_t)W * Cin * 2, (cuuint64_t)H * W * Cin * 2};
    const cuuint32_t abox[4] = {64, (cuuint32_t)tw, (cuuint32_t)th, (cuuint32_t)tb};
    const cuuint64_t wdims[2] = {(cuuint64_t)Cin, (cuuint64_t)p.taps * Cout_pad};
    const cuuint64_t wstr[1] = {(cuuint64_t)Cin * 2};
    const cuuint32_t wbox[2] = {64, (cuuint32_t)NT};
    int rc;
    if ((rc = make_map(&ma_hi, a_hi, 4, adims, astr, abox)) != DSEP_OK) return rc;
    if ((rc = make_map(&mw_hi, w_hi, 2, wdims, wstr, wbox)) != DSEP_OK) return rc;
    if (passes == 3) {
        if ((rc = make_map(&ma_lo, a_lo, 4, adims, astr, abox)) != DSEP_OK) return rc;
        if ((rc = make_map(&mw_lo, w_lo, 2, wdims, wstr, wbox)) != DSEP_OK) return rc;
    } else {
        ma_lo = ma_hi;
        mw_lo = mw_hi;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    switch (NT) {
        case 16: return launch_conv<16>(ma_hi, ma_lo, mw_hi, mw_lo, p, s);
        case 64: return launch_conv<64>(ma_hi, ma_lo, mw_hi, mw_lo, p, s);
        default: return launch_conv<128>(ma_hi, ma_lo, mw_hi, mw_lo, p, s);
    }
}
