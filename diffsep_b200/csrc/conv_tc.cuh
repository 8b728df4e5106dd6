// Shared declarations of the tensor-core convolution kernels (conv_tc.cu: per-tap / halo kernels and the host
// dispatch; conv_fused.cu: the role-split halo kernel with in-kernel prologue).
#pragma once
#include <stdlib.h>

#include <mutex>

#include "common.cuh"

// passes = 2 (the product default; -DDSEP_FP8_CORR=0 compiles it out): per K=16 step ONE fp16 product hi*hi plus
// ONE e4m3 tensor-core product that carries both correction terms, [A_lo8 | A_hi8] x [W_hi8 ; W_lo8] (K = 32
// bytes, twice the fp16 rate) = 2 tensor-core units per MAC instead of 3.  tools/numerics_study.py: network error
// 4.7e-5 from the operand rounding (budget 1e-4; dropping either correction in a single level-0 conv costs
// 2.4e-4).  Halo kernel only (maps of at least 16 x 8, Cout >= 64); the main operand is either built in-kernel
// or arrives as (fp16 hi, e4m3 correction) planes by TMA (dsep_fir_resample8 writes them).
#ifndef DSEP_FP8_CORR
#define DSEP_FP8_CORR 1
#endif

namespace dsep {

struct ConvParams {
    int B, H, W, Cin, Cout_pad, cout_store;
    int taps;              // 1 or 9
    int tw_log2, th_log2;  // pixel tile: tw x th x tb = 128
    int tiles_w, tiles_h, tiles_b, tiles_n, total_items;   // item = (pair of M-adjacent tiles, channel tile)
    int kblocks;           // Cin / kBK
    int kblocks2;          // Cin2 / kBK of the fused 1x1 shortcut (0: none)
    int passes;            // 1 or 3
    const float* bias;
    const float* film;
    int film_stride;
    const float* residual;
    float scale, acc_scale;
    float* out;
    double* stats;         // [B, cout_store, 2] or null
    // fused prologue (halo mode): the A patch is built in-kernel from fp32 activations instead of
    // arriving as split planes by TMA.  main operand: act(x * sc + sh) of the channel-concatenated
    // [fx0 (fC0 ch) | fx1 (fC1 ch)]; shortcut operand: the raw [gx0 | gx1] (identity).
    const float* fx0; const float* fx1; int fC0, fC1;
    const float* fsc; const float* fsh; int fact;
    const float* gx0; const float* gx1; int gC0, gC1;
#if DSEP_FP8_CORR
    float corr_rel;        // passes = 2: weight of the e4m3 correction accumulator relative to the fp16 one
    float a8_hi, a8_lo;    // passes = 2: power-of-two prescales of the e4m3 activation planes (A_hi, A_lo)
#endif
    int debug;             // DSEP_CONV_DEBUG bitmask: 1 skip MMA issue, 2 skip TMA loads, 4 skip epilogue stores (timing experiments)
};

constexpr int kEpiWarps = 8;
constexpr int kThreads = 128 + kEpiWarps * 32;
// K-block: channels per pipeline stage.  64 (128-byte swizzle) is the default; 32 (64-byte swizzle,
// twice as many half-size stages) was measured 16 % slower: the operand feed is bound by L2->SM
// delivery (~12.6 TB/s with 128-byte rows, ~7 TB/s with 64-byte rows), not by ring depth.
#ifndef DSEP_CONV_BK
#define DSEP_CONV_BK 64
#endif
constexpr int kBK = DSEP_CONV_BK;
static_assert(kBK == 32 || kBK == 64, "K-block must be 32 or 64 channels");
constexpr int kABytes = 128 * kBK * 2;     // one A plane of a stage

// ---- "halo" mode (3x3, maps of at least 16 x 8): the A operand of all nine taps comes from ONE
// shared-memory patch.  The output tile is 8 (w) x 16 (h) pixels; its 10 x 18 input patch of 64
// channels is a single TMA box (out-of-image pixels zero-filled).  Patch row = y_p * 10 + x_p, so the
// 128 operand rows of tap (dy, dx) are 16 groups of 8 consecutive patch rows starting at row
// dy * 10 + dx, 10 rows (1280 B) apart: exactly a K-major UMMA descriptor with SBO = 1280 and a
// shifted start address (the 128-byte swizzle is a function of the shared-memory address, which TMA
// and tcgen05.mma share).  Shared-memory ingest per tile drops from 1152 KB to 668 KB, which is what
// bounded the per-tap kernel (TMA-only time 1.0 ms vs MMA-only 1.2 ms on the level-0 conv).
constexpr int kPatchW = 10, kPatchH = 18;
constexpr int kPatchBytes = kPatchW * kPatchH * 128;          // 23040: one plane of a patch
constexpr int kPatchPlane = 23 * 1024;                         // its 1024-aligned slot
constexpr int kHaloAStages = 2;

// Builds one 64-channel A patch (PH x PW pixels, row = py * PW + px, 128-byte-swizzled K-major rows of
// (hi, lo) fp16) from fp32 activations: y = act(x * sc[c] + sh[c]), zero outside the image (the conv
// pads the ACTIVATED tensor).  Called by the kBuilders builder threads; wtid = 0..255.  Replaces the
// GroupNorm-apply + SiLU + split pass (and the channel concat) that used to run as its own kernel.
constexpr int kBuilders = 256;   // the 8 worker warps (two warpgroups, 224 registers each after setmaxnreg)

// (hi, lo) fp16 pairs of two floats; out-of-range values saturate to +-65504 instead of becoming NaN
__device__ __forceinline__ void split2_f16(float a, float b, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));
    const float2 back = __half22float2(*reinterpret_cast<const __half2*>(&hi));
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(b - back.y), "f"(a - back.x));
}

__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}

// streaming 16-byte load that does not allocate in L1: with ~210 KB of the SM's 228 KB configured as shared memory
// the L1 is a few KB, and the activation rows streaming through it (46 KB per patch) kept evicting the per-channel
// tables (GroupNorm affine, bias, FiLM), whose reloads then cost an L2 round trip per patch (ncu: 12 % of the
// builders' samples sat on the first use of sc / sh)
// 32 bytes (8 channels of one pixel) in ONE 256-bit load: a whole sector per lane.  Two 16-byte loads would each touch
// half of every sector — harmless with L1 allocation (the second hits), twice the L2 requests without it.
__device__ __forceinline__ void ldg_stream8(const float* p, float4& a, float4& b) {
    asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
}
__device__ __forceinline__ float4 ldg_stream(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

#if DSEP_FP8_CORR
// 8 activations -> 16 bytes of the e4m3 correction plane: [A_lo8 x 8 | A_hi8 x 8] (the weight plane holds
// [W_hi8 x 8 | W_lo8 x 8] at the same bytes, so the K = 32 product sums A_lo*W_hi + A_hi*W_lo)
__device__ __forceinline__ uint32_t e4m3x4(float a, float b, float c, float d) {
    uint16_t p0, p1;
    asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(p0) : "f"(b), "f"(a));
    asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(p1) : "f"(d), "f"(c));
    return static_cast<uint32_t>(p0) | (static_cast<uint32_t>(p1) << 16);
}
#endif

struct ConvMaps {
    CUtensorMap a_hi, a_lo, w_hi, w_lo, a2_hi, a2_lo, w2_hi, w2_lo;
};

int conv_num_sms();
// conv_fused.cu: launches conv_fused_kernel<NT, FP8> (NT = 64 / 128; p.passes = 2 or 3; p.fx0 != nullptr)
int launch_conv_fused(const ConvMaps& m, const ConvParams& p, int NT, cudaStream_t stream);
// conv_wide.cu: conv_wide_kernel (Cout tile 128, passes = 2 with corr_rel == 1, 8 x 32 pixel tiles: H % 32 == 0, W % 8 == 0)
int launch_conv_wide(const ConvMaps& m, const ConvParams& p, cudaStream_t stream);

}  // namespace dsep
