// Patch builders shared by the role-split convolution kernels (conv_fused.cu: 128-pixel tiles, pixels on the MMA's M
// side; conv_wide.cu: 256-pixel tiles, pixels on the N side): per-thread plan of one patch (or half patch), the
// burst loads, the in-place GroupNorm affine and the SiLU + (hi, lo) / (hi, e4m3) split + swizzled stores.
#pragma once
#include "conv_tc.cuh"

namespace dsep {

// One builder thread's view of one patch: its rows are r = r0 + u * krows (u < niter), all in one patch column,
// kdy image rows apart — 10 x 18 halo patch: krows 30, niter 6 (threads with r0 >= 30 idle); 8 x 16: 32, 4.
struct PatchPlan {
    const float* src;      // this thread's 8 channels of row u = 0 (only dereferenced where inb says so)
    uint32_t step;         // elements between rows u and u + 1
    uint32_t inb;          // bit u: row u exists and lies inside the image
    uint32_t smask;        // bit u: row u is stored at all (conv_wide.cu: the second half patch ends after 160 rows)
    uint32_t krows;        // 30 / 32 (conv_wide.cu: 60 / 64)
    uint32_t niter;        // 6 / 4 (conv_wide.cu: 3 / 2; 0: this thread has no rows)
    uint32_t so;           // index of this thread's 8 channels in the sc / sh tables
    int mode;              // 0 raw (shortcut operand), 1 affine, 2 affine + SiLU
    bool second;           // shortcut K-block (fp16 (hi, lo) planes even in the e4m3 mode)
};

// d = a * b + c with fp16 a, b and fp32 c, d (sm_100 mixed-precision FMA, SASS FHFMA): y - hi in ONE instruction
// instead of a half -> float conversion and a subtraction (exact either way: both products are exact in fp32)
__device__ __forceinline__ float fma_f32_f16(uint16_t a, uint16_t b, float c) {
    float d;
    asm("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(d) : "h"(a), "h"(b), "f"(c));
    return d;
}

__device__ __forceinline__ uint32_t e4m3x2_from_f16x2(uint32_t h2) {
    uint16_t r;
    asm("cvt.rn.satfinite.e4m3x2.f16x2 %0, %1;" : "=h"(r) : "r"(h2));
    return r;
}

// Builder pipeline per patch and thread (rows live in a register set v[6][2]):
//   load_rows   — ONE burst of global loads (6 x 32 bytes in flight per thread) into the set;
//   touch_rows  — a whole patch later: the first arithmetic on the loaded values, in place (GroupNorm affine
//                 x * sc + sh, or an exact identity add for raw shortcut operands).  This is where the warp waits
//                 for the burst — and it runs BEFORE the next burst (into the other set) is issued, with the slot
//                 wait's spin loop in between so that ptxas cannot reorder the two.  Reason: ptxas puts every one
//                 of these loads on the same scoreboard; a consumer that waits for "its" loads therefore waits for
//                 everything in flight on that scoreboard.  The first prefetching version re-issued each row's
//                 loads right after converting that row and lost a full memory round trip per ROW (1.28 ms for the
//                 builders alone against 1.11 ms without any prefetch);
//   convert_rows — SiLU, (hi, lo) / (hi, e4m3) split, swizzled stores: depends on touch_rows' results only.
// NR: rows per thread and register set (6 with 8 builder warps, 3 with conv_wide.cu's 16)
template <int NR>
__device__ __forceinline__ void load_rows(float4 (&v)[NR][2], const PatchPlan& d) {
    // rows that are not loaded (outside the image: the conv pads the ACTIVATED tensor with zeros) are zero-filled here,
    // and stay zero through touch_rows (skipped for them) and convert_store (SiLU(0) = 0, both planes of 0 are 0) — so
    // the row body needs no per-row selects.  Interior tiles (every row loaded) skip the fill with one branch.
    if (d.inb != (1u << d.niter) - 1u) {
#pragma unroll
        for (int u = 0; u < NR; ++u) {
            if (!((d.inb >> u) & 1u)) {
                v[u][0] = make_float4(0.f, 0.f, 0.f, 0.f);
                v[u][1] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    }
#pragma unroll
    for (int u = 0; u < NR; ++u) {
        if ((d.inb >> u) & 1u) {
            const float* q = d.src + static_cast<size_t>(u) * d.step;
            ldg_stream8(q, v[u][0], v[u][1]);
        }
    }
}

// mode 2 (SiLU follows): the affine is pre-scaled by c = -log2(e), so v holds u = c * t and convert_store needs no
// multiply in front of ex2: y = t / (1 + 2^u) = u * rcp(c + c * 2^u).
constexpr float kNegLog2e = -1.4426950408889634f;
template <int NR>
__device__ __forceinline__ void touch_rows(float4 (&v)[NR][2], const PatchPlan& d, const float* __restrict__ sc,
                                           const float* __restrict__ sh, float negzero) {
    if (d.mode == 0 || (d.mode == 1 && sc == nullptr)) {
        // raw operand (shortcut K-blocks, FIR-fed / im2col main operands): nothing to apply — ONE exact identity
        // (x * 1 + (-0) == x; negzero is opaque to the compiler) per row is enough to make the warp wait for that row's
        // load here, before the next burst goes out (the load writes all eight registers under one scoreboard)
#pragma unroll
        for (int u = 0; u < NR; ++u)
            if ((d.inb >> u) & 1u) v[u][0].x = fmaf(v[u][0].x, 1.0f, negzero);
        return;
    }
    float k_sc[8], k_sh[8];
    if (d.mode != 0 && sc != nullptr) {
        const float4 a0 = __ldg(reinterpret_cast<const float4*>(sc + d.so));
        const float4 a1 = __ldg(reinterpret_cast<const float4*>(sc + d.so + 4));
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(sh + d.so));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(sh + d.so + 4));
        k_sc[0] = a0.x; k_sc[1] = a0.y; k_sc[2] = a0.z; k_sc[3] = a0.w;
        k_sc[4] = a1.x; k_sc[5] = a1.y; k_sc[6] = a1.z; k_sc[7] = a1.w;
        k_sh[0] = b0.x; k_sh[1] = b0.y; k_sh[2] = b0.z; k_sh[3] = b0.w;
        k_sh[4] = b1.x; k_sh[5] = b1.y; k_sh[6] = b1.z; k_sh[7] = b1.w;
    } else {      // raw operand: x * 1 + (-0) == x exactly (negzero is opaque to the compiler)
#pragma unroll
        for (int e = 0; e < 8; ++e) { k_sc[e] = 1.0f; k_sh[e] = negzero; }
    }
    if (d.mode == 2) {
#pragma unroll
        for (int e = 0; e < 8; ++e) { k_sc[e] *= kNegLog2e; k_sh[e] *= kNegLog2e; }
    }
#pragma unroll
    for (int u = 0; u < NR; ++u) {
        if ((d.inb >> u) & 1u) {
            v[u][0].x = fmaf(v[u][0].x, k_sc[0], k_sh[0]); v[u][0].y = fmaf(v[u][0].y, k_sc[1], k_sh[1]);
            v[u][0].z = fmaf(v[u][0].z, k_sc[2], k_sh[2]); v[u][0].w = fmaf(v[u][0].w, k_sc[3], k_sh[3]);
            v[u][1].x = fmaf(v[u][1].x, k_sc[4], k_sh[4]); v[u][1].y = fmaf(v[u][1].y, k_sc[5], k_sh[5]);
            v[u][1].z = fmaf(v[u][1].z, k_sc[6], k_sh[6]); v[u][1].w = fmaf(v[u][1].w, k_sc[7], k_sh[7]);
        }
    }
}

// One pair of float4 (8 consecutive channels of one patch row, affine already applied) -> the row chunk of both
// planes.  silu: the values are u = -log2(e) * t and y = t / (1 + 2^u) = u * rcp(c + c 2^u) (ex2.approx.ftz +
// rcp.approx.ftz, no range fix-ups).  E4M3: second plane = [A_lo8 x 8 | A_hi8 x 8] with A_hi8 converted straight from
// the packed fp16 pairs (the host passes a8_exp = 0) and A_lo8 = e4m3(lo * a8_lo); else the fp16 lo plane.
// SILU: 1 / 0 = known at compile time (no copy of the row in front of a run-time branch), -1 = the `silu` argument.
template <bool E4M3, int SILU = -1>
__device__ __forceinline__ void convert_store(const float4 v0, const float4 v1, bool silu, uint32_t dst_hi,
                                              uint32_t dst_2, float a8_lo) {
    if (SILU >= 0) silu = SILU != 0;
    // rows outside the image arrive as zeros (load_rows) and come out as zeros: one straight-line body per row
    uint32_t hi[4], lo[4];
    float y[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    if (silu) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            float ex, rc;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(y[e]));
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(fmaf(ex, kNegLog2e, kNegLog2e)));
            y[e] *= rc;
        }
    }
    if (E4M3) {
        // a8_lo = 2^(11 + a8_exp): l = (y - hi) * a8_lo = y * a8_lo - hi * a8_lo, the second term folded into an FHFMA
        // with the fp16 constant -a8_lo (exact: power of two; the dispatch admits a8_exp = 0 only, -2048 is 0xE800)
        float l[8];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi[e]) : "f"(y[2 * e + 1]), "f"(y[2 * e]));
            const uint16_t h0 = static_cast<uint16_t>(hi[e] & 0xFFFFu), h1 = static_cast<uint16_t>(hi[e] >> 16);
            l[2 * e] = fma_f32_f16(h0, static_cast<uint16_t>(0xE800u), y[2 * e] * a8_lo);
            l[2 * e + 1] = fma_f32_f16(h1, static_cast<uint16_t>(0xE800u), y[2 * e + 1] * a8_lo);
        }
        lo[0] = e4m3x4(l[0], l[1], l[2], l[3]);
        lo[1] = e4m3x4(l[4], l[5], l[6], l[7]);
        lo[2] = e4m3x2_from_f16x2(hi[0]) | (e4m3x2_from_f16x2(hi[1]) << 16);
        lo[3] = e4m3x2_from_f16x2(hi[2]) | (e4m3x2_from_f16x2(hi[3]) << 16);
    } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) {      // (hi, lo) fp16 pairs; lo = y - hi by FHFMA (hi * -1 + y)
            asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi[e]) : "f"(y[2 * e + 1]), "f"(y[2 * e]));
            const uint16_t h0 = static_cast<uint16_t>(hi[e] & 0xFFFFu), h1 = static_cast<uint16_t>(hi[e] >> 16);
            const float d0 = fma_f32_f16(h0, static_cast<uint16_t>(0xBC00u), y[2 * e]);
            const float d1 = fma_f32_f16(h1, static_cast<uint16_t>(0xBC00u), y[2 * e + 1]);
            asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo[e]) : "f"(d1), "f"(d0));
        }
    }
    sts128(dst_hi, hi[0], hi[1], hi[2], hi[3]);
    sts128(dst_2, lo[0], lo[1], lo[2], lo[3]);
}

// One thread's rows of a (half) patch -> both planes: rows r0 + u * KROWS (u < NITER) of the patch part that starts at
// slot_addr, whose first row has index rbias (mod 8) inside its 1024-byte swizzle atom.  Compile-time geometry keeps
// the row / swizzle arithmetic in immediates.  MASKED: rows without their smask bit are not stored at all.
template <bool E4M3, int KROWS, int NITER, bool MASKED, int NR, int SILU = -1>
__device__ __forceinline__ void convert_rows_g(const float4 (&v)[NR][2], const PatchPlan& cur, uint32_t slot_addr,
                                               uint32_t r0, uint32_t jchunk, float a8_lo, uint32_t plane_stride,
                                               uint32_t rbias) {
    static_assert(NITER <= NR, "convert_rows_g: register set too small");
    if (cur.niter == 0u) return;                 // threads beyond the patch (r0 >= KROWS)
    const bool silu = cur.mode == 2;
    const uint32_t base = slot_addr + r0 * 128u;
#pragma unroll
    for (int u = 0; u < NITER; ++u) {
        // (r0 + u * KROWS) & 7 == (r0 + (u * KROWS & 7)) & 7: only the low bits of r0 are runtime
        const uint32_t sw = ((jchunk ^ ((r0 + rbias + ((u * KROWS) & 7u)) & 7u)) << 4);
        const uint32_t off = base + u * KROWS * 128u + sw;
        if (MASKED && !((cur.smask >> u) & 1u)) continue;
        convert_store<E4M3, SILU>(v[u][0], v[u][1], silu, off, off + plane_stride, a8_lo);
    }
}

// conv_fused.cu's geometry.  HALO3: the 10 x 18 patch of a 3x3 conv (6 rows per thread, 30 patch rows apart) — else
// 8 x 16 (4 rows, 32 apart).
template <bool E4M3, bool HALO3, bool MASKED = false>
__device__ __forceinline__ void convert_rows(const float4 (&v)[6][2], const PatchPlan& cur, uint32_t slot_addr,
                                             uint32_t r0, uint32_t jchunk, float a8_lo,
                                             uint32_t plane_stride = kPatchPlane, uint32_t rbias = 0u) {
    convert_rows_g<E4M3, HALO3 ? 30 : 32, HALO3 ? 6 : 4, MASKED, 6>(v, cur, slot_addr, r0, jchunk, a8_lo, plane_stride,
                                                                   rbias);
}

}  // namespace dsep
