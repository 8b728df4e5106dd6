// The wide-tile form of the fused convolution (conv_fused.cu): same operation, same operand planes, same roles —
//
//   out = scale * ( acc_scale * ( conv(act(x * sc + sh)) + conv1x1(shortcut) ) + bias + film + residual )
//
// — with the GEMM transposed.  conv_fused_kernel computes D[pixel, cout] per 128-pixel tile: both tcgen05 operands
// (128 patch rows, 128 weight rows) come from shared memory, 8 KB per 64-clock instruction, i.e. the tensor core alone
// keeps the shared-memory banks busy every cycle, and the builders' stores, the weight fill (576 KB per 128 pixels) and
// the epilogue's transpose staging compete for what is left (profiles/conv_r2a.md: tensor pipe 60 % active, epilogue
// warps 98 % busy, MIO stalls on every shared-memory access).  Here
//
//   D[cout, pixel] = W[cout, K] x P[pixel, K]^T,   M = 128 output channels, N = 256 pixels (8 wide x 32 high),
//
// so that
//   * one instruction (128 clocks) reads 4 KB of weights + 8 KB of patch: 96 B/clk instead of 128;
//   * a weight stage serves 256 pixels: the weight fill per pixel halves (and so do the L2 -> SM bytes);
//   * TMEM lanes are output CHANNELS: a warp's 32 lanes hold 32 consecutive channels of one pixel, so the epilogue
//     stores 128 contiguous bytes per instruction straight from registers — no shared-memory transpose — and the
//     GroupNorm statistics, bias and FiLM row are per-lane scalars (no shuffles, no scratch).
// Shared-memory traffic per 128 pixels: 864 (operands) + 288 (weight fill) + 92 (builders) KB against 1152 + 576 + 220.
//
// Both correction products accumulate into the SAME TMEM columns as the fp16 product (the e4m3 planes are scaled so
// that A_lo8 * W_hi8 and A_hi8 * W_lo8 come out at the main product's scale: corr_rel == 1, see ConvWeight.planes8),
// which is what lets two 256-column accumulators double-buffer in the 512 TMEM columns.
//
// Layout: 2 patch slots x 2 planes x 340 rows x 128 B (10 x 34 halo patch of 64 channels; a 1x1 conv / shortcut uses
// 256 rows), 3 weight stages of ONE plane of one (tap, K-block) each (128 rows x 128 B, fetched half by each CTA of the
// pair and multicast to both).  The builders fill a patch as two halves with the geometry of conv_fused.cu's patch
// (10 x 18 rows / 8 x 16 rows), so the builder code is shared (conv_builders.cuh).
#include "conv_builders.cuh"

namespace dsep {

constexpr int kWThreads = 512;
constexpr int kWBuilderWarps = 8;
constexpr int kWEpiWarps = 4;

struct WideCfg {
    static constexpr int kTileH = 32, kTileW = 8;
    static constexpr int kPlaneBytes = 43 * 1024;               // 340 rows x 128 B = 43520 in a 1024-aligned slot
    static constexpr int kNA = 2;                               // patch ring slots
    static constexpr int kAStage = 2 * kPlaneBytes;
    static constexpr int kWStage = 128 * 128;                   // one plane of a weight stage: 128 rows x 64 channels
    static constexpr int kWStages = 3;
    static constexpr int kRingBytes = kNA * kAStage + kWStages * kWStage;
    static constexpr int kSmemBytes = kRingBytes + 512 + 1024;
    static constexpr int kTmemCols = 512;                       // 2 accumulators x 256 pixels
    static constexpr uint32_t kHalfRows3 = 180, kHalfRows1 = 128;   // rows of one half patch (3x3 halo / 1x1)
};
static_assert(WideCfg::kSmemBytes <= 232448, "conv_wide: shared memory");
static_assert(340 * 128 <= WideCfg::kPlaneBytes, "conv_wide: patch plane");

// CS: the output's channel count when it is 128 or 256 (the pixel stride of the epilogue's per-lane loads / stores
// becomes an immediate: their address arithmetic was ~40 % of the epilogue's instructions), 0 = run-time value.
template <int CS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kWThreads, 1)
conv_wide_kernel(const __grid_constant__ CUtensorMap tm_w_hi, const __grid_constant__ CUtensorMap tm_w_lo,
                 const __grid_constant__ CUtensorMap tm_w2_hi, const __grid_constant__ CUtensorMap tm_w2_lo,
                 const ConvParams p) {
    using Cfg = WideCfg;
    constexpr int NS = Cfg::kWStages;
    constexpr int NA = Cfg::kNA;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* stage_base = smem;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kRingBytes);
    uint64_t* full = bars;             // [NS]   weights: TMA -> MMA
    uint64_t* empty = bars + NS;       // [NS]   MMA -> TMA (both CTAs of the pair commit)
    uint64_t* tfull = bars + 2 * NS;   // [2]    MMA -> epilogue
    uint64_t* tempty = tfull + 2;      // [2]    epilogue -> MMA
    uint64_t* afull = tempty + 2;      // [NA]   patches: builders -> MMA (8 warps x 2 halves)
    uint64_t* aempty = afull + NA;     // [NA]   MMA -> builders
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(aempty + NA);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_w_hi);
        tma_prefetch_desc(&tm_w_lo);
        if (p.kblocks2 > 0) { tma_prefetch_desc(&tm_w2_hi); tma_prefetch_desc(&tm_w2_lo); }
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < NS; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 2); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], kWEpiWarps); }
        for (int i = 0; i < NA; ++i) { mbar_init(&afull[i], 2 * kWBuilderWarps); mbar_init(&aempty[i], 1); }
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc<Cfg::kTmemCols>(tmem_ptr);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();          // both CTAs' barriers are initialised before any multicast targets them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const int total_patches = p.kblocks + p.kblocks2;
    // tile of this CTA for a schedule item: (pair of M-adjacent pixel tiles, channel tile)
    auto locate = [&](int item, int& w0, int& h0, int& b0) {
        int r = 2 * (item / p.tiles_n) + static_cast<int>(rank);
        const int wt = r % p.tiles_w; r /= p.tiles_w;
        const int ht = r % p.tiles_h; r /= p.tiles_h;
        w0 = wt * Cfg::kTileW; h0 = ht * Cfg::kTileH; b0 = r;       // b0 >= B: the odd tile out of the last pair
    };

    if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (warp == 0 && lane == 0) {
            // ------------------------------------------------------------------ TMA producer: weight stages
            int bs = 0;
            uint32_t bph = 0;
            auto load_plane = [&](const CUtensorMap* map, int kcol, int wrow) {
                mbar_wait(&empty[bs], bph ^ 1u);
                uint8_t* sb = stage_base + NA * Cfg::kAStage + bs * Cfg::kWStage;
                if (p.debug & 2) {
                    mbar_arrive(&full[bs]);
                } else {
                    // each CTA of the pair fetches half of the rows and multicasts them to both
                    mbar_arrive_expect_tx(&full[bs], static_cast<uint32_t>(Cfg::kWStage));
                    tma_load_2d_mc(sb + rank * (Cfg::kWStage / 2), map, &full[bs], kcol,
                                   wrow + static_cast<int>(rank) * 64, 0x3);
                }
                if (++bs == NS) { bs = 0; bph ^= 1u; }
            };
            for (int item = cluster_id; item < p.total_items; item += num_clusters) {
                const int n0 = (item % p.tiles_n) * 128;
                for (int kb = 0; kb < p.kblocks2; ++kb) {
                    load_plane(&tm_w2_hi, kb * 64, n0);
                    load_plane(&tm_w2_lo, kb * 64, n0);
                }
                for (int kb = 0; kb < p.kblocks; ++kb)
                    for (int tap = 0; tap < p.taps; ++tap) {
                        load_plane(&tm_w_hi, kb * 64, tap * p.Cout_pad + n0);
                        load_plane(&tm_w_lo, kb * 64, tap * p.Cout_pad + n0);
                    }
            }
        } else if (warp == 1) {
            // ------------------------------------------------------------------ MMA issuer
            // the whole warp walks the pipeline convergently and one elected lane issues (descriptors and barrier
            // addresses stay in uniform registers, see conv_tc.cu)
            const bool leader = elect_one();
            constexpr uint32_t idesc = umma_idesc_f16(128, 256);
            constexpr uint32_t kHiPatch = ((kPatchW * 128) >> 4) | (1u << 14) | (2u << 29);
            constexpr uint32_t kHiPlain = (1024u >> 4) | (1u << 14) | (2u << 29);
            auto desc = [](uint32_t lo, uint32_t hi) {
                uint64_t d;
                asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
                return d;
            };
            const uint32_t a_ring = smem_u32(stage_base) >> 4;
            const uint32_t w_ring = smem_u32(stage_base + NA * Cfg::kAStage) >> 4;
            const bool do_mma = !(p.debug & 1);
            int bs = 0, as_ = 0;
            uint32_t bph = 0, aph = 0;
            int it = 0;
            for (int item = cluster_id; item < p.total_items; item += num_clusters, ++it) {
                const int acc = it & 1;
                mbar_wait(&tempty[acc], ((it >> 1) & 1) ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * 256;
                uint32_t accumulate = 0;
                // one weight stage (a plane of 128 rows x 64 channels) against one or two patch planes
                auto issue = [&](bool fp8, uint32_t p_word, uint32_t p_word2, uint32_t p_hiword) {
                    mbar_wait(&full[bs], bph);
                    tc_fence_after();
                    const uint32_t w_word = w_ring + bs * (Cfg::kWStage >> 4);
                    if (leader) {
                        if (do_mma) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const uint64_t wd = desc(w_word + 2 * k, kHiPlain);
                                const uint64_t pd = desc(p_word + 2 * k, p_hiword);
                                if (fp8) umma_e4m3(d_tmem, wd, pd, idesc, accumulate);
                                else umma_f16(d_tmem, wd, pd, idesc, accumulate);
                                accumulate = 1;
                            }
                            if (p_word2 != 0u) {
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    umma_f16(d_tmem, desc(w_word + 2 * k, kHiPlain), desc(p_word2 + 2 * k, p_hiword),
                                             idesc, 1);
                            }
                        }
                        umma_commit_mc(&empty[bs], 0x3);              // frees the weight slot in BOTH CTAs
                    }
                    __syncwarp();
                    if (++bs == NS) { bs = 0; bph ^= 1u; }
                };
                for (int pi = 0; pi < total_patches; ++pi) {
                    const bool second = pi < p.kblocks2;
                    mbar_wait(&afull[as_], aph);
                    tc_fence_after();
                    const uint32_t sa = a_ring + as_ * (Cfg::kAStage >> 4);
                    const uint32_t sa2 = sa + (Cfg::kPlaneBytes >> 4);
                    if (second) {
                        // fp16 shortcut K-block: W_hi x (P_hi, P_lo), then W_lo x P_hi
                        issue(false, sa, sa2, kHiPlain);
                        issue(false, sa, 0u, kHiPlain);
                    } else if (p.taps == 9) {
#pragma unroll 1
                        for (int tap = 0; tap < 9; ++tap) {
                            const uint32_t off = static_cast<uint32_t>((tap / 3) * kPatchW + tap % 3) * 8u;
                            issue(false, sa + off, 0u, kHiPatch);     // W_hi  x A_hi            (fp16)
                            issue(true, sa2 + off, 0u, kHiPatch);     // [W_hi8 ; W_lo8] x [A_lo8 | A_hi8]  (e4m3)
                        }
                    } else {
                        issue(false, sa, 0u, kHiPlain);
                        issue(true, sa2, 0u, kHiPlain);
                    }
                    if (leader) umma_commit(&aempty[as_]);
                    __syncwarp();
                    if (++as_ == NA) { as_ = 0; aph ^= 1u; }
                }
                if (leader) umma_commit(&tfull[acc]);
                __syncwarp();
            }
        }
    } else if (warp < 4 + kWBuilderWarps) {
        // ---------------------------------------------------------------------- patch builders
        asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");
        const int wtid = static_cast<int>(threadIdx.x) - 128;
        uint32_t jchunk = static_cast<uint32_t>(wtid & 7);
        uint32_t r0 = static_cast<uint32_t>(wtid >> 3);
        uint32_t a_ring = smem_u32(stage_base);
        // opaque to ptxas: it otherwise recomputes all three from %tid / the shared-memory window per patch ROW
        asm volatile("" : "+r"(jchunk), "+r"(r0), "+r"(a_ring));
        const bool skip = (p.debug & 2) != 0;
        const float a8_lo = p.a8_lo;

        // generator of this CTA's half-patch sequence: tiles in schedule order, per tile the shortcut K-blocks then
        // the main ones (the order the MMA issuer consumes them in), per patch the upper then the lower half.
        // Everything that depends only on (thread, tile) is computed when the tile changes (retile): the pixel index
        // of the thread's first row and the in-image masks of both halves for both patch forms; a plan then costs a
        // source select and one 64-bit multiply-add (it used to redo the index arithmetic and the six-row mask loop
        // per half patch: 11 % of the kernel's instructions, profiles/conv_regions_r2b.md).
        int g_item = cluster_id, g_pi = 0, g_half = 0;
        bool g_valid = g_item < p.total_items;
        const int py3 = static_cast<int>(r0) / kPatchW, px3 = static_cast<int>(r0) - py3 * kPatchW;   // 3x3 halo form
        const int py1 = static_cast<int>(r0) >> 3, px1 = static_cast<int>(r0) & 7;                     // 1x1 form
        const bool act3 = r0 < 30u;
        uint32_t sm3 = 0;                       // bits 6h + u: row u of half h exists (the patch has 34 rows of 10)
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int u = 0; u < 6; ++u)
                if (act3 && 18 * h + py3 + 3 * u < 34) sm3 |= 1u << (6 * h + u);
        int pix3 = 0, pix1 = 0, g_bs = 0;       // pixel index of row u = 0 of half 0 (3x3 / 1x1 form); batch entry
        uint32_t inb3 = 0, inb1 = 0;            // in-image masks: bits 6h + u (3x3), 4 bits (1x1: whole tiles)
        auto retile = [&]() {
            int w0, h0, b0;
            locate(g_item, w0, h0, b0);
            const bool b_ok = b0 < p.B;
            g_bs = b_ok ? b0 : 0;
            const int w3 = w0 - 1 + px3;
            const bool col3 = b_ok && w3 >= 0 && w3 < p.W;
            uint32_t m = 0;
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int u = 0; u < 6; ++u) {
                    const int hh = h0 - 1 + 18 * h + py3 + 3 * u;
                    if (col3 && hh >= 0 && hh < p.H) m |= 1u << (6 * h + u);
                }
            inb3 = skip ? 0u : (m & sm3);
            inb1 = (b_ok && !skip) ? 0xFu : 0u;
            pix3 = (g_bs * p.H + h0 - 1 + py3) * p.W + w3;
            pix1 = (g_bs * p.H + h0 + py1) * p.W + w0 + px1;
        };
        if (g_valid) retile();
        // A plan is kept COMPACT between its uses — source pointer + row step (dead once the loads are issued), table
        // index, and one word of bits: in-image mask (6) | stored-row mask (6) << 8 | form << 16 | half << 18 — and
        // expanded into a PatchPlan only where a helper wants one: with two full plans live across the row bodies the
        // builders sat at their register cap and ptxas rematerialised thread indices and the shared-memory base per
        // ROW (~15 of ~110 instructions).
        struct Plan { const float* src; uint32_t step, so, bits; };
        constexpr uint32_t kFormHalo = 0u, kFormPlain = 1u, kFormShort = 2u;      // 3x3 main, 1x1 main, fp16 shortcut
        auto expand = [&](const Plan& q) {
            PatchPlan d;
            const uint32_t form = (q.bits >> 16) & 3u;
            d.src = q.src; d.step = q.step; d.so = q.so;
            d.inb = q.bits & 63u; d.smask = (q.bits >> 8) & 63u;
            d.krows = form == kFormHalo ? 30u : 32u;
            d.niter = form == kFormHalo ? (act3 ? 6u : 0u) : 4u;
            d.second = form == kFormShort;
            d.mode = form == kFormShort ? 0 : (p.fact ? 2 : 1);
            return d;
        };
        auto next_plan = [&](Plan& q) {            // plan of (g_item, g_pi, g_half), then advance
            const bool second = g_pi < p.kblocks2;
            const bool halo3 = !second && p.taps == 9;
            const int kb = second ? g_pi : g_pi - p.kblocks2;
            const int c = kb * 64 + static_cast<int>(jchunk) * 8;
            const float* x0 = second ? p.gx0 : p.fx0;
            const float* x1 = second ? p.gx1 : p.fx1;
            const int C0 = second ? p.gC0 : p.fC0, C1 = second ? p.gC1 : p.fC1;
            const float* src; int cs, cl;
            if (c < C0) { src = x0; cs = C0; cl = c; } else { src = x1; cs = C1; cl = c - C0; }
            const int pix = halo3 ? pix3 + g_half * 18 * p.W : pix1 + g_half * 16 * p.W;
            q.src = src + (static_cast<long long>(pix) * cs + cl);
            q.step = static_cast<uint32_t>((halo3 ? 3 : 4) * p.W * cs);
            q.so = static_cast<uint32_t>(g_bs * (C0 + C1) + c);
            const uint32_t inb = halo3 ? (inb3 >> (6 * g_half)) & 63u : inb1;
            const uint32_t smask = halo3 ? (sm3 >> (6 * g_half)) & 63u : 0xFu;
            const uint32_t form = halo3 ? kFormHalo : (second ? kFormShort : kFormPlain);
            q.bits = inb | (smask << 8) | (form << 16) | (static_cast<uint32_t>(g_half) << 18);
            if (++g_half == 2) {
                g_half = 0;
                if (++g_pi == total_patches) {
                    g_pi = 0;
                    g_item += num_clusters;
                    g_valid = g_item < p.total_items;
                    if (g_valid) retile();
                }
            }
        };

        // two register sets: while one half patch is converted out of set X, the loads of the following one are in
        // flight into set Y, and vice versa (conv_builders.cuh: load_rows / touch_rows and why they are ordered so)
        float4 vx[6][2], vy[6][2];
        Plan qx, qy;
        int as_ = 0, halves_done = 0;
        uint32_t aph = 0;
        const float negzero = -(p.acc_scale * 0.0f);
        const bool silu = p.fact != 0;
        auto wait_slot = [&]() { if (halves_done == 0) mbar_wait(&aempty[as_], aph ^ 1u); };
        auto load = [&](float4 (&v)[6][2], const Plan& q) { const PatchPlan d = expand(q); load_rows(v, d); };
        auto touch = [&](float4 (&v)[6][2], const Plan& q) {
            const PatchPlan d = expand(q);
            touch_rows(v, d, p.fsc, p.fsh, negzero);
        };
        auto build = [&](const float4 (&v)[6][2], const Plan& q) {
            const PatchPlan d = expand(q);
            const uint32_t form = (q.bits >> 16) & 3u, half = (q.bits >> 18) & 1u;
            const uint32_t off = half * (form == kFormHalo ? Cfg::kHalfRows3 : Cfg::kHalfRows1) * 128u;
            uint32_t slot = a_ring + static_cast<uint32_t>(as_) * Cfg::kAStage + off;
            asm volatile("" : "+r"(slot));      // once per half patch, not per row
            constexpr uint32_t PS = Cfg::kPlaneBytes;
            if (form == kFormShort) {
                convert_rows_g<false, 32, 4, true, 6, 0>(v, d, slot, r0, jchunk, 0.f, PS, 0u);
            } else if (form == kFormHalo) {
                const uint32_t rbias = half * (Cfg::kHalfRows3 & 7u);
                if (silu) convert_rows_g<true, 30, 6, true, 6, 1>(v, d, slot, r0, jchunk, a8_lo, PS, rbias);
                else convert_rows_g<true, 30, 6, true, 6, 0>(v, d, slot, r0, jchunk, a8_lo, PS, rbias);
            } else {
                if (silu) convert_rows_g<true, 32, 4, true, 6, 1>(v, d, slot, r0, jchunk, a8_lo, PS, 0u);
                else convert_rows_g<true, 32, 4, true, 6, 0>(v, d, slot, r0, jchunk, a8_lo, PS, 0u);
            }
            // each builder warp publishes its share of each half (afull counts 8 warps x 2 halves)
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> async proxy (tensor core)
            __syncwarp();
            if (lane == 0) mbar_arrive(&afull[as_]);
            if (++halves_done == 2) {
                halves_done = 0;
                if (++as_ == NA) { as_ = 0; aph ^= 1u; }
            }
        };
        bool has_x = g_valid, has_y = false;
        if (has_x) {
            next_plan(qx);
            load(vx, qx);
            has_y = g_valid;
            if (has_y) next_plan(qy);
        }
        while (has_x) {
            touch(vx, qx);
            wait_slot();
            if (has_y) load(vy, qy);
            build(vx, qx);
            has_x = g_valid;
            if (has_x) next_plan(qx);
            if (!has_y) break;
            touch(vy, qy);
            wait_slot();
            if (has_x) load(vx, qx);
            build(vy, qy);
            has_y = g_valid;
            if (has_y) next_plan(qy);
        }
    } else {
        // ---------------------------------------------------------------------- epilogue (4 warps)
        asm volatile("setmaxnreg.inc.sync.aligned.u32 168;");
        const int wq = warp & 3;               // TMEM lane quarter == warp_id % 4: channels n0 + 32 wq + lane
        const float as2 = p.acc_scale * p.scale;
        const bool store = !(p.debug & 4);
        const uint32_t C = CS ? static_cast<uint32_t>(CS) : static_cast<uint32_t>(p.cout_store);
        int run_b = -1, run_n0 = -1;
        float s1 = 0.f, s2 = 0.f;              // running GroupNorm sums of this lane's channel over one batch entry
        float bz = 0.f;                        // (bias + FiLM) * scale of this lane's channel
        auto flush_stats = [&]() {
            if (p.stats == nullptr || run_b < 0 || run_b >= p.B) return;
            const int n = run_n0 + wq * 32 + lane;
            if (n < p.cout_store) {
                double* st = p.stats + (static_cast<size_t>(run_b) * p.cout_store + n) * 2;
                atomicAdd(st + 0, static_cast<double>(s1));
                atomicAdd(st + 1, static_cast<double>(s2));
            }
        };
        float res[2][32];          // residual values of this lane's channel, double-buffered over the 32-pixel chunks
        const float negzero_e = -(p.acc_scale * 0.0f);      // opaque -0: x * 1 + (-0) == x exactly
        int it = 0;
        for (int item = cluster_id; item < p.total_items; item += num_clusters, ++it) {
            const int as = it & 1;
            const int n0 = (item % p.tiles_n) * 128;
            int w0, h0, b0;
            locate(item, w0, h0, b0);
            const int n = n0 + wq * 32 + lane;
            const bool ok = b0 < p.B && n < p.cout_store;        // whole tiles only (H % 32 == 0, W % 8 == 0)
            if (b0 != run_b || n0 != run_n0) {
                flush_stats();
                run_b = b0; run_n0 = n0;
                s1 = 0.f; s2 = 0.f;
                bz = 0.f;
                if (ok) {
                    if (p.bias != nullptr) bz = __ldg(p.bias + n);
                    if (p.film != nullptr) bz += __ldg(p.film + static_cast<size_t>(b0) * p.film_stride + n);
                }
                bz *= p.scale;
            }
            // pixel j of chunk c: image row h0 + 4 c + (j >> 3), column w0 + (j & 7)
            const size_t e0 = ((static_cast<size_t>(ok ? b0 : 0) * p.H + h0) * p.W + w0) * C + (ok ? n : 0);
            float* const out0 = p.out + e0;
            const float* const res0 = p.residual != nullptr ? p.residual + e0 : nullptr;
            const uint32_t d_row = static_cast<uint32_t>(p.W) * C;
            auto load_res = [&](int c, float (&r)[32]) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    float v = 0.f;
                    if (ok) {
                        const float* q = res0 + static_cast<size_t>(4 * c + (j >> 3)) * d_row + (j & 7) * C;
                        asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(q));
                    }
                    r[j] = v;
                }
            };
            // (an L2 prefetch of the NEXT tile's residual lines from here was measured 3-4 % slower, burst and sustained)
            if (res0 != nullptr) load_res(0, res[0]);
            const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + as * 256;
            mbar_wait(&tfull[as], (it >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                if (res0 != nullptr) {
                    // touch the current buffer BEFORE the next burst is issued (all these loads share one scoreboard:
                    // a consumer of the current chunk would otherwise wait for the next chunk's loads as well); one
                    // dependent instruction on its newest load is enough, the scale rides in the FMA below
                    res[c & 1][31] = fmaf(res[c & 1][31], 1.0f, negzero_e);   // exact identity on the NEWEST load
                    __syncwarp();
                    if (c + 1 < 8) load_res(c + 1, res[(c + 1) & 1]);
                }
                uint32_t v[32];
                tmem_ld_32x32(t_addr + c * 32, v);
                tmem_ld_wait();
                if (c == 7) {   // TMEM fully drained by this warp: hand the accumulator back
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tempty[as]);
                }
                if (ok) {
                    float* const orow = out0 + static_cast<size_t>(4 * c) * d_row;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        float a = fmaf(__uint_as_float(v[j]), as2, bz);
                        if (res0 != nullptr) a = fmaf(res[c & 1][j], p.scale, a);
                        if (store) orow[static_cast<size_t>(j >> 3) * d_row + (j & 7) * C] = a;
                        s1 += a;
                        s2 = fmaf(a, a, s2);
                    }
                }
            }
        }
        flush_stats();
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();          // the peer may still multicast into / arrive on this CTA's shared memory
    if (warp == 2) tmem_dealloc<Cfg::kTmemCols>(tmem_base);
}

template <int CS>
static int launch_wide(const ConvMaps& m, const ConvParams& p, cudaStream_t stream) {
    constexpr int kSmem = WideCfg::kSmemBytes;
    static PerDeviceAttr attr;
    const cudaError_t attr_err = set_max_smem_once(attr, conv_wide_kernel<CS>, kSmem);
    if (attr_err != cudaSuccess) {
        set_error("cudaFuncSetAttribute(conv_wide_kernel): %s", cudaGetErrorString(attr_err));
        return DSEP_ERR_CUDA;
    }
    const int max_clusters = conv_num_sms() / 2;
    const int grid = 2 * (p.total_items < max_clusters ? p.total_items : max_clusters);
    conv_wide_kernel<CS><<<grid, kWThreads, kSmem, stream>>>(m.w_hi, m.w_lo, m.w2_hi, m.w2_lo, p);
    return check_launch("conv_wide_kernel");
}

int launch_conv_wide(const ConvMaps& m, const ConvParams& p, cudaStream_t stream) {
    if (p.cout_store == 128) return launch_wide<128>(m, p, stream);
    if (p.cout_store == 256) return launch_wide<256>(m, p, stream);
    return launch_wide<0>(m, p, stream);
}

}  // namespace dsep
