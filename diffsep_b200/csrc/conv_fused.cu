// The network's dominant launch: 3x3 / 1x1 convolution on tcgen05 with the GroupNorm-apply + SiLU + channel
// concat + operand split of its input done in-kernel, role-split so that no role waits on another's latency.
//
//   out = scale * ( acc_scale * ( conv(act(x * sc + sh)) + conv1x1(shortcut) ) + bias + film + residual )
//
// Same arithmetic, operand layouts and barriers-by-name as conv_tc_kernel<NT, halo> (conv_tc.cu), which this
// kernel replaces for Cout >= 64 whenever the main operand is built in-kernel.  What changed, and why
// (DESIGN.md section 5): in the 2-unit mode (fp16 hi*hi + one e4m3 product for both corrections) the
// tensor core needs 0.74 ms for the level-0 conv but the 8 worker warps that both BUILT the operand patches and
// RAN the epilogue needed 1.11 ms on their own — issue slots 38 % used, the rest exposed latency: each patch's
// global loads were issued and immediately waited for, and the epilogue's TMEM / shared / global round trips
// sat in the same instruction stream.  Here:
//   * 16 warps.  Warpgroup 0: TMA weight producer (warp 0), MMA issuer (warp 1, convergent, descriptors in
//     uniform registers), TMEM allocator (warp 2).  Warpgroups 1-2: eight BUILDER warps.  Warpgroup 3: four
//     EPILOGUE warps (one per TMEM lane quarter).  setmaxnreg: 40 / 144 / 184 (128 x 40 + 256 x 144 + 128 x 184
//     = 65536 registers).
//   * builders prefetch: while patch n is converted and stored out of one register set, the global loads of patch
//     n + 1 are in flight into a second set (issued in one burst at the start of patch n), so a patch's load latency
//     overlaps a whole patch of arithmetic — across tile and K-block boundaries.  One code path serves both patch
//     geometries (10 x 18 halo patch of a 3x3 conv, 8 x 16 of a 1x1 / shortcut).
//   * the epilogue never blocks a builder: it runs a full tile behind the MMAs (double-buffered TMEM).
#include "conv_builders.cuh"

namespace dsep {

constexpr int kFThreads = 512;
constexpr int kFBuilderWarps = 8;
constexpr int kFEpiWarps = 4;

// TWO: the CTA pair issues ONE tcgen05.mma.cta_group::2 per step from the leader CTA (M = 256: 128 pixels from each
// CTA, the weight rows split between the two CTAs' shared memory).  Per SM that halves the weight bytes written by
// TMA and read by the tensor core — in the 2-unit mode the 1-CTA kernel needs 1152 KB of operand reads + 576 KB of
// weight fill + 220 KB of builder / epilogue traffic per 128-pixel tile through a 128 B/clk shared-memory pipe:
// 15.2 k clk against 12.4 k clk of tensor time (measured floor 0.92 ms against 0.75 ms).  With the pair: 864 + 288
// + 220 KB = 10.7 k clk.  The smaller weight stages also pay for a third patch slot.
template <int NT, bool TWO>
struct FusedCfg {
    static constexpr int kNA = TWO ? 3 : 2;                    // A-patch ring slots
    static constexpr int kBBytes = NT * 128;                   // one whole weight plane of a stage (64 channels)
    static constexpr int kBPlane = TWO ? kBBytes / 2 : kBBytes;   // this CTA's share of it
    static constexpr int kAStage = 2 * kPatchPlane;
    static constexpr int kBStage = 2 * kBPlane;
    static constexpr int kStatBytes = kFEpiWarps * (NT / 32) * 8 * 32;          // epilogue: running GroupNorm sums
    static constexpr int kBiasBytes = kFEpiWarps * (NT / 32) * 32 * 16;         // epilogue: (bias + FiLM) * scale per lane
    static constexpr int kStagingBytes = kFEpiWarps * 32 * 32 * 4 + kStatBytes + kBiasBytes;
    static constexpr int kAvail = 232448 - 1024 - 512 - kStagingBytes - kNA * kAStage;
    static constexpr int kBStagesMax = kAvail / kBStage;
    static constexpr int kBStages = kBStagesMax > 8 ? 8 : kBStagesMax;
    static constexpr int kRingBytes = kNA * kAStage + kBStages * kBStage;
    static constexpr int kSmemBytes = kRingBytes + kStagingBytes + 512 + 1024;
    static constexpr int kTmemCols = 4 * NT;                   // 2 accumulator stages x (main | correction) x NT
};


template <int NT, bool FP8, bool TWO>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kFThreads, 1)
conv_fused_kernel(const __grid_constant__ CUtensorMap tm_w_hi, const __grid_constant__ CUtensorMap tm_w_lo,
                  const __grid_constant__ CUtensorMap tm_w2_hi, const __grid_constant__ CUtensorMap tm_w2_lo,
                  const ConvParams p) {
    using Cfg = FusedCfg<NT, TWO>;
    static_assert(!TWO || FP8, "the CTA-pair MMA is implemented for the 2-unit mode only");
    constexpr int NS = Cfg::kBStages;
    constexpr int NA = Cfg::kNA;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* stage_base = smem;
    float* staging = reinterpret_cast<float*>(smem + Cfg::kRingBytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kRingBytes + Cfg::kStagingBytes);
    uint64_t* full = bars;             // [NS]   weights: TMA -> MMA
    uint64_t* empty = bars + NS;       // [NS]   MMA -> TMA (both CTAs of the pair commit)
    uint64_t* tfull = bars + 2 * NS;   // [2]    MMA -> epilogue
    uint64_t* tempty = tfull + 2;      // [2]    epilogue -> MMA
    uint64_t* afull = tempty + 2;      // [NA]   patches: builders -> MMA
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
        // 1-CTA MMAs: each CTA's issuer commits to both CTAs' weight slots (multicast weights).  TWO: the leader's
        // issuer is the only one; it collects the builder / epilogue warps of both CTAs on ITS barriers.
        for (int i = 0; i < NS; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], TWO ? 1 : 2); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], (TWO ? 2 : 1) * kFEpiWarps); }
        for (int i = 0; i < NA; ++i) { mbar_init(&afull[i], (TWO ? 2 : 1) * kFBuilderWarps); mbar_init(&aempty[i], 1); }
        fence_mbar_init();
    }
    if (warp == 2) {
        if constexpr (TWO) tmem_alloc_2cta<Cfg::kTmemCols>(tmem_ptr);
        else tmem_alloc<Cfg::kTmemCols>(tmem_ptr);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();          // both CTAs' barriers are initialised before any multicast targets them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const int total_patches = p.kblocks + p.kblocks2;
    // TWO: the barriers a role waits on are completed from the other CTA (multicast commits of the leader's tensor
    // core); such completions do not wake a suspended try_wait early (see mbar_wait_poll), so keep the hint short
    auto role_wait = [](uint64_t* bar, uint32_t parity) {
        if constexpr (TWO) mbar_wait<400u>(bar, parity);
        else mbar_wait(bar, parity);
    };

    if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (warp == 0 && lane == 0) {
            // ------------------------------------------------------------------ TMA producer: weight stages
            int bs = 0;
            uint32_t bph = 0;
            auto load_weights = [&](const CUtensorMap* whi, const CUtensorMap* wlo, int kcol, int wrow) {
                role_wait(&empty[bs], bph ^ 1u);
                uint8_t* sb = stage_base + NA * Cfg::kAStage + bs * Cfg::kBStage;
                const int wrow_h = wrow + static_cast<int>(rank) * (NT / 2);      // my half of the weight rows
                if (p.debug & 2) {
                    if (!TWO || rank == 0) mbar_arrive(&full[bs]);
                } else if constexpr (TWO) {
                    // each CTA keeps its half of the rows of both planes; everything completes on the leader's barrier
                    if (rank == 0) mbar_arrive_expect_tx(&full[bs], 2u * Cfg::kBBytes);
                    tma_load_2d_2sm(sb, whi, &full[bs], kcol, wrow_h);
                    tma_load_2d_2sm(sb + Cfg::kBPlane, wlo, &full[bs], kcol, wrow_h);
                } else {
                    // each CTA of the pair fetches half of the rows of both planes and multicasts them to both
                    mbar_arrive_expect_tx(&full[bs], 2u * Cfg::kBBytes);
                    const int boff = static_cast<int>(rank) * (Cfg::kBBytes / 2);
                    tma_load_2d_mc(sb + boff, whi, &full[bs], kcol, wrow_h, 0x3);
                    tma_load_2d_mc(sb + Cfg::kBBytes + boff, wlo, &full[bs], kcol, wrow_h, 0x3);
                }
                if (++bs == NS) { bs = 0; bph ^= 1u; }
            };
            for (int item = cluster_id; item < p.total_items; item += num_clusters) {
                const int n0 = (item % p.tiles_n) * NT;
                for (int kb = 0; kb < p.kblocks2; ++kb) load_weights(&tm_w2_hi, &tm_w2_lo, kb * 64, n0);
                for (int kb = 0; kb < p.kblocks; ++kb)
                    for (int tap = 0; tap < p.taps; ++tap)
                        load_weights(&tm_w_hi, &tm_w_lo, kb * 64, tap * p.Cout_pad + n0);
            }
        } else if (warp == 1 && !(TWO && rank != 0)) {
            // ------------------------------------------------------------------ MMA issuer (TWO: the leader CTA's)
            // the WHOLE warp walks the pipeline convergently and one elected lane issues: barrier addresses,
            // descriptors and the accumulate flag stay in uniform registers (see conv_tc.cu)
            const bool leader = elect_one();
            constexpr uint32_t idesc_n = umma_idesc_f16(TWO ? 256 : 128, NT);
            constexpr uint32_t idesc_2n = umma_idesc_f16(128, 2 * NT);
            auto mma16 = [&](uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc_flag) {
                if constexpr (TWO) umma_f16_2cta(d, a, b, idesc, acc_flag);
                else umma_f16(d, a, b, idesc, acc_flag);
            };
            auto mma8 = [&](uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc_flag) {
                if constexpr (TWO) umma_e4m3_2cta(d, a, b, idesc, acc_flag);
                else umma_e4m3(d, a, b, idesc, acc_flag);
            };
            auto commit_both = [&](uint64_t* bar) {       // arrive on this barrier in BOTH CTAs
                if constexpr (TWO) umma_commit_2cta_mc(bar, 0x3);
                else umma_commit_mc(bar, 0x3);
            };
            auto commit_own = [&](uint64_t* bar) {        // TWO: the follower's roles wait on their own copies
                if constexpr (TWO) umma_commit_2cta_mc(bar, 0x3);
                else umma_commit(bar);
            };
            constexpr uint32_t kHiPatch = ((kPatchW * 128) >> 4) | (1u << 14) | (2u << 29);
            constexpr uint32_t kHiPlain = (1024u >> 4) | (1u << 14) | (2u << 29);
            auto desc = [](uint32_t lo, uint32_t hi) {
                uint64_t d;
                asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
                return d;
            };
            const uint32_t a_ring = smem_u32(stage_base) >> 4;
            const uint32_t b_ring = smem_u32(stage_base + NA * Cfg::kAStage) >> 4;
            const bool do_mma = !(p.debug & 1);
            // TWO: these barriers are completed by the peer CTA's arrivals / TMA bytes as well -> poll (see mbar_wait_poll)
            auto mma_wait = [&](uint64_t* bar, uint32_t parity) {
                if constexpr (TWO) mbar_wait_poll(bar, parity);
                else mbar_wait(bar, parity);
            };
            int bs = 0, as_ = 0;
            uint32_t bph = 0, aph = 0;
            int it = 0;
            for (int item = cluster_id; item < p.total_items; item += num_clusters, ++it) {
                const int acc = it & 1;
                mma_wait(&tempty[acc], ((it >> 1) & 1) ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * 2 * NT;
                uint32_t accumulate = 0, accumulate8 = 0;
                auto issue = [&](uint32_t a_word, uint32_t a_hiword, bool main_kb) {
                    mma_wait(&full[bs], bph);
                    tc_fence_after();
                    const uint32_t b_word = b_ring + bs * (Cfg::kBStage >> 4);
                    if (leader) {
                        if (do_mma) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const uint64_t a_hi = desc(a_word + 2 * k, a_hiword);
                                const uint64_t a_2 = desc(a_word + (kPatchPlane >> 4) + 2 * k, a_hiword);
                                const uint64_t b_hi = desc(b_word + 2 * k, kHiPlain);
                                const uint64_t b_2 = desc(b_word + (Cfg::kBPlane >> 4) + 2 * k, kHiPlain);
                                if (FP8) {
                                    mma16(d_tmem, a_hi, b_hi, idesc_n, accumulate);                // hi*hi -> [0, NT)
                                    if (main_kb) {   // [A_lo8 | A_hi8] x [W_hi8 ; W_lo8], K = 32 -> [NT, 2NT)
                                        mma8(d_tmem + NT, a_2, b_2, idesc_n, accumulate8);
                                        accumulate8 = 1;
                                    } else {         // fp16 shortcut K-block: the two corrections join [0, NT)
                                        mma16(d_tmem, a_hi, b_2, idesc_n, 1);
                                        mma16(d_tmem, a_2, b_hi, idesc_n, 1);
                                    }
                                } else {             // A_hi x [W_hi ; W_lo] (N = 2NT) and A_lo x W_hi
                                    umma_f16(d_tmem, a_hi, b_hi, idesc_2n, accumulate);
                                    umma_f16(d_tmem, a_2, b_hi, idesc_n, 1);
                                }
                                accumulate = 1;
                            }
                        }
                        commit_both(&empty[bs]);              // frees the weight slot in BOTH CTAs
                    }
                    __syncwarp();
                    if (++bs == NS) { bs = 0; bph ^= 1u; }
                };
                for (int pi = 0; pi < total_patches; ++pi) {
                    const bool second = pi < p.kblocks2;
                    mma_wait(&afull[as_], aph);
                    tc_fence_after();
                    const uint32_t sa = a_ring + as_ * (Cfg::kAStage >> 4);
                    if (!second && p.taps == 9) {
#pragma unroll
                        for (int tap = 0; tap < 9; ++tap)
                            issue(sa + ((tap / 3) * kPatchW + tap % 3) * 8, kHiPatch, true);
                    } else {
                        issue(sa, kHiPlain, !second);
                    }
                    if (leader) commit_own(&aempty[as_]);
                    __syncwarp();
                    if (++as_ == NA) { as_ = 0; aph ^= 1u; }
                }
                if (leader) commit_own(&tfull[acc]);
                __syncwarp();
            }
        }
    } else if (warp < 4 + kFBuilderWarps) {
        // ---------------------------------------------------------------------- patch builders
        asm volatile("setmaxnreg.inc.sync.aligned.u32 144;");
        const int wtid = static_cast<int>(threadIdx.x) - 128;
        uint32_t jchunk = static_cast<uint32_t>(wtid & 7);
        uint32_t r0 = static_cast<uint32_t>(wtid >> 3);
        uint32_t a_ring = smem_u32(stage_base);
        // opaque to ptxas: it otherwise recomputes all three from %tid / the shared-memory window per patch ROW
        asm volatile("" : "+r"(jchunk), "+r"(r0), "+r"(a_ring));
        const bool skip = (p.debug & 2) != 0;
        const float a8_lo = FP8 ? p.a8_lo : 0.f;         // a8_hi == 1 (the dispatch sends other prescales elsewhere)

        // generator of this CTA's patch sequence: tiles in schedule order, per tile the shortcut K-blocks then the
        // main ones (the order the MMA issuer consumes them in)
        int g_item = cluster_id, g_pi = 0;
        int g_w0 = 0, g_h0 = 0, g_b0 = 0;
        auto locate = [&]() {
            int r = 2 * (g_item / p.tiles_n) + static_cast<int>(rank);
            const int wt = r % p.tiles_w; r /= p.tiles_w;
            const int ht = r % p.tiles_h; r /= p.tiles_h;
            g_w0 = wt << 3; g_h0 = ht << 4; g_b0 = r;
        };
        bool g_valid = g_item < p.total_items;
        if (g_valid) locate();
        auto next_plan = [&](PatchPlan& d) {            // plan of (g_item, g_pi), then advance
            const bool second = g_pi < p.kblocks2;
            const bool halo3 = !second && p.taps == 9;
            const int kb = second ? g_pi : g_pi - p.kblocks2;
            const int c = kb * 64 + static_cast<int>(jchunk) * 8;
            const float* x0 = second ? p.gx0 : p.fx0;
            const float* x1 = second ? p.gx1 : p.fx1;
            const int C0 = second ? p.gC0 : p.fC0, C1 = second ? p.gC1 : p.fC1;
            const float* src; int cs, cl;
            if (c < C0) { src = x0; cs = C0; cl = c; } else { src = x1; cs = C1; cl = c - C0; }
            const int PW = halo3 ? kPatchW : 8, kdy = halo3 ? 3 : 4;
            d.krows = halo3 ? 30u : 32u;
            d.niter = r0 < d.krows ? (halo3 ? 6u : 4u) : 0u;
            const int py0 = static_cast<int>(r0) / PW, px0 = static_cast<int>(r0) - py0 * PW;
            const int w = g_w0 - (halo3 ? 1 : 0) + px0;
            const int h0 = g_h0 - (halo3 ? 1 : 0) + py0;
            const bool col_ok = d.niter != 0u && g_b0 < p.B && w >= 0 && w < p.W;
            d.src = src + (((static_cast<long long>(g_b0) * p.H + h0) * p.W + w) * cs + cl);
            d.step = static_cast<uint32_t>(kdy * p.W * cs);
            uint32_t inb = 0;
#pragma unroll
            for (int u = 0; u < 6; ++u) {
                const int h = h0 + kdy * u;
                if (col_ok && u < static_cast<int>(d.niter) && h >= 0 && h < p.H) inb |= 1u << u;
            }
            d.inb = skip ? 0u : inb;
            d.so = static_cast<uint32_t>((g_b0 < p.B ? g_b0 : 0) * (C0 + C1) + c);
            d.second = second;
            d.mode = second ? 0 : (p.fact ? 2 : 1);
            if (++g_pi == total_patches) {
                g_pi = 0;
                g_item += num_clusters;
                g_valid = g_item < p.total_items;
                if (g_valid) locate();
            }
        };

        // two register sets: while one patch is converted out of set X, the loads of the following patch are in
        // flight into set Y, and vice versa (see load_rows / touch_rows above for the ordering and why)
        float4 vx[6][2], vy[6][2];
        PatchPlan px, py;
        int as_ = 0;
        uint32_t aph = 0;
        const float negzero = -(p.acc_scale * 0.0f);
        auto wait_slot = [&]() { role_wait(&aempty[as_], aph ^ 1u); };
        auto build = [&](const float4 (&v)[6][2], const PatchPlan& cur) {
            uint32_t slot = a_ring + static_cast<uint32_t>(as_) * Cfg::kAStage;
            asm volatile("" : "+r"(slot));              // once per patch, not per row
            const bool halo3 = cur.krows == 30u;        // shortcut patches are never halo patches
            if (FP8) {
                // SiLU known at compile time in the row bodies of the 2-unit mode (conv_builders.cuh)
                constexpr uint32_t PS = kPatchPlane;
                const bool silu = cur.mode == 2;
                if (cur.second) convert_rows_g<false, 32, 4, false, 6, 0>(v, cur, slot, r0, jchunk, 0.f, PS, 0u);
                else if (halo3 && silu) convert_rows_g<true, 30, 6, false, 6, 1>(v, cur, slot, r0, jchunk, a8_lo, PS, 0u);
                else if (halo3) convert_rows_g<true, 30, 6, false, 6, 0>(v, cur, slot, r0, jchunk, a8_lo, PS, 0u);
                else if (silu) convert_rows_g<true, 32, 4, false, 6, 1>(v, cur, slot, r0, jchunk, a8_lo, PS, 0u);
                else convert_rows_g<true, 32, 4, false, 6, 0>(v, cur, slot, r0, jchunk, a8_lo, PS, 0u);
            } else {
                if (halo3) convert_rows<false, true>(v, cur, slot, r0, jchunk, 0.f);
                else convert_rows<false, false>(v, cur, slot, r0, jchunk, 0.f);
            }
            // each builder warp publishes its own share (afull counts the builder warps)
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> async proxy (tensor core)
            __syncwarp();
            if (lane == 0) {
                if (TWO && rank != 0) mbar_arrive_cluster(&afull[as_], 0);     // the leader's MMA warp waits
                else mbar_arrive(&afull[as_]);
            }
            if (++as_ == NA) { as_ = 0; aph ^= 1u; }
        };
        bool has_x = g_valid, has_y = false;
        if (has_x) {
            next_plan(px);
            load_rows(vx, px);
            has_y = g_valid;
            if (has_y) next_plan(py);
        }
        while (has_x) {
            touch_rows(vx, px, p.fsc, p.fsh, negzero);
            wait_slot();
            if (has_y) load_rows(vy, py);
            build(vx, px);
            has_x = g_valid;
            if (has_x) next_plan(px);
            if (!has_y) break;
            touch_rows(vy, py, p.fsc, p.fsh, negzero);
            wait_slot();
            if (has_x) load_rows(vx, px);
            build(vy, py);
            has_y = g_valid;
            if (has_y) next_plan(py);
        }
    } else {
        // ---------------------------------------------------------------------- epilogue (4 warps)
        asm volatile("setmaxnreg.inc.sync.aligned.u32 184;");
        const int wq = warp & 3;               // TMEM lane quarter == warp_id % 4
        const float crel = FP8 ? p.corr_rel : 1.0f;
        constexpr int kChunks = NT / 32;       // 32-column chunks of the tile, all handled by this warp
        // GroupNorm partial sums of the tile sequence of one (batch entry, channel tile): [chunk][channel quad q] x
        // (sum, sum of squares) x 4 channels, kept in this warp's shared-memory scratch by lanes 0-7 (they used to be
        // 32 registers per thread, which the residual double buffer below needs) and flushed with fp64 atomics when
        // the batch entry / channel tile changes
        const uint32_t sstat = smem_u32(reinterpret_cast<uint8_t*>(staging) + kFEpiWarps * 32 * 32 * 4 +
                                        wq * (kChunks * 8 * 32)) + (lane & 7) * 32;
        // (bias + FiLM) * scale of this lane's 4 channels per chunk: loaded in ONE burst at the top of a tile into a
        // lane-private shared-memory slot (per chunk it was a global load with its latency exposed four times a tile)
        const uint32_t sbias = smem_u32(reinterpret_cast<uint8_t*>(staging) + kFEpiWarps * 32 * 32 * 4 + Cfg::kStatBytes +
                                        wq * (kChunks * 32 * 16)) + lane * 16;
        int run_b = -1, run_n0 = -1;
        auto zero_stats = [&]() {
            if (lane < 8) {
#pragma unroll
                for (int c = 0; c < kChunks; ++c) {
                    sts128(sstat + c * 256, 0u, 0u, 0u, 0u);
                    sts128(sstat + c * 256 + 16, 0u, 0u, 0u, 0u);
                }
            }
        };
        auto flush_stats = [&]() {
            if (p.stats == nullptr || run_b < 0) return;           // warp-uniform
            if (run_b >= p.B || lane >= 8) return;
#pragma unroll
            for (int c = 0; c < kChunks; ++c) {
                const int n = run_n0 + c * 32 + lane * 4;
                if (n < p.cout_store) {
                    const float4 r1 = lds128(sstat + c * 256), r2 = lds128(sstat + c * 256 + 16);
                    double* st = p.stats + (static_cast<size_t>(run_b) * p.cout_store + n) * 2;
                    atomicAdd(st + 0, (double)r1.x); atomicAdd(st + 1, (double)r2.x);
                    atomicAdd(st + 2, (double)r1.y); atomicAdd(st + 3, (double)r2.y);
                    atomicAdd(st + 4, (double)r1.z); atomicAdd(st + 5, (double)r2.z);
                    atomicAdd(st + 6, (double)r1.w); atomicAdd(st + 7, (double)r2.w);
                }
            }
        };
        const uint32_t stg_s = smem_u32(staging + wq * (32 * 32));
        const int q = lane & 7, rg = lane >> 3;
        const uint32_t st_base = stg_s + lane * 128 + ((lane & 7) << 4);     // chunk j at ^ (j << 4)
        const uint32_t ld_base = stg_s + rg * 128 + ((q ^ rg) << 4);         // row i at + i*512, ^ ((i&1) << 6)
        const float as2 = p.acc_scale * p.scale;
        const bool store = !(p.debug & 4);
        float4 res[2][8];          // residual rows, double-buffered over the chunks (and across tiles, see below)
        bool have0 = false;        // res[0] already holds chunk 0 of the tile about to be processed
        int it = 0;
        for (int item = cluster_id; item < p.total_items; item += num_clusters, ++it) {
            const int as = it & 1;
            const int nt = item % p.tiles_n;
            int r = 2 * (item / p.tiles_n) + static_cast<int>(rank);      // this CTA's M tile of the pair
            const int wt = r % p.tiles_w; r /= p.tiles_w;
            const int ht = r % p.tiles_h; r /= p.tiles_h;
            // r >= B only for the odd tile out of the last pair: its patches were zero and every store is masked
            const int w0 = wt << 3, h0 = ht << 4, b0 = r;
            const int n0 = nt * NT;
            const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + as * 2 * NT;
            if (p.stats != nullptr && (b0 != run_b || n0 != run_n0)) {
                flush_stats();
                run_b = b0; run_n0 = n0;
                zero_stats();
            }
            const bool whole = b0 < p.B && h0 + 16 <= p.H && w0 + 8 <= p.W && n0 + NT <= p.cout_store;
            const float* next_res0 = nullptr;      // chunk 0 of the NEXT tile, if that tile lies whole inside the map
            if (p.residual != nullptr) {
                // pull the NEXT tile's residual rows into L2 now (no registers, a whole tile period of lead time)
                const int nitem = item + num_clusters;
                if (nitem < p.total_items) {
                    int r2 = 2 * (nitem / p.tiles_n) + static_cast<int>(rank);
                    const int wt2 = r2 % p.tiles_w; r2 /= p.tiles_w;
                    const int ht2 = r2 % p.tiles_h; r2 /= p.tiles_h;
                    const int hr2 = (ht2 << 4) + wq * 4, wc2 = (wt2 << 3) + rg, n02 = (nitem % p.tiles_n) * NT;
                    if (r2 < p.B && (ht2 << 4) + 16 <= p.H && (wt2 << 3) + 8 <= p.W && n02 + NT <= p.cout_store)
                        next_res0 = p.residual + ((static_cast<size_t>(r2) * p.H + hr2) * p.W + wc2) * p.cout_store
                                    + n02 + q * 4;
                    if (r2 < p.B && q < kChunks) {        // lane q of each row group takes the row's q-th 128-byte line
                        const float* base = p.residual + ((static_cast<size_t>(r2) * p.H + hr2) * p.W + wc2) * p.cout_store
                                            + n02 + q * 32;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            if (hr2 + (i >> 1) < p.H && wc2 + 4 * (i & 1) < p.W && n02 + q * 32 < p.cout_store)
                                asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (i >> 1) * static_cast<size_t>(p.W) * p.cout_store
                                                                              + (i & 1) * 4 * static_cast<size_t>(p.cout_store)));
                        }
                    }
                }
            }
            // row i of this thread: pixel (h0 + wq*4 + (i >> 1), w0 + rg + 4*(i & 1)), 4 channels from n
            const int hrow = h0 + wq * 4, wcol = w0 + rg;
            const uint32_t C = static_cast<uint32_t>(p.cout_store);
            const size_t e0 = ((static_cast<size_t>(b0 < p.B ? b0 : 0) * p.H + hrow) * p.W + wcol) * C + n0 + q * 4;
            float* const out0 = p.out + e0;
            const float* const res0 = p.residual != nullptr ? p.residual + e0 : nullptr;
            const uint32_t d_row = static_cast<uint32_t>(p.W) * C;      // i -> i + 2: next image row
            const uint32_t d_half = 4u * C;                            // odd i: 4 pixels to the right
            // rows of this thread that exist (partial tiles at the map's edge; all 8 when `whole`)
            uint32_t rowmask = 0xFFu;
            if (!whole) {
                rowmask = 0;
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (b0 < p.B && hrow + (i >> 1) < p.H && wcol + 4 * (i & 1) < p.W) rowmask |= 1u << i;
            }
            bool waited = false;
            // residual rows: chunk c + 1's loads are issued while chunk c is
            // processed (one chunk = ~1.5 k clk of distance; issued at their point of use they made the epilogue the
            // kernel's critical path: 1.16 ms for MMAs + epilogue alone against 0.94 ms without a residual).  Same
            // scoreboard trap as in the builders: the current buffer is "touched" (scaled in place) BEFORE the next
            // burst is issued, with a warp barrier in between as a scheduling fence.
            auto load_res = [&](int c, float4 (&r)[8]) {
                const bool n_ok_c = whole || n0 + c * 32 + q * 4 < p.cout_store;
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    r[i] = (n_ok_c && ((rowmask >> i) & 1u))
                               ? ldg_stream(res0 + c * 32 + (i >> 1) * d_row + (i & 1) * d_half)
                               : make_float4(0.f, 0.f, 0.f, 0.f);
            };
            if (res0 != nullptr && !have0) load_res(0, res[0]);
            have0 = false;
            {
                float4 bz_all[kChunks];
#pragma unroll
                for (int c = 0; c < kChunks; ++c) {
                    const int n = n0 + c * 32 + q * 4;
                    float4 bz = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (whole || n < p.cout_store) {
                        if (p.bias != nullptr) bz = __ldg(reinterpret_cast<const float4*>(p.bias + n));
                        if (p.film != nullptr && b0 < p.B) {
                            const float4 f = __ldg(reinterpret_cast<const float4*>(
                                p.film + static_cast<size_t>(b0) * p.film_stride + n));
                            bz.x += f.x; bz.y += f.y; bz.z += f.z; bz.w += f.w;
                        }
                    }
                    bz_all[c] = bz;
                }
#pragma unroll
                for (int c = 0; c < kChunks; ++c)
                    sts128(sbias + c * 512, __float_as_uint(bz_all[c].x * p.scale), __float_as_uint(bz_all[c].y * p.scale),
                           __float_as_uint(bz_all[c].z * p.scale), __float_as_uint(bz_all[c].w * p.scale));
            }
#pragma unroll
            for (int c = 0; c < kChunks; ++c) {
                const int n = n0 + c * 32 + q * 4;
                const bool n_ok = whole || n < p.cout_store;
                if (res0 != nullptr) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        res[c & 1][i].x *= p.scale; res[c & 1][i].y *= p.scale;
                        res[c & 1][i].z *= p.scale; res[c & 1][i].w *= p.scale;
                    }
                    __syncwarp();
                    if (c + 1 < kChunks) {
                        load_res(c + 1, res[(c + 1) & 1]);
                    } else if (next_res0 != nullptr && (kChunks & 1) == 0) {
                        // last chunk: res[0] is free — chunk 0 of the next tile goes out now, so that no tile starts
                        // with an exposed memory round trip
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            res[0][i] = ldg_stream(next_res0 + (i >> 1) * d_row + (i & 1) * d_half);
                        have0 = true;
                    }
                }
                const float4 bz = lds128(sbias + c * 512);
                if (!waited) {
                    role_wait(&tfull[as], (it >> 1) & 1);
                    tc_fence_after();
                    waited = true;
                }
                uint32_t v[32];
                tmem_ld_32x32(t_addr + c * 32, v);
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {   // add the correction / hi*lo half, 16 columns at a time
                    uint32_t u[16];
                    tmem_ld_32x16(t_addr + NT + c * 32 + hh * 16, u);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        v[hh * 16 + j] = __float_as_uint(fmaf(__uint_as_float(u[j]), crel, __uint_as_float(v[hh * 16 + j])));
                }
                if (c == kChunks - 1) {   // TMEM fully drained by this warp: hand the buffer back early
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if (TWO && rank != 0) mbar_arrive_cluster(&tempty[as], 0);
                        else mbar_arrive(&tempty[as]);
                    }
                }
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    sts128(st_base ^ (j << 4), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                __syncwarp();
                float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float4 a = lds128((ld_base ^ ((i & 1) << 6)) + i * 512);
                    if (n_ok && ((rowmask >> i) & 1u)) {
                        a.x = fmaf(a.x, as2, bz.x); a.y = fmaf(a.y, as2, bz.y);
                        a.z = fmaf(a.z, as2, bz.z); a.w = fmaf(a.w, as2, bz.w);
                        if (res0 != nullptr) {
                            a.x += res[c & 1][i].x; a.y += res[c & 1][i].y;
                            a.z += res[c & 1][i].z; a.w += res[c & 1][i].w;
                        }
                        if (store)
                            *reinterpret_cast<float4*>(out0 + c * 32 + (i >> 1) * d_row + (i & 1) * d_half) = a;
                        s1.x += a.x; s1.y += a.y; s1.z += a.z; s1.w += a.w;
                        s2.x = fmaf(a.x, a.x, s2.x); s2.y = fmaf(a.y, a.y, s2.y);
                        s2.z = fmaf(a.z, a.z, s2.z); s2.w = fmaf(a.w, a.w, s2.w);
                    }
                }
                if (p.stats != nullptr) {      // rows of the 4 lane groups -> lanes 0..7 -> the running sums
#pragma unroll
                    for (int o = 8; o <= 16; o <<= 1) {
                        s1.x += __shfl_xor_sync(0xffffffffu, s1.x, o); s1.y += __shfl_xor_sync(0xffffffffu, s1.y, o);
                        s1.z += __shfl_xor_sync(0xffffffffu, s1.z, o); s1.w += __shfl_xor_sync(0xffffffffu, s1.w, o);
                        s2.x += __shfl_xor_sync(0xffffffffu, s2.x, o); s2.y += __shfl_xor_sync(0xffffffffu, s2.y, o);
                        s2.z += __shfl_xor_sync(0xffffffffu, s2.z, o); s2.w += __shfl_xor_sync(0xffffffffu, s2.w, o);
                    }
                    if (lane < 8) {
                        const float4 r1 = lds128(sstat + c * 256), r2 = lds128(sstat + c * 256 + 16);
                        sts128(sstat + c * 256, __float_as_uint(r1.x + s1.x), __float_as_uint(r1.y + s1.y),
                               __float_as_uint(r1.z + s1.z), __float_as_uint(r1.w + s1.w));
                        sts128(sstat + c * 256 + 16, __float_as_uint(r2.x + s2.x), __float_as_uint(r2.y + s2.y),
                               __float_as_uint(r2.z + s2.z), __float_as_uint(r2.w + s2.w));
                    }
                }
                __syncwarp();
            }
        }
        flush_stats();
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();          // the peer may still multicast into / arrive on this CTA's shared memory
    if (warp == 2) {
        if constexpr (TWO) tmem_dealloc_2cta<Cfg::kTmemCols>(tmem_base);
        else tmem_dealloc<Cfg::kTmemCols>(tmem_base);
    }
}

template <int NT, bool FP8, bool TWO>
static int launch_fused(const ConvMaps& m, const ConvParams& p, cudaStream_t stream) {
    constexpr int kSmem = FusedCfg<NT, TWO>::kSmemBytes;
    static PerDeviceAttr attr;
    const cudaError_t attr_err = set_max_smem_once(attr, conv_fused_kernel<NT, FP8, TWO>, kSmem);
    if (attr_err != cudaSuccess) {
        set_error("cudaFuncSetAttribute(conv_fused_kernel<%d,%d>): %s", NT, (int)FP8, cudaGetErrorString(attr_err));
        return DSEP_ERR_CUDA;
    }
    const int max_clusters = conv_num_sms() / 2;
    const int grid = 2 * (p.total_items < max_clusters ? p.total_items : max_clusters);
    conv_fused_kernel<NT, FP8, TWO><<<grid, kFThreads, kSmem, stream>>>(m.w_hi, m.w_lo, m.w2_hi, m.w2_lo, p);
    return check_launch("conv_fused_kernel");
}

int launch_conv_fused(const ConvMaps& m, const ConvParams& p, int NT, cudaStream_t stream) {
    // DSEP_CONV_PAIR=1: the cta_group::2 variant (experimental: measured slower than 1-CTA MMAs with multicast
    // weights — 1.33 vs 1.22 ms on the level-0 conv — and not the default)
    static const int pair_env = getenv("DSEP_CONV_PAIR") ? atoi(getenv("DSEP_CONV_PAIR")) : 0;
    if (p.passes == 2 && pair_env != 0)
        return NT == 64 ? launch_fused<64, true, true>(m, p, stream) : launch_fused<128, true, true>(m, p, stream);
    if (p.passes == 2)
        return NT == 64 ? launch_fused<64, true, false>(m, p, stream) : launch_fused<128, true, false>(m, p, stream);
    return NT == 64 ? launch_fused<64, false, false>(m, p, stream) : launch_fused<128, false, false>(m, p, stream);
}

}  // namespace dsep
