// Implicit-GEMM 3x3 / 1x1 convolution on tcgen05 tensor cores (sm_100a).
//
//   out[b,h,w,n] = scale * ( acc_scale * ( sum_{tap,c} A[b,h+dy,w+dx,c] * Wt[tap,n,c]
//                                          + sum_c A2[b,h,w,c] * W2[n,c] )
//                            + bias[n] + film[b,n] + residual[b,h,w,n] )
//   stats[b,n] += (sum, sum of squares) of out over the pixels of the tile        (optional)
//
// Replaces the cuDNN convolutions behind ddpm_conv3x3 / ddpm_conv1x1 / NIN of the reference
// (models/ncsnpp_utils/layers.py:112-156, 678-689) together with the Dense_0 bias add, the 1x1
// shortcut Conv_2 and the (x + h)/sqrt(2) residual of ResnetBlockBigGANpp.forward
// (layerspp.py:311-323), and produces the statistics the NEXT GroupNorm needs, so that tensor is
// never re-read just to be reduced.
//
// Design (B200-first):
//   * activations are channels-last, so the A tile of one filter tap (128 pixels x 64 channels)
//     is ONE 4-D TMA box [64 ch, tw, th, tb] at coordinates shifted by (dx, dy); out-of-image
//     taps are zero-filled by TMA — no im2col, no padding copies;
//   * operands are (hi, lo) fp16 planes (weights pre-scaled by a power of two, undone by
//     acc_scale).  Per K=16 step two tcgen05.mma: A_hi x [W_hi ; W_lo] (the two weight planes
//     stacked along N, N = 2 NT, into TMEM columns [0, 2NT)) and A_lo x W_hi (N = NT, columns
//     [0, NT)); the epilogue adds the halves.  Same MAC count as three passes with 17 % less
//     shared-memory operand traffic, and the small hi*lo terms accumulate apart from the large
//     ones.  passes = 1 issues A_hi x W_hi only (11-bit operands: TF32-grade, what cuDNN runs
//     for the reference by default on a GPU);
//   * the 1x1 shortcut convolution is extra K-blocks from a second operand (A2, W2) accumulated
//     into the same TMEM tile: its output is never written to or re-read from HBM;
//   * persistent CTAs (one per SM), warp-specialised: warp 0 TMA producer, warp 1 MMA issuer (the
//     whole warp runs convergently so descriptors live in uniform registers; one elected lane
//     issues), warp 2 TMEM allocator, warps 4-11 workers: patch builders + epilogue (two warps per
//     TMEM lane quarter, splitting the columns).  Halo kernels move registers from the control
//     warpgroup to the worker warpgroups with setmaxnreg (56 / 224).  smem ring of K-blocks with
//     mbarrier full/empty pairs; the TMEM accumulator is double-buffered so the epilogue of tile i
//     overlaps the MMAs of tile i+1;
//   * epilogue: residual prefetched into registers, tcgen05.ld -> XOR-swizzled per-warp smem
//     transpose (conflict-free both ways) -> bias/FiLM/residual/scale fused -> 128-byte
//     coalesced fp32 stores; per-channel (sum, sum^2) reduced in registers + 2 shuffles and
//     accumulated with fp64 atomics.
#include "conv_tc.cuh"

namespace dsep {

template <int NT>
struct ConvCfg {
    static constexpr int kBBytes = NT * kBK * 2;
    static constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;
    static constexpr int kStagingBytes = NT >= 64 ? kEpiWarps * 32 * 32 * 4 : 0;
    static constexpr int kAvail = 232448 - 1024 - 512 - kStagingBytes;
    static constexpr int kStagesMax = kAvail / kStageBytes;
    static constexpr int kStages = kStagesMax > 8 ? 8 : kStagesMax;
    static constexpr int kTmemCols = 4 * NT < 32 ? 32 : 4 * NT;   // 2 stages x (hi*hi+lo*hi | hi*lo)
    static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + 512 + 1024;
};


template <int NT>
struct HaloCfg {
    static constexpr int kBBytes = NT * 128;                   // one weight plane of a stage (64 ch)
    static constexpr int kAStage = 2 * kPatchPlane;
    static constexpr int kBStage = 2 * kBBytes;
    static constexpr int kStagingBytes = kEpiWarps * 32 * 32 * 4;
    static constexpr int kAvail = 232448 - 1024 - 512 - kStagingBytes - kHaloAStages * kAStage;
    static constexpr int kBStagesMax = kAvail / kBStage;
    static constexpr int kBStages = kBStagesMax > 8 ? 8 : kBStagesMax;
    static constexpr int kSmemBytes = kHaloAStages * kAStage + kBStages * kBStage + kStagingBytes + 512 + 1024;
};


// One builder thread's share of a patch.  Item = (patch row r, 8-channel chunk j); thread wtid owns chunk
// j = wtid & 7 of rows r0 + kRows u (r0 = wtid >> 3).  kRows is a multiple of the patch width (32 for 8-wide
// patches; 30 for the 10-wide halo patch, the last 16 builder threads idle), so all of a thread's rows sit in
// ONE patch column, kDy image rows apart: one address and one bounds test per thread, then a constant
// stride — no per-item division, and no partial last iteration (16 x 8 = 4 x 32, 18 x 10 = 6 x 30).
template <int PW, int PH>
struct PatchRegs {
    static constexpr int kRows = (kBuilders / 8) / PW * PW;
    static constexpr int kDy = kRows / PW;
    static constexpr int kIter = (PW * PH) / kRows;
    static_assert(kIter * kRows == PW * PH, "patch rows must split evenly over the builder threads");
    float4 v[kIter][2];
    float k_sc[8], k_sh[8];
    uint32_t inb;          // bit u: item u lies inside the image (loaded)
    bool active;           // this thread has rows in the patch
    uint32_t off0;         // byte offset of row r0 inside a patch plane
    uint32_t jchunk;       // this thread's 16-byte chunk of the row (before the swizzle XOR)
};

// phase 1: issue the global loads (activation rows + the GroupNorm affine of this thread's 8 channels)
template <int PW, int PH>
__device__ __forceinline__ void patch_load(PatchRegs<PW, PH>& R, const float* x0, int C0, const float* x1, int C1,
                                           int kb, const float* sc, const float* sh, int b, int h_org, int w_org,
                                           int B, int H, int W, int wtid) {
    using PR = PatchRegs<PW, PH>;
    const int j = wtid & 7;
    const int c = kb * 64 + j * 8;                          // first of this thread's 8 channels (concatenated)
    const float* src;
    int cs, cl;
    if (c < C0) { src = x0; cs = C0; cl = c; } else { src = x1; cs = C1; cl = c - C0; }
    if (sc != nullptr) {
        const size_t so = static_cast<size_t>(b < B ? b : 0) * (C0 + C1) + c;
        const float4 a0 = __ldg(reinterpret_cast<const float4*>(sc + so));
        const float4 a1 = __ldg(reinterpret_cast<const float4*>(sc + so + 4));
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(sh + so));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(sh + so + 4));
        R.k_sc[0] = a0.x; R.k_sc[1] = a0.y; R.k_sc[2] = a0.z; R.k_sc[3] = a0.w;
        R.k_sc[4] = a1.x; R.k_sc[5] = a1.y; R.k_sc[6] = a1.z; R.k_sc[7] = a1.w;
        R.k_sh[0] = b0.x; R.k_sh[1] = b0.y; R.k_sh[2] = b0.z; R.k_sh[3] = b0.w;
        R.k_sh[4] = b1.x; R.k_sh[5] = b1.y; R.k_sh[6] = b1.z; R.k_sh[7] = b1.w;
    }
    const int r0 = wtid >> 3;
    R.active = r0 < PR::kRows;
    const int py0 = r0 / PW, px0 = r0 - py0 * PW;
    const int w = w_org + px0;
    const int h0 = h_org + py0;
    const bool col_ok = R.active && b < B && w >= 0 && w < W;
    R.jchunk = static_cast<uint32_t>(j);
    R.off0 = static_cast<uint32_t>(r0) * 128u;              // + u * kRows * 128; the swizzle phase (r & 7) varies with u
    const long long e0 = ((static_cast<long long>(b) * H + h0) * W + w) * cs + cl;
    const long long step = static_cast<long long>(PR::kDy) * W * cs;
    R.inb = 0;
#pragma unroll
    for (int u = 0; u < PR::kIter; ++u) {                   // all loads first: ~12 x 16 B in flight per thread
        const int h = h0 + PR::kDy * u;
        if (col_ok && h >= 0 && h < H) {
            const float* q = src + (e0 + u * step);
            R.v[u][0] = __ldg(reinterpret_cast<const float4*>(q));
            R.v[u][1] = __ldg(reinterpret_cast<const float4*>(q + 4));
            R.inb |= 1u << u;
        } else {
            R.v[u][0] = R.v[u][1] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
}

// phase 2: y = act(x * sc + sh) (zero outside the image: the conv pads the ACTIVATED tensor), fp32 -> (hi, lo)
// fp16 split, 128-byte-swizzled K-major rows.  dst_hi / dst_lo are shared-window addresses of the two planes.

template <int PW, int PH>
__device__ __forceinline__ void patch_store(const PatchRegs<PW, PH>& R, uint32_t dst_hi, uint32_t dst_lo, bool want_lo,
                                            bool affine, int act, float a8_hi = 0.f, float a8_lo = 0.f) {
    using PR = PatchRegs<PW, PH>;
    if (!R.active) return;
#pragma unroll
    for (int u = 0; u < PR::kIter; ++u) {
        const uint32_t r = (R.off0 >> 7) + static_cast<uint32_t>(u * PR::kRows);
        const uint32_t off = r * 128u + ((R.jchunk ^ (r & 7u)) << 4);
        uint32_t hi[4] = {0u, 0u, 0u, 0u}, lo[4] = {0u, 0u, 0u, 0u};
        if ((R.inb >> u) & 1u) {
            float y[8] = {R.v[u][0].x, R.v[u][0].y, R.v[u][0].z, R.v[u][0].w,
                          R.v[u][1].x, R.v[u][1].y, R.v[u][1].z, R.v[u][1].w};
            if (affine) {
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    float t = fmaf(y[e], R.k_sc[e], R.k_sh[e]);
                    if (act) {   // SiLU = t / (1 + 2^(-t log2 e)): ex2.approx.ftz + rcp.approx.ftz, no range fix-ups
                        float ex, rc;
                        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(t * -1.4426950408889634f));
                        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(1.0f + ex));
                        t *= rc;
                    }
                    y[e] = t;
                }
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) split2_f16(y[2 * e], y[2 * e + 1], hi[e], lo[e]);
#if DSEP_FP8_CORR
            if (a8_hi != 0.f) {      // second plane = e4m3 corrections instead of the fp16 lo plane
                float h[8], l[8];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&hi[e]));
                    h[2 * e] = b.x; h[2 * e + 1] = b.y;
                    l[2 * e] = y[2 * e] - b.x; l[2 * e + 1] = y[2 * e + 1] - b.y;
                }
                lo[0] = e4m3x4(l[0] * a8_lo, l[1] * a8_lo, l[2] * a8_lo, l[3] * a8_lo);
                lo[1] = e4m3x4(l[4] * a8_lo, l[5] * a8_lo, l[6] * a8_lo, l[7] * a8_lo);
                lo[2] = e4m3x4(h[0] * a8_hi, h[1] * a8_hi, h[2] * a8_hi, h[3] * a8_hi);
                lo[3] = e4m3x4(h[4] * a8_hi, h[5] * a8_hi, h[6] * a8_hi, h[7] * a8_hi);
            }
#endif
        }
        sts128(dst_hi + off, hi[0], hi[1], hi[2], hi[3]);
        if (want_lo) sts128(dst_lo + off, lo[0], lo[1], lo[2], lo[3]);
    }
}

// TWO (halo mode only): the pair issues ONE tcgen05.mma.cta_group::2 per step from the leader CTA
// (M = 256: 128 pixels from each CTA; the B rows are split between the two CTAs' shared memory), which
// halves the B-operand shared-memory reads and ingest per SM — the resource the 1-CTA kernel saturates.
template <int NT, bool HALO, bool TWO>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
               const __grid_constant__ CUtensorMap tm_w_hi, const __grid_constant__ CUtensorMap tm_w_lo,
               const __grid_constant__ CUtensorMap tm_a2_hi, const __grid_constant__ CUtensorMap tm_a2_lo,
               const __grid_constant__ CUtensorMap tm_w2_hi, const __grid_constant__ CUtensorMap tm_w2_lo,
               const ConvParams p) {
    using Cfg = ConvCfg<NT>;
    using HCfg = HaloCfg<NT>;
    // per-tap mode: one ring of NS stages {A_hi, A_lo, W_hi, W_lo}; halo mode: a ring of NA patches
    // {A_hi, A_lo} (barriers afull/aempty) and a ring of NS weight stages {W_hi, W_lo} (full/empty)
    constexpr int NS = HALO ? HCfg::kBStages : Cfg::kStages;
    constexpr int NA = kHaloAStages;
    constexpr int kRingBytes = HALO ? NA * HCfg::kAStage + NS * HCfg::kBStage : NS * Cfg::kStageBytes;
    constexpr int kStagingBytes = HALO ? HCfg::kStagingBytes : Cfg::kStagingBytes;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* stage_base = smem;
    float* staging = reinterpret_cast<float*>(smem + kRingBytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kRingBytes + kStagingBytes);
    uint64_t* full = bars;             // [NS]   TMA -> MMA
    uint64_t* empty = bars + NS;       // [NS]   MMA -> TMA
    uint64_t* tfull = bars + 2 * NS;   // [2]    MMA -> epilogue
    uint64_t* tempty = tfull + 2;      // [2]    epilogue -> MMA
    uint64_t* afull = tempty + 2;      // [NA]   halo patches: TMA -> MMA
    uint64_t* aempty = afull + NA;     // [NA]   MMA -> TMA
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(aempty + NA);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // warp-uniform for the compiler too
    const int lane = threadIdx.x & 31;
#if DSEP_FP8_CORR
    const bool three = p.passes != 1;        // two operand planes per stage (fp16 lo, or the e4m3 correction plane)
    const bool fp8c = p.passes == 2;         // main K-blocks: hi*hi in fp16 + both corrections in one e4m3 product
#else
    const bool three = p.passes == 3;
#endif
    // CTA pair: the two CTAs of a cluster work on M-adjacent tiles of the same channel tile, each
    // fetches half of every weight tile and multicasts it to both (halves the weight traffic)
    const uint32_t rank = cluster_ctarank();
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_a_hi);
        tma_prefetch_desc(&tm_w_hi);
        if (three) { tma_prefetch_desc(&tm_a_lo); tma_prefetch_desc(&tm_w_lo); }
        if (p.kblocks2 > 0) { tma_prefetch_desc(&tm_a2_hi); tma_prefetch_desc(&tm_w2_hi); }
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < NS; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], TWO ? 1 : 2);    // 1-CTA MMAs: one tcgen05.commit from each CTA of the pair
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull[i], 1);
            // narrow tiles: only 4 warps drain TMEM; TWO: the leader collects both CTAs' epilogue warps
            mbar_init(&tempty[i], NT >= 64 ? (TWO ? 2 * kEpiWarps : kEpiWarps) : 4);
        }
        for (int i = 0; i < NA; ++i) {
            // built patches: one arrival per builder warp (of both CTAs under TWO); TMA-fed patches: the producer's
            mbar_init(&afull[i], p.fx0 != nullptr ? (TWO ? 2 : 1) * (kBuilders / 32) : 1);
            mbar_init(&aempty[i], 1);
        }
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

    const int tb_log2 = 7 - p.tw_log2 - p.th_log2;
    const int kiters = p.taps * p.kblocks + p.kblocks2;
    const uint32_t stage_tx = (three ? 2u : 1u) * static_cast<uint32_t>(kABytes + Cfg::kBBytes);

    // builds, in ring order (shortcut K-blocks, then main), every A patch of one tile; run by kBuilders threads
    auto build_tile_patches = [&](int item, int wtid, int& as_, uint32_t& aph) {
        if constexpr (HALO) {
            int r = 2 * (item / p.tiles_n) + static_cast<int>(rank);
            const int wt = r % p.tiles_w; r /= p.tiles_w;
            const int ht = r % p.tiles_h; r /= p.tiles_h;
            const int w0 = wt << 3, h0 = ht << 4, b0 = r;
            const int total = p.kblocks + p.kblocks2;
            const bool skip = (p.debug & 2) != 0;
            // each builder warp publishes its own share (afull counts the builder warps): no CTA-wide barrier,
            // so warps drift apart and one warp's load latency hides behind another's arithmetic
            auto publish = [&]() {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> async proxy (tensor core)
                __syncwarp();
                if (lane == 0) {
                    if (TWO && rank != 0) mbar_arrive_cluster(&afull[as_], 0);   // the leader's MMA warp waits
                    else mbar_arrive(&afull[as_]);
                }
                if (++as_ == kHaloAStages) { as_ = 0; aph ^= 1u; }
            };
            for (int pi = 0; pi < total; ++pi) {
                const bool second = pi < p.kblocks2;
                const uint32_t sa = smem_u32(stage_base + as_ * HaloCfg<NT>::kAStage);
                mbar_wait(&aempty[as_], aph ^ 1u);
                if (second || p.taps != 9) {
                    PatchRegs<8, 16> R;
                    if (!skip) {
                        if (second)
                            patch_load<8, 16>(R, p.gx0, p.gC0, p.gx1, p.gC1, pi, nullptr, nullptr, b0, h0, w0, p.B, p.H,
                                              p.W, wtid);
                        else
                            patch_load<8, 16>(R, p.fx0, p.fC0, p.fx1, p.fC1, pi - p.kblocks2, p.fsc, p.fsh, b0, h0, w0,
                                              p.B, p.H, p.W, wtid);
                    }
#if DSEP_FP8_CORR
                    if (!skip)
                        patch_store<8, 16>(R, sa, sa + kPatchPlane, three, !second && p.fsc != nullptr, p.fact,
                                           (fp8c && !second) ? p.a8_hi : 0.f, p.a8_lo);
#else
                    if (!skip) patch_store<8, 16>(R, sa, sa + kPatchPlane, three, !second && p.fsc != nullptr, p.fact);
#endif
                } else {
                    PatchRegs<kPatchW, kPatchH> R;
                    if (!skip)
                        patch_load<kPatchW, kPatchH>(R, p.fx0, p.fC0, p.fx1, p.fC1, pi - p.kblocks2, p.fsc, p.fsh, b0,
                                                     h0 - 1, w0 - 1, p.B, p.H, p.W, wtid);
#if DSEP_FP8_CORR
                    if (!skip)
                        patch_store<kPatchW, kPatchH>(R, sa, sa + kPatchPlane, three, p.fsc != nullptr, p.fact,
                                                      fp8c ? p.a8_hi : 0.f, p.a8_lo);
#else
                    if (!skip) patch_store<kPatchW, kPatchH>(R, sa, sa + kPatchPlane, three, p.fsc != nullptr, p.fact);
#endif
                }
                publish();
            }
        }
    };

    // halo kernels: the control warpgroup (TMA producer, MMA issuer, two idle warps) hands registers to the
    // two worker warpgroups, whose patch builders + epilogue otherwise spill at the 168-register launch bound
    // (0.9 GB of local-memory traffic per launch, measured): 128 x 56 + 256 x 224 = 64512 registers.  Each
    // setmaxnreg sits at the top of its role branch so that the code it dominates is allocated to that bound.
    if (HALO && warp < 4) {
      if constexpr (HALO) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
      if (warp == 0 && lane == 0) {
        // ------------------------------------------------------------------ TMA producer (halo mode)
        if constexpr (HALO) {
            const uint32_t planes = three ? 2u : 1u;
            int bs = 0, as_ = 0;
            uint32_t bph = 0, aph = 0;
            auto load_weights = [&](const CUtensorMap* whi, const CUtensorMap* wlo, int kcol, int wrow) {
                mbar_wait(&empty[bs], bph ^ 1u);
                uint8_t* sb = stage_base + NA * HCfg::kAStage + bs * HCfg::kBStage;
                if (p.debug & 2) {
                    if (!TWO || rank == 0) mbar_arrive(&full[bs]);
                    if (++bs == NS) { bs = 0; bph ^= 1u; }
                    return;
                }
                if constexpr (TWO) {
                    // B rows are split between the CTAs.  X (NT rows): W_hi in the leader, W_lo in the
                    // follower = the two halves of the stacked [W_hi ; W_lo]; Y (NT/2 rows): this CTA's
                    // half of W_hi for the A_lo x W_hi product.  Everything completes on the leader's barrier.
#if DSEP_FP8_CORR
                    if (fp8c) {
                        // each CTA holds its half of the rows of both weight planes (fp16 hi; e4m3 corrections, or
                        // the fp16 lo plane of a shortcut K-block): N = NT products only
                        if (rank == 0) mbar_arrive_expect_tx(&full[bs], 2u * HCfg::kBBytes);
                        const int wr = wrow + static_cast<int>(rank) * (NT / 2);
                        tma_load_2d_2sm(sb, whi, &full[bs], kcol, wr);
                        tma_load_2d_2sm(sb + HCfg::kBBytes, wlo, &full[bs], kcol, wr);
                        if (++bs == NS) { bs = 0; bph ^= 1u; }
                        return;
                    }
#endif
                    if (rank == 0)
                        mbar_arrive_expect_tx(&full[bs], three ? 3u * HCfg::kBBytes : 1u * HCfg::kBBytes);
                    if (three) {
                        const CUtensorMap* wx = rank == 0 ? whi : wlo;
                        tma_load_2d_2sm(sb, wx, &full[bs], kcol, wrow);
                        tma_load_2d_2sm(sb + HCfg::kBBytes / 2, wx, &full[bs], kcol, wrow + NT / 2);
                    }
                    tma_load_2d_2sm(sb + HCfg::kBBytes, whi, &full[bs], kcol, wrow + static_cast<int>(rank) * (NT / 2));
                } else {
                    mbar_arrive_expect_tx(&full[bs], planes * HCfg::kBBytes);
                    const int wrow_h = wrow + static_cast<int>(rank) * (NT / 2);
                    const int boff = static_cast<int>(rank) * (HCfg::kBBytes / 2);
                    tma_load_2d_mc(sb + boff, whi, &full[bs], kcol, wrow_h, 0x3);
                    if (three) tma_load_2d_mc(sb + HCfg::kBBytes + boff, wlo, &full[bs], kcol, wrow_h, 0x3);
                }
                if (++bs == NS) { bs = 0; bph ^= 1u; }
            };
            for (int item = cluster_id; item < p.total_items; item += num_clusters) {
                const int nt = item % p.tiles_n;
                int r = 2 * (item / p.tiles_n) + static_cast<int>(rank);
                const int wt = r % p.tiles_w; r /= p.tiles_w;
                const int ht = r % p.tiles_h; r /= p.tiles_h;
                const int w0 = wt << 3, h0 = ht << 4, b0 = r;
                const int n0 = nt * NT;
                const int hal = p.taps == 9 ? 1 : 0;
                const uint32_t patch_bytes = p.taps == 9 ? kPatchBytes : 16384u;
                // patch order per tile: shortcut K-blocks first (one tap each), then the main ones, so
                // that the slot the NEXT tile's first patch needs is released half-way through this
                // tile's MMAs.  Every agent waits for every slot in ring order (even those it does not
                // fill): an mbarrier parity wait must never fall a whole phase behind.
                for (int kb = 0; kb < p.kblocks2; ++kb) {
                    mbar_wait(&aempty[as_], aph ^ 1u);
                    if (p.gx0 == nullptr) {       // split planes by TMA (otherwise the worker warps build it)
                        uint8_t* sa = stage_base + as_ * HCfg::kAStage;
                        if (p.debug & 2) {
                            mbar_arrive(&afull[as_]);
                        } else if constexpr (TWO) {
                            if (rank == 0) mbar_arrive_expect_tx(&afull[as_], 2u * planes * 16384u);
                            tma_load_4d_2sm(sa, &tm_a2_hi, &afull[as_], kb * 64, w0, h0, b0);
                            if (three) tma_load_4d_2sm(sa + kPatchPlane, &tm_a2_lo, &afull[as_], kb * 64, w0, h0, b0);
                        } else {
                            mbar_arrive_expect_tx(&afull[as_], planes * 16384u);
                            tma_load_4d(sa, &tm_a2_hi, &afull[as_], kb * 64, w0, h0, b0);
                            if (three) tma_load_4d(sa + kPatchPlane, &tm_a2_lo, &afull[as_], kb * 64, w0, h0, b0);
                        }
                    }
                    if (++as_ == NA) { as_ = 0; aph ^= 1u; }
                    load_weights(&tm_w2_hi, &tm_w2_lo, kb * 64, n0);
                }
                for (int kb = 0; kb < p.kblocks; ++kb) {
                    mbar_wait(&aempty[as_], aph ^ 1u);
                    if (p.fx0 == nullptr) {
                        uint8_t* sa = stage_base + as_ * HCfg::kAStage;
                        if (p.debug & 2) {
                            mbar_arrive(&afull[as_]);
                        } else if constexpr (TWO) {
                            if (rank == 0) mbar_arrive_expect_tx(&afull[as_], 2u * planes * patch_bytes);
                            tma_load_4d_2sm(sa, &tm_a_hi, &afull[as_], kb * 64, w0 - hal, h0 - hal, b0);
                            if (three)
                                tma_load_4d_2sm(sa + kPatchPlane, &tm_a_lo, &afull[as_], kb * 64, w0 - hal, h0 - hal, b0);
                        } else {
                            mbar_arrive_expect_tx(&afull[as_], planes * patch_bytes);
                            tma_load_4d(sa, &tm_a_hi, &afull[as_], kb * 64, w0 - hal, h0 - hal, b0);
                            if (three)
                                tma_load_4d(sa + kPatchPlane, &tm_a_lo, &afull[as_], kb * 64, w0 - hal, h0 - hal, b0);
                        }
                    }
                    if (++as_ == NA) { as_ = 0; aph ^= 1u; }
                    for (int tap = 0; tap < p.taps; ++tap)
                        load_weights(&tm_w_hi, &tm_w_lo, kb * 64, tap * p.Cout_pad + n0);
                }
            }
        }
      } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (halo mode)
        // The WHOLE warp walks the pipeline in convergent control flow and one elected lane issues, so
        // that barrier addresses, descriptors and the accumulate flag stay in uniform registers: a
        // single-thread issuer spent ~20 vector instructions (elect / R2UR / vote loops) per
        // tcgen05.mma and capped the tensor pipe at 68 % active (ncu, round 1).  Descriptors are
        // (address >> 4) words advanced by immediates; the high word is a constant.
        if constexpr (HALO) {
            if (!(TWO && rank != 0)) {
                const bool leader = elect_one();
                constexpr uint32_t idesc_n = umma_idesc_f16(TWO ? 256 : 128, NT);
                constexpr uint32_t idesc_2n = umma_idesc_f16(TWO ? 256 : 128, 2 * NT);
                // high descriptor word: SBO >> 4 | version 1 (bit 46) | SWIZZLE_128B (bits 61-63)
                constexpr uint32_t kHiPatch = ((kPatchW * 128) >> 4) | (1u << 14) | (2u << 29);
                constexpr uint32_t kHiPlain = (1024u >> 4) | (1u << 14) | (2u << 29);
                auto desc = [](uint32_t lo, uint32_t hi) {
                    uint64_t d;
                    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
                    return d;
                };
                auto mma = [&](uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc_flag) {
                    if constexpr (TWO) umma_f16_2cta(d, a, b, idesc, acc_flag);
                    else umma_f16(d, a, b, idesc, acc_flag);
                };
#if DSEP_FP8_CORR
                auto mma8 = [&](uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc_flag) {
                    if constexpr (TWO) umma_e4m3_2cta(d, a, b, idesc, acc_flag);
                    else umma_e4m3(d, a, b, idesc, acc_flag);
                };
#endif
                auto commit_pair = [&](uint64_t* bar) {      // arrive in BOTH CTAs when the MMAs so far retire
                    if constexpr (TWO) umma_commit_2cta_mc(bar, 0x3);
                    else umma_commit_mc(bar, 0x3);
                };
                auto commit_local = [&](uint64_t* bar) {     // TWO: the follower's barriers need the arrival too
                    if constexpr (TWO) umma_commit_2cta_mc(bar, 0x3);
                    else umma_commit(bar);
                };
                const uint32_t a_ring = smem_u32(stage_base) >> 4;
                const uint32_t b_ring = smem_u32(stage_base + NA * HCfg::kAStage) >> 4;
                const bool do_mma = !(p.debug & 1);
                int bs = 0, as_ = 0;
                uint32_t bph = 0, aph = 0;
                int it = 0;
                for (int item = cluster_id; item < p.total_items; item += num_clusters, ++it) {
                    const int acc = it & 1;
                    mbar_wait(&tempty[acc], ((it >> 1) & 1) ^ 1u);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * 2 * NT;
                    uint32_t accumulate = 0;
#if DSEP_FP8_CORR
                    uint32_t accumulate8 = 0;        // first e4m3 product of the tile initialises columns [NT, 2NT)
#endif
                    // one weight stage against the A rows whose hi-plane descriptor word is a_word
                    auto issue = [&](uint32_t a_word, uint32_t a_hiword, bool main_kb) {
                        mbar_wait(&full[bs], bph);
                        tc_fence_after();
                        const uint32_t b_word = b_ring + bs * (HCfg::kBStage >> 4);   // W_hi rows, then W_lo rows
                        if (leader) {
                            if (do_mma) {
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    const uint64_t a_hi = desc(a_word + 2 * k, a_hiword);
                                    const uint64_t b_hi = desc(b_word + 2 * k, kHiPlain);
                                    // TWO: region X (this CTA's half of [W_hi ; W_lo]) / region Y (its half of W_hi)
                                    const uint64_t b_y = TWO ? desc(b_word + (HCfg::kBBytes >> 4) + 2 * k, kHiPlain) : b_hi;
#if DSEP_FP8_CORR
                                    if (fp8c) {
                                        const uint64_t a_2 = desc(a_word + (kPatchPlane >> 4) + 2 * k, a_hiword);
                                        const uint64_t b_2 = desc(b_word + (HCfg::kBBytes >> 4) + 2 * k, kHiPlain);
                                        mma(d_tmem, a_hi, b_hi, idesc_n, accumulate);              // hi*hi -> [0, NT)
                                        if (main_kb) {     // [A_lo8 | A_hi8] x [W_hi8 ; W_lo8], K = 32 -> [NT, 2NT)
                                            mma8(d_tmem + NT, a_2, b_2, idesc_n, accumulate8);
                                            accumulate8 = 1;
                                        } else {           // fp16 shortcut K-block: all three products into [0, NT)
                                            mma(d_tmem, a_hi, b_2, idesc_n, 1);
                                            mma(d_tmem, a_2, b_hi, idesc_n, 1);
                                        }
                                    } else
#endif
                                    if (three) {
                                        mma(d_tmem, a_hi, b_hi, idesc_2n, accumulate);
                                        mma(d_tmem, desc(a_word + (kPatchPlane >> 4) + 2 * k, a_hiword), b_y, idesc_n, 1);
                                    } else {
                                        mma(d_tmem, a_hi, b_y, idesc_n, accumulate);
                                    }
                                    accumulate = 1;
                                }
                            }
                            commit_pair(&empty[bs]);
                        }
                        __syncwarp();
                        if (++bs == NS) { bs = 0; bph ^= 1u; }
                    };
                    for (int kb = 0; kb < p.kblocks2; ++kb) {      // fused 1x1 shortcut K-blocks
                        mbar_wait(&afull[as_], aph);
                        tc_fence_after();
                        issue(a_ring + as_ * (HCfg::kAStage >> 4), kHiPlain, false);
                        if (leader) commit_local(&aempty[as_]);
                        __syncwarp();
                        if (++as_ == NA) { as_ = 0; aph ^= 1u; }
                    }
                    for (int kb = 0; kb < p.kblocks; ++kb) {
                        mbar_wait(&afull[as_], aph);
                        tc_fence_after();
                        const uint32_t sa = a_ring + as_ * (HCfg::kAStage >> 4);
                        if (p.taps == 9) {
#pragma unroll
                            for (int tap = 0; tap < 9; ++tap)
                                issue(sa + ((tap / 3) * kPatchW + tap % 3) * 8, kHiPatch, true);
                        } else {
                            issue(sa, kHiPlain, true);
                        }
                        if (leader) commit_local(&aempty[as_]);
                        __syncwarp();
                        if (++as_ == NA) { as_ = 0; aph ^= 1u; }
                    }
                    if (leader) commit_local(&tfull[acc]);
                    __syncwarp();
                }
            }
        }
      }
    } else if (!HALO && warp == 0 && lane == 0) {
        // ------------------------------------------------------------------ TMA producer (per-tap mode)
        int stage = 0;
        uint32_t phase = 0;
        for (int item = cluster_id; item < p.total_items; item += num_clusters) {
            const int nt = item % p.tiles_n;
            int r = 2 * (item / p.tiles_n) + static_cast<int>(rank);      // this CTA's M tile of the pair
            const int wt = r % p.tiles_w; r /= p.tiles_w;
            const int ht = r % p.tiles_h; r /= p.tiles_h;
            // r >= tiles_b only for the odd tile out of the last pair: b0 >= B, so TMA zero-fills
            // and every store is masked
            const int w0 = wt << p.tw_log2, h0 = ht << p.th_log2, b0 = r << tb_log2;
            const int n0 = nt * NT;
            for (int ki = 0; ki < kiters; ++ki) {
                const bool second = ki >= p.taps * p.kblocks;
                int tap = 0, kb, dy = 0, dx = 0;
                if (second) {
                    kb = ki - p.taps * p.kblocks;
                } else {
                    tap = ki / p.kblocks;
                    kb = ki - tap * p.kblocks;
                    if (p.taps == 9) { dy = tap / 3 - 1; dx = tap % 3 - 1; }
                }
                mbar_wait(&empty[stage], phase ^ 1u);
                uint8_t* s = stage_base + stage * Cfg::kStageBytes;
                if (p.debug & 2) {
                    mbar_arrive(&full[stage]);
                    if (++stage == NS) { stage = 0; phase ^= 1u; }
                    continue;
                }
                mbar_arrive_expect_tx(&full[stage], stage_tx);
                const int wrow = second ? n0 : tap * p.Cout_pad + n0;
                const int wrow_h = wrow + static_cast<int>(rank) * (NT / 2);        // my half of the weight rows
                const int boff = static_cast<int>(rank) * (Cfg::kBBytes / 2);
                tma_load_4d(s, second ? &tm_a2_hi : &tm_a_hi, &full[stage], kb * kBK, w0 + dx, h0 + dy, b0);
                tma_load_2d_mc(s + 2 * kABytes + boff, second ? &tm_w2_hi : &tm_w_hi, &full[stage], kb * kBK,
                               wrow_h, 0x3);
                if (three) {
                    tma_load_4d(s + kABytes, second ? &tm_a2_lo : &tm_a_lo, &full[stage], kb * kBK, w0 + dx,
                                h0 + dy, b0);
                    tma_load_2d_mc(s + 2 * kABytes + Cfg::kBBytes + boff, second ? &tm_w2_lo : &tm_w_lo,
                                   &full[stage], kb * kBK, wrow_h, 0x3);
                }
                if (++stage == NS) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (!HALO && warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (per-tap mode)
        // whole warp, convergent, one elected lane issues: descriptors stay in uniform registers (see halo mode)
        const bool leader = elect_one();
        constexpr uint32_t idesc_n = umma_idesc_f16(128, NT);
        constexpr uint32_t idesc_2n = umma_idesc_f16(128, 2 * NT);
        constexpr uint32_t kHi = kBK == 64 ? ((1024u >> 4) | (1u << 14) | (2u << 29)) : ((512u >> 4) | (1u << 14) | (4u << 29));
        auto desc = [](uint32_t lo, uint32_t hi) {
            uint64_t d;
            asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
            return d;
        };
        const uint32_t ring = smem_u32(stage_base) >> 4;
        const bool do_mma = !(p.debug & 1);
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        for (int item = cluster_id; item < p.total_items; item += num_clusters, ++it) {
            const int as = it & 1;
            mbar_wait(&tempty[as], ((it >> 1) & 1) ^ 1u);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + as * 2 * NT;
            uint32_t accumulate = 0;
            for (int ki = 0; ki < kiters; ++ki) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                const uint32_t s = ring + stage * (Cfg::kStageBytes >> 4);
                if (leader) {
                    if (do_mma) {
#pragma unroll
                        for (int k = 0; k < kBK / 16; ++k) {
                            const uint64_t a_hi = desc(s + 2 * k, kHi);
                            const uint64_t b_hi = desc(s + ((2 * kABytes) >> 4) + 2 * k, kHi);   // W_hi rows, then W_lo rows
                            if (three) {
                                umma_f16(d_tmem, a_hi, b_hi, idesc_2n, accumulate);
                                umma_f16(d_tmem, desc(s + (kABytes >> 4) + 2 * k, kHi), b_hi, idesc_n, 1);
                            } else {
                                umma_f16(d_tmem, a_hi, b_hi, idesc_n, accumulate);
                            }
                            accumulate = 1;
                        }
                    }
                    umma_commit_mc(&empty[stage], 0x3);   // frees this slot in BOTH CTAs once these MMAs retire
                }
                __syncwarp();
                if (++stage == NS) { stage = 0; phase ^= 1u; }
            }
            if (leader) umma_commit(&tfull[as]);          // accumulator complete -> epilogue
            __syncwarp();
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------------ epilogue
        if constexpr (HALO) asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
        const int ew = warp - 4;
        const int wq = ew & 3;                // TMEM lane quarter == warp_id % 4
        // sum of the two accumulator halves: hi*hi (+ lo*hi) and hi*lo — or, with e4m3 corrections, the correction
        // accumulator weighted by the ratio of the operand prescales
#if DSEP_FP8_CORR
        const float crel = fp8c ? p.corr_rel : 1.0f;
#define EPI_ADD(main_, corr_) fmaf((corr_), crel, (main_))
#else
#define EPI_ADD(main_, corr_) ((main_) + (corr_))
#endif
        const int tw_mask = (1 << p.tw_log2) - 1, th_mask = (1 << p.th_log2) - 1;
        // GroupNorm partial sums of this thread's 4 channels, carried across consecutive tiles of the
        // same (batch entry, channel tile) and flushed with fp64 atomics only when that changes
        float4 run1[NT >= 64 ? NT / 64 : 1], run2[NT >= 64 ? NT / 64 : 1];
        int run_b = -1, run_n0 = -1;
        auto flush_stats = [&]() {
            if constexpr (NT >= 64) {
                if (p.stats == nullptr || run_b < 0) return;           // warp-uniform
#pragma unroll
                for (int c = 0; c < NT / 64; ++c) {                     // rows of the 4 lane groups -> lanes 0..7
#pragma unroll
                    for (int o = 8; o <= 16; o <<= 1) {
                        run1[c].x += __shfl_xor_sync(0xffffffffu, run1[c].x, o);
                        run1[c].y += __shfl_xor_sync(0xffffffffu, run1[c].y, o);
                        run1[c].z += __shfl_xor_sync(0xffffffffu, run1[c].z, o);
                        run1[c].w += __shfl_xor_sync(0xffffffffu, run1[c].w, o);
                        run2[c].x += __shfl_xor_sync(0xffffffffu, run2[c].x, o);
                        run2[c].y += __shfl_xor_sync(0xffffffffu, run2[c].y, o);
                        run2[c].z += __shfl_xor_sync(0xffffffffu, run2[c].z, o);
                        run2[c].w += __shfl_xor_sync(0xffffffffu, run2[c].w, o);
                    }
                }
                if (run_b >= p.B || (lane >> 3) != 0) return;
#pragma unroll
                for (int c = 0; c < NT / 64; ++c) {
                    const int n = run_n0 + (ew >> 2) * (NT / 2) + c * 32 + (lane & 7) * 4;
                    if (n < p.cout_store) {
                        double* st = p.stats + (static_cast<size_t>(run_b) * p.cout_store + n) * 2;
                        atomicAdd(st + 0, (double)run1[c].x); atomicAdd(st + 1, (double)run2[c].x);
                        atomicAdd(st + 2, (double)run1[c].y); atomicAdd(st + 3, (double)run2[c].y);
                        atomicAdd(st + 4, (double)run1[c].z); atomicAdd(st + 5, (double)run2[c].z);
                        atomicAdd(st + 6, (double)run1[c].w); atomicAdd(st + 7, (double)run2[c].w);
                    }
                }
            }
        };
        auto do_epilogue = [&](int item, int it) {
            const int as = it & 1;
            const int nt = item % p.tiles_n;
            int r = 2 * (item / p.tiles_n) + static_cast<int>(rank);      // this CTA's M tile of the pair
            const int wt = r % p.tiles_w; r /= p.tiles_w;
            const int ht = r % p.tiles_h; r /= p.tiles_h;
            // r >= tiles_b only for the odd tile out of the last pair: b0 >= B, so TMA zero-fills
            // and every store is masked
            const int w0 = wt << p.tw_log2, h0 = ht << p.th_log2, b0 = r << tb_log2;
            const int n0 = nt * NT;
            const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + as * 2 * NT;
            if constexpr (NT >= 64) {
                if (p.stats != nullptr && (b0 != run_b || n0 != run_n0)) {
                    flush_stats();
                    run_b = b0; run_n0 = n0;
#pragma unroll
                    for (int c = 0; c < NT / 64; ++c) run1[c] = run2[c] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }

            // ---- fast path (halo tiles entirely inside the image and the stored channels: every tile of the
            // network's maps >= 16 x 8): no per-row predicates, one 64-bit base per thread + constant strides,
            // explicit shared-space staging accesses, scale folded into the FMAs.  ~2.5x fewer instructions
            // than the generic path below (which spent ~55 per float4 on predicates and 64-bit addresses).
            if constexpr (HALO && NT >= 64) {
                if (b0 < p.B && h0 + 16 <= p.H && w0 + 8 <= p.W && n0 + NT <= p.cout_store) {
                    constexpr int kChunks = NT / 64;
                    const int chalf = ew >> 2;
                    const int q = lane & 7, rg = lane >> 3;
                    const uint32_t C = static_cast<uint32_t>(p.cout_store);
                    // row i of this thread: pixel (h0 + wq*4 + (i >> 1), w0 + rg + 4*(i & 1)), 4 channels from n
                    const int n = n0 + chalf * (NT / 2) + q * 4;
                    const size_t e0 = ((static_cast<size_t>(b0) * p.H + h0 + wq * 4) * p.W + w0 + rg) * C + n;
                    float* const out0 = p.out + e0;
                    const float* const res0 = p.residual != nullptr ? p.residual + e0 : nullptr;
                    const uint32_t d_row = static_cast<uint32_t>(p.W) * C;      // i -> i + 2: next image row
                    const uint32_t d_half = 4u * C;                            // odd i: 4 pixels to the right
                    const uint32_t stg_s = smem_u32(staging + ew * (32 * 32));
                    const uint32_t st_base = stg_s + lane * 128 + ((lane & 7) << 4);   // chunk j at ^ (j << 4)
                    const uint32_t ld_base = stg_s + rg * 128 + ((q ^ rg) << 4);        // row i at + i*512, ^ ((i&1) << 6)
                    const float as2 = p.acc_scale * p.scale;
                    const bool store = !(p.debug & 4);
                    bool waited = false;
#pragma unroll
                    for (int c = 0; c < kChunks; ++c) {
                        float4 res[8];
                        if (res0 != nullptr) {     // prefetch the residual while the accumulator is still being produced
#pragma unroll
                            for (int i = 0; i < 8; ++i)
                                res[i] = __ldg(reinterpret_cast<const float4*>(res0 + c * 32 + (i >> 1) * d_row + (i & 1) * d_half));
                        }
                        float4 bz = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (p.bias != nullptr) bz = __ldg(reinterpret_cast<const float4*>(p.bias + n + c * 32));
                        if (p.film != nullptr) {
                            const float4 f = __ldg(reinterpret_cast<const float4*>(
                                p.film + static_cast<size_t>(b0) * p.film_stride + n + c * 32));
                            bz.x += f.x; bz.y += f.y; bz.z += f.z; bz.w += f.w;
                        }
                        bz.x *= p.scale; bz.y *= p.scale; bz.z *= p.scale; bz.w *= p.scale;
                        if (!waited) {
                            mbar_wait(&tfull[as], (it >> 1) & 1);
                            tc_fence_after();
                            waited = true;
                        }
                        const int col0 = chalf * (NT / 2) + c * 32;
                        uint32_t v[32];
                        tmem_ld_32x32(t_addr + col0, v);
                        if (three) {   // add the hi*lo half, 16 columns at a time (register pressure)
#pragma unroll
                            for (int hh = 0; hh < 2; ++hh) {
                                uint32_t u[16];
                                tmem_ld_32x16(t_addr + NT + col0 + hh * 16, u);
                                tmem_ld_wait();
#pragma unroll
                                for (int j = 0; j < 16; ++j)
                                    v[hh * 16 + j] = __float_as_uint(EPI_ADD(__uint_as_float(v[hh * 16 + j]), __uint_as_float(u[j])));
                            }
                        } else {
                            tmem_ld_wait();
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
                            a.x = fmaf(a.x, as2, bz.x); a.y = fmaf(a.y, as2, bz.y);
                            a.z = fmaf(a.z, as2, bz.z); a.w = fmaf(a.w, as2, bz.w);
                            if (res0 != nullptr) {
                                a.x = fmaf(res[i].x, p.scale, a.x); a.y = fmaf(res[i].y, p.scale, a.y);
                                a.z = fmaf(res[i].z, p.scale, a.z); a.w = fmaf(res[i].w, p.scale, a.w);
                            }
                            if (store)
                                *reinterpret_cast<float4*>(out0 + c * 32 + (i >> 1) * d_row + (i & 1) * d_half) = a;
                            s1.x += a.x; s1.y += a.y; s1.z += a.z; s1.w += a.w;
                            s2.x = fmaf(a.x, a.x, s2.x); s2.y = fmaf(a.y, a.y, s2.y);
                            s2.z = fmaf(a.z, a.z, s2.z); s2.w = fmaf(a.w, a.w, s2.w);
                        }
                        if (p.stats != nullptr) {
                            run1[c].x += s1.x; run1[c].y += s1.y; run1[c].z += s1.z; run1[c].w += s1.w;
                            run2[c].x += s2.x; run2[c].y += s2.y; run2[c].z += s2.z; run2[c].w += s2.w;
                        }
                        __syncwarp();
                    }
                    return;
                }
            }

            if constexpr (NT >= 64) {
                constexpr int kChunks = NT / 64;          // 32-column chunks per warp
                const int chalf = ew >> 2;                // which half of the tile's columns
                float* stg = staging + ew * (32 * 32);
                const int q = lane & 7, rg = lane >> 3;
                // pixel of each of this thread's 8 output rows (row = i*4 + rg of this warp's 32);
                // kNoPix marks rows outside the image / batch
                constexpr uint32_t kNoPix = 0xFFFFFFFFu;
                uint32_t pix[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int m = wq * 32 + i * 4 + rg;
                    const int w = w0 + (m & tw_mask);
                    const int h = h0 + ((m >> p.tw_log2) & th_mask);
                    const int b = b0 + (m >> (p.tw_log2 + p.th_log2));
                    pix[i] = (b < p.B && h < p.H && w < p.W) ? static_cast<uint32_t>((b * p.H + h) * p.W + w) : kNoPix;
                }
                bool waited = false;
#pragma unroll
                for (int c = 0; c < kChunks; ++c) {
                    const int col0 = chalf * (NT / 2) + c * 32;
                    const int n = n0 + col0 + q * 4;
                    const bool n_ok = n < p.cout_store;
                    // prefetch the residual while the accumulator is still being produced / loaded
                    float4 res[8];
                    if (p.residual != nullptr) {
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            res[i] = (n_ok && pix[i] != kNoPix)
                                         ? __ldg(reinterpret_cast<const float4*>(
                                               p.residual + static_cast<size_t>(pix[i]) * p.cout_store + n))
                                         : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    if (!waited) {
                        mbar_wait(&tfull[as], (it >> 1) & 1);
                        tc_fence_after();
                        waited = true;
                    }
                    uint32_t v[32];
                    tmem_ld_32x32(t_addr + col0, v);
                    if (three) {   // add the hi*lo half, 16 columns at a time (register pressure)
#pragma unroll
                        for (int hh = 0; hh < 2; ++hh) {
                            uint32_t u[16];
                            tmem_ld_32x16(t_addr + NT + col0 + hh * 16, u);
                            tmem_ld_wait();
#pragma unroll
                            for (int j = 0; j < 16; ++j)
                                v[hh * 16 + j] = __float_as_uint(EPI_ADD(__uint_as_float(v[hh * 16 + j]), __uint_as_float(u[j])));
                        }
                    } else {
                        tmem_ld_wait();
                    }
                    if (c == kChunks - 1) {   // TMEM fully drained by this warp: hand the buffer back early
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) {
                            if (TWO && rank != 0) mbar_arrive_cluster(&tempty[as], 0);   // the leader issues the MMAs
                            else mbar_arrive(&tempty[as]);
                        }
                    }
                    // transpose through smem: row = lane, 16-byte chunk j stored at j ^ (lane & 7)
                    float4* dst = reinterpret_cast<float4*>(stg + lane * 32);
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        dst[j ^ (lane & 7)] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                                          __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
                    __syncwarp();
                    float4 bz = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (p.bias != nullptr && n_ok) bz = __ldg(reinterpret_cast<const float4*>(p.bias + n));
                    // one batch entry per tile (every map of at least 128 pixels): FiLM joins the bias once
                    const bool film_rows = p.film != nullptr && tb_log2 != 0;
                    if (p.film != nullptr && tb_log2 == 0 && n_ok && b0 < p.B) {
                        const float4 f = __ldg(reinterpret_cast<const float4*>(
                            p.film + static_cast<size_t>(b0) * p.film_stride + n));
                        bz.x += f.x; bz.y += f.y; bz.z += f.z; bz.w += f.w;
                    }
                    float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int row = i * 4 + rg;
                        if (n_ok && pix[i] != kNoPix) {
                            float4 a = *reinterpret_cast<const float4*>(stg + row * 32 + ((q ^ (row & 7)) << 2));
                            a.x = fmaf(a.x, p.acc_scale, bz.x); a.y = fmaf(a.y, p.acc_scale, bz.y);
                            a.z = fmaf(a.z, p.acc_scale, bz.z); a.w = fmaf(a.w, p.acc_scale, bz.w);
                            if (film_rows) {
                                const int bi = b0 + ((wq * 32 + row) >> (p.tw_log2 + p.th_log2));
                                const float4 f = __ldg(reinterpret_cast<const float4*>(
                                    p.film + static_cast<size_t>(bi) * p.film_stride + n));
                                a.x += f.x; a.y += f.y; a.z += f.z; a.w += f.w;
                            }
                            if (p.residual != nullptr) {
                                a.x += res[i].x; a.y += res[i].y; a.z += res[i].z; a.w += res[i].w;
                            }
                            a.x *= p.scale; a.y *= p.scale; a.z *= p.scale; a.w *= p.scale;
                            if (!(p.debug & 4))
                                *reinterpret_cast<float4*>(p.out + static_cast<size_t>(pix[i]) * p.cout_store + n) = a;
                            s1.x += a.x; s1.y += a.y; s1.z += a.z; s1.w += a.w;
                            s2.x = fmaf(a.x, a.x, s2.x); s2.y = fmaf(a.y, a.y, s2.y);
                            s2.z = fmaf(a.z, a.z, s2.z); s2.w = fmaf(a.w, a.w, s2.w);
                        }
                    }
                    if (p.stats != nullptr) {
                        // all 32 rows of this warp belong to one batch entry (host guarantees H*W >= 128);
                        // the cross-lane reduction waits until the flush
                        run1[c].x += s1.x; run1[c].y += s1.y; run1[c].z += s1.z; run1[c].w += s1.w;
                        run2[c].x += s2.x; run2[c].y += s2.y; run2[c].z += s2.z; run2[c].w += s2.w;
                    }
                    __syncwarp();
                }
            } else {
                // narrow output (pyramid convs, Cout = 6 padded to 16): thread == pixel; warps 8-11 idle
                if (ew < 4) {
                    mbar_wait(&tfull[as], (it >> 1) & 1);
                    tc_fence_after();
                    uint32_t v[16];
                    tmem_ld_32x16(t_addr, v);
                    if (three) {
                        uint32_t u[16];
                        tmem_ld_32x16(t_addr + NT, u);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(EPI_ADD(__uint_as_float(v[j]), __uint_as_float(u[j])));
                    } else {
                        tmem_ld_wait();
                    }
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
        };

        bool worker_builds = false;
        if constexpr (HALO) worker_builds = p.fx0 != nullptr || p.gx0 != nullptr;
        if (!worker_builds) {
            int it = 0;
            for (int item = cluster_id; item < p.total_items; item += num_clusters, ++it) do_epilogue(item, it);
        } else {
            if constexpr (HALO) {
                // Worker schedule per tile i: build the patches of tile i that are theirs (in ring order:
                // shortcut K-blocks, then main), then run the epilogue of tile i-1 while the tensor core
                // chews on the main patches.  Patches live in the 2-slot A ring shared with the TMA
                // producer (which fills those that still arrive as planes); the epilogue of tile i-1
                // only has to finish before the MMAs of tile i+1 (double-buffered TMEM).
                const int wtid = static_cast<int>(threadIdx.x) - 128;
                int as_ = 0;
                uint32_t aph = 0;
                int it = 0, prev_item = -1;
                for (int item = cluster_id;; item += num_clusters, ++it) {      // one call site each: one inlined copy
                    const bool more = item < p.total_items;
                    if (more) build_tile_patches(item, wtid, as_, aph);
                    if (prev_item >= 0) do_epilogue(prev_item, it - 1);
                    if (!more) break;
                    prev_item = item;
                }
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

#undef EPI_ADD

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
    const CUtensorMapSwizzle swz = kBK == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    EncodeTiledFn enc = get_encode_tiled();
    if (enc == nullptr) {
        set_error("cuTensorMapEncodeTiled not available from the CUDA driver");
        return DSEP_ERR_CUDA;
    }
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, const_cast<void*>(base), dims,
                     strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
        return DSEP_ERR_CUDA;
    }
    return DSEP_OK;
}

int conv_num_sms() {
    static int n[kMaxDevices] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) dev = 0;
    if (n[dev] == 0) {
        int v = 0;
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        n[dev] = v > 0 ? v : 148;
    }
    return n[dev];
}

template <int NT, bool HALO, bool TWO = false>
static int launch_conv(const ConvMaps& m, const ConvParams& p, cudaStream_t stream) {
    constexpr int kSmem = HALO ? HaloCfg<NT>::kSmemBytes : ConvCfg<NT>::kSmemBytes;
    static PerDeviceAttr attr;
    const cudaError_t attr_err = set_max_smem_once(attr, conv_tc_kernel<NT, HALO, TWO>, kSmem);
    if (attr_err != cudaSuccess) {
        set_error("cudaFuncSetAttribute(conv_tc_kernel<%d,%d>): %s", NT, (int)HALO, cudaGetErrorString(attr_err));
        return DSEP_ERR_CUDA;
    }
    const int max_clusters = conv_num_sms() / 2;
    const int grid = 2 * (p.total_items < max_clusters ? p.total_items : max_clusters);
    conv_tc_kernel<NT, HALO, TWO><<<grid, kThreads, kSmem, stream>>>(
        m.a_hi, m.a_lo, m.w_hi, m.w_lo, m.a2_hi, m.a2_lo, m.w2_hi, m.w2_lo, p);
    return check_launch("conv_tc_kernel");
}

static int ilog2(int v) {
    int l = 0;
    while ((1 << l) < v) ++l;
    return l;
}

}  // namespace dsep

extern "C" int dsep_conv_kblock(void) { return dsep::kBK; }

static int g_wide_launches = 0;

// corr_rel / a8_exp: only for passes = 2 (dsep_conv2d_fused8); dsep_conv2d_fused passes (1, 0)
static int conv_dispatch(const dsep_conv_args* g, float corr_rel, int a8_exp, dsep_stream_t stream) {
    using namespace dsep;
    DSEP_REQUIRE(g != nullptr, "conv2d: null argument block");
    const int B = g->B, H = g->H, W = g->W, Cin = g->Cin, Cout_pad = g->Cout_pad, ksize = g->ksize;
    const int Cin2 = g->Cin2, cout_store = g->cout_store, passes = g->passes;
    const bool main_fused = g->x0 != nullptr, short_fused = g->s0 != nullptr;
    DSEP_REQUIRE((g->a_hi || main_fused) && g->w_hi && g->out, "conv2d_tc: null operand");
#if DSEP_FP8_CORR
    DSEP_REQUIRE(passes == 1 || passes == 2 || passes == 3, "conv2d_tc: passes must be 1, 2 or 3 (got %d)", passes);
    DSEP_REQUIRE(passes != 2 || (kBK == 64 && Cout_pad % 64 == 0 && H >= 16 && W >= 8 && corr_rel > 0.f),
                 "conv2d_fused8: e4m3 corrections need Cout >= 64 and a map of at least 16 x 8");
#else
    DSEP_REQUIRE(passes == 1 || passes == 3, "conv2d_tc: passes must be 1 or 3 (got %d)", passes);
    (void)corr_rel; (void)a8_exp;
#endif
    DSEP_REQUIRE(passes == 1 || ((g->a_lo || main_fused) && g->w_lo), "conv2d_tc: passes=3 needs the lo planes");
    DSEP_REQUIRE(ksize == 1 || ksize == 3, "conv2d_tc: ksize must be 1 or 3 (got %d)", ksize);
    DSEP_REQUIRE(B > 0 && H > 0 && W > 0, "conv2d_tc: empty tensor");
    DSEP_REQUIRE(Cin > 0 && Cin % kBK == 0, "conv2d_tc: Cin must be a multiple of %d (got %d)", kBK, Cin);
    DSEP_REQUIRE(Cout_pad == 16 || (Cout_pad > 0 && Cout_pad % 64 == 0),
                 "conv2d_tc: Cout_pad must be 16 or a multiple of 64 (got %d)", Cout_pad);
    DSEP_REQUIRE(cout_store > 0 && cout_store <= Cout_pad, "conv2d_tc: bad cout_store %d", cout_store);
    DSEP_REQUIRE(Cout_pad == 16 || cout_store % 4 == 0, "conv2d_tc: cout_store must be a multiple of 4");
    DSEP_REQUIRE(Cin2 >= 0 && Cin2 % kBK == 0, "conv2d_tc: Cin2 must be a multiple of %d (got %d)", kBK, Cin2);
    DSEP_REQUIRE(Cin2 == 0 || ((g->a2_hi || short_fused) && g->w2_hi &&
                               (passes == 1 || ((g->a2_lo || short_fused) && g->w2_lo))),
                 "conv2d_tc: fused 1x1 operand (Cin2=%d) needs its activation and weight planes", Cin2);
    int NT = Cout_pad == 16 ? 16 : (Cout_pad % 128 == 0 ? 128 : 64);
    // small maps (the 4x4 / 8x8 levels: 4 - 16 pixel tiles): with 128-channel tiles only 8 - 32 CTAs have work and each
    // walks its whole K loop alone (~30 us per launch, MMA-bound per CTA); 64-channel tiles put twice the SMs on it
    if (NT == 128 && !main_fused && !short_fused &&
        (int64_t)ceil_div(B * H * W, 128) * (Cout_pad / 128) * 2 <= conv_num_sms() / 2)
        NT = 64;

    ConvParams p{};
    p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.Cout_pad = Cout_pad; p.cout_store = cout_store;
    p.taps = ksize * ksize;
    int tw = 1 << ilog2(W); if (tw > 16) tw = 16;
    int th = 1 << ilog2(H); if (th > 128 / tw) th = 128 / tw;
    // halo mode: maps of at least 16 x 8 and wide output tiles; fixed 8 (w) x 16 (h) tile
    static const int halo_env = getenv("DSEP_CONV_HALO") ? atoi(getenv("DSEP_CONV_HALO")) : 1;
    // narrow outputs (Cout_pad = 16: the pyramid convs) take the halo kernel only with the in-kernel prologue
    const bool halo_ok = kBK == 64 && W >= 8 && H >= 16 && (Cout_pad != 16 || (main_fused && Cin2 == 0));
    const bool halo = halo_ok && (main_fused || short_fused || passes == 2 || (halo_env != 0 && ksize == 3));
    DSEP_REQUIRE(!(main_fused || short_fused) || halo_ok,
                 "conv2d_fused: the in-kernel prologue needs a map of at least 16 x 8 (and no fused shortcut when "
                 "Cout_pad is 16; got %dx%d, Cout_pad %d)", H, W, Cout_pad);
    // all A patches of a launch come from ONE agent (TMA producer or worker warps): two agents sharing the
    // patch ring could fall a whole mbarrier phase apart
    DSEP_REQUIRE(Cin2 == 0 || main_fused == short_fused,
                 "conv2d_fused: main and shortcut operands must both be fp32 (in-kernel prologue) or both planes");
    if (main_fused) {
        DSEP_REQUIRE(g->C0 > 0 && g->C0 % 64 == 0 && g->C1 >= 0 && g->C1 % 64 == 0 && g->C0 + g->C1 == Cin &&
                         (g->C1 == 0 || g->x1 != nullptr),
                     "conv2d_fused: main operand channels (%d + %d) must be multiples of 64 adding up to Cin=%d",
                     g->C0, g->C1, Cin);
        DSEP_REQUIRE((g->sc == nullptr) == (g->sh == nullptr), "conv2d_fused: sc and sh come together");
    }
    if (short_fused) {
        DSEP_REQUIRE(g->S0 > 0 && g->S0 % 64 == 0 && g->S1 >= 0 && g->S1 % 64 == 0 && g->S0 + g->S1 == Cin2 &&
                         (g->S1 == 0 || g->s1 != nullptr),
                     "conv2d_fused: shortcut operand channels (%d + %d) must be multiples of 64 adding up to Cin2=%d",
                     g->S0, g->S1, Cin2);
    }
    // wide form (conv_wide.cu): 8 x 32 pixel tiles on the MMA's N side, one accumulator for all three products
    static const int wide_env = getenv("DSEP_CONV_WIDE") ? atoi(getenv("DSEP_CONV_WIDE")) : 1;
    static const int v2_env = getenv("DSEP_CONV_V2") ? atoi(getenv("DSEP_CONV_V2")) : 1;
    static const int two_env = getenv("DSEP_CONV_2CTA") ? atoi(getenv("DSEP_CONV_2CTA")) : 0;
    const bool wide = halo && main_fused && NT == 128 && passes == 2 && a8_exp == 0 && corr_rel == 1.0f &&
                      H % 32 == 0 && W % 8 == 0 && wide_env != 0 && v2_env != 0 && !two_env;
    if (halo) { tw = 8; th = wide ? 32 : 16; }
    const int tb = wide ? 1 : 128 / (tw * th);
    DSEP_REQUIRE(g->stats == nullptr || (Cout_pad != 16 && tb == 1),
                 "conv2d_tc: fused statistics need Cout >= 64 and 128-pixel tiles inside one batch entry "
                 "(got %dx%d, tile %dx%dx%d)", H, W, th, tw, tb);
    p.tw_log2 = ilog2(tw); p.th_log2 = ilog2(th);
    p.tiles_w = ceil_div(W, tw); p.tiles_h = ceil_div(H, th); p.tiles_b = ceil_div(B, tb);
    p.tiles_n = Cout_pad / NT;
    p.total_items = ((p.tiles_w * p.tiles_h * p.tiles_b + 1) / 2) * p.tiles_n;
    p.kblocks = Cin / kBK;
    p.kblocks2 = Cin2 / kBK;
    p.passes = passes;
    p.bias = g->bias; p.film = g->film; p.film_stride = g->film_stride; p.residual = g->residual;
    p.scale = g->scale; p.acc_scale = g->acc_scale; p.out = g->out; p.stats = g->stats;
    p.fx0 = g->x0; p.fx1 = g->x1; p.fC0 = g->C0; p.fC1 = g->C1; p.fsc = g->sc; p.fsh = g->sh; p.fact = g->act;
    p.gx0 = g->s0; p.gx1 = g->s1; p.gC0 = g->S0; p.gC1 = g->S1;
#if DSEP_FP8_CORR
    p.corr_rel = corr_rel;
    p.a8_hi = ldexpf(1.0f, a8_exp);
    p.a8_lo = ldexpf(1.0f, a8_exp + 11);
#endif
    {
        static const int dbg = getenv("DSEP_CONV_DEBUG") ? atoi(getenv("DSEP_CONV_DEBUG")) : 0;
        p.debug = dbg;
    }

    ConvMaps m;
    int rc;
    const bool patch3 = halo && ksize == 3;
    {
        const cuuint64_t wdims[2] = {(cuuint64_t)Cin, (cuuint64_t)p.taps * Cout_pad};
        const cuuint64_t wstr[1] = {(cuuint64_t)Cin * 2};
        const cuuint32_t wbox[2] = {(cuuint32_t)kBK, (cuuint32_t)(NT / 2)};
        if ((rc = make_map(&m.w_hi, g->w_hi, 2, wdims, wstr, wbox)) != DSEP_OK) return rc;
        if (passes != 1) {
            if ((rc = make_map(&m.w_lo, g->w_lo, 2, wdims, wstr, wbox)) != DSEP_OK) return rc;
        } else {
            m.w_lo = m.w_hi;
        }
        if (!main_fused) {
            const cuuint64_t adims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
            const cuuint64_t astr[3] = {(cuuint64_t)Cin * 2, (cuuint64_t)W * Cin * 2, (cuuint64_t)H * W * Cin * 2};
            const cuuint32_t abox[4] = {(cuuint32_t)kBK, (cuuint32_t)(patch3 ? kPatchW : tw),
                                        (cuuint32_t)(patch3 ? kPatchH : th), (cuuint32_t)tb};
            if ((rc = make_map(&m.a_hi, g->a_hi, 4, adims, astr, abox)) != DSEP_OK) return rc;
            if (passes != 1) {
                if ((rc = make_map(&m.a_lo, g->a_lo, 4, adims, astr, abox)) != DSEP_OK) return rc;
            } else {
                m.a_lo = m.a_hi;
            }
        } else {
            m.a_hi = m.w_hi; m.a_lo = m.w_hi;     // unused by the kernel
        }
    }
    m.a2_hi = m.a_hi; m.a2_lo = m.a_lo; m.w2_hi = m.w_hi; m.w2_lo = m.w_lo;
    if (Cin2 > 0) {
        const cuuint64_t wdims[2] = {(cuuint64_t)Cin2, (cuuint64_t)Cout_pad};
        const cuuint64_t wstr[1] = {(cuuint64_t)Cin2 * 2};
        const cuuint32_t wbox[2] = {(cuuint32_t)kBK, (cuuint32_t)(NT / 2)};
        if ((rc = make_map(&m.w2_hi, g->w2_hi, 2, wdims, wstr, wbox)) != DSEP_OK) return rc;
        if (passes != 1) {
            if ((rc = make_map(&m.w2_lo, g->w2_lo, 2, wdims, wstr, wbox)) != DSEP_OK) return rc;
        } else {
            m.w2_lo = m.w2_hi;
        }
        if (!short_fused) {
            const cuuint64_t adims[4] = {(cuuint64_t)Cin2, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
            const cuuint64_t astr[3] = {(cuuint64_t)Cin2 * 2, (cuuint64_t)W * Cin2 * 2, (cuuint64_t)H * W * Cin2 * 2};
            const cuuint32_t abox[4] = {(cuuint32_t)kBK, (cuuint32_t)tw, (cuuint32_t)th, (cuuint32_t)tb};
            if ((rc = make_map(&m.a2_hi, g->a2_hi, 4, adims, astr, abox)) != DSEP_OK) return rc;
            if (passes != 1) {
                if ((rc = make_map(&m.a2_lo, g->a2_lo, 4, adims, astr, abox)) != DSEP_OK) return rc;
            } else {
                m.a2_lo = m.a2_hi;
            }
        }
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    // Cout >= 64 with the operand built in-kernel: the role-split kernels of conv_wide.cu / conv_fused.cu
    // (DSEP_CONV_WIDE=0: 128-pixel tiles only; DSEP_CONV_V2=0: the previous one-worker-role halo kernel below, kept
    // for A/B timing)
    if (wide) {
        ++g_wide_launches;
        return launch_conv_wide(m, p, s);
    }
    if (halo && main_fused && NT >= 64 && passes != 1 && v2_env != 0 && !two_env && (passes != 2 || a8_exp == 0))
        return launch_conv_fused(m, p, NT, s);
    if (halo && NT == 16) return launch_conv<16, true>(m, p, s);
#if DSEP_FP8_CORR
    if (passes == 2 && two_env)
        return NT == 64 ? launch_conv<64, true, true>(m, p, s) : launch_conv<128, true, true>(m, p, s);
    if (passes == 2) return NT == 64 ? launch_conv<64, true>(m, p, s) : launch_conv<128, true>(m, p, s);
#endif
    if (halo && two_env)
        return NT == 64 ? launch_conv<64, true, true>(m, p, s) : launch_conv<128, true, true>(m, p, s);
    if (halo) return NT == 64 ? launch_conv<64, true>(m, p, s) : launch_conv<128, true>(m, p, s);
    switch (NT) {
        case 16: return launch_conv<16, false>(m, p, s);
        case 64: return launch_conv<64, false>(m, p, s);
        default: return launch_conv<128, false>(m, p, s);
    }
}

extern "C" int dsep_conv2d_fused(const dsep_conv_args* g, dsep_stream_t stream) {
    DSEP_REQUIRE(g == nullptr || g->passes != 2, "conv2d_fused: passes = 2 goes through dsep_conv2d_fused8");
    return conv_dispatch(g, 1.0f, 0, stream);
}

extern "C" int dsep_has_fp8_corr(void) { return DSEP_FP8_CORR; }
extern "C" int dsep_conv_wide_launches(void) { return g_wide_launches; }

extern "C" int dsep_conv2d_fused8(const dsep_conv_args* g, float corr_rel, int a8_exp, dsep_stream_t stream) {
#if DSEP_FP8_CORR
    DSEP_REQUIRE(g != nullptr && g->passes == 2, "conv2d_fused8: passes must be 2");
    return conv_dispatch(g, corr_rel, a8_exp, stream);
#else
    (void)g; (void)corr_rel; (void)a8_exp; (void)stream;
    dsep::set_error("conv2d_fused8: this libdsep.so was built without DSEP_FP8_CORR");
    return DSEP_ERR_UNSUPPORTED;
#endif
}

extern "C" int dsep_conv2d_tc(const void* a_hi, const void* a_lo, int B, int H, int W, int Cin,
                              const void* w_hi, const void* w_lo, int Cout_pad, int ksize,
                              const void* a2_hi, const void* a2_lo, int Cin2, const void* w2_hi,
                              const void* w2_lo, const float* bias, const float* film, int film_stride,
                              const float* residual, float scale, float acc_scale, float* out,
                              int cout_store, double* stats, int passes, dsep_stream_t stream) {
    dsep_conv_args g{};
    g.a_hi = a_hi; g.a_lo = a_lo; g.B = B; g.H = H; g.W = W; g.Cin = Cin;
    g.w_hi = w_hi; g.w_lo = w_lo; g.Cout_pad = Cout_pad; g.ksize = ksize;
    g.a2_hi = a2_hi; g.a2_lo = a2_lo; g.Cin2 = Cin2; g.w2_hi = w2_hi; g.w2_lo = w2_lo;
    g.bias = bias; g.film = film; g.film_stride = film_stride; g.residual = residual;
    g.scale = scale; g.acc_scale = acc_scale; g.out = out; g.cout_store = cout_store; g.stats = stats;
    g.passes = passes;
    return dsep_conv2d_fused(&g, stream);
}
