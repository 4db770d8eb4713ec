// k_small.cuh -- draws of 2, 3 or 4 reads (the CLI default -M 4).
//
// A group of LPD lanes owns one draw, a warp owns a tile of 32/LPD draws, and every warp of a
// persistent one-CTA-per-SM grid (16 warps) walks its tiles through a shared-memory pipeline:
//
//   producer  the NR*32/LPD lanes that hold the tile's read indices each issue ONE bulk copy
//             (cp.async.bulk, the 1-D TMA path: UBLKCP) of a whole read mask, HBM/L2 -> shared,
//             completing on the warp's mbarrier.  The buffer is refilled with the NEXT tile as soon
//             as the current tile's masks have been pulled into registers, so the copy lands while
//             the tile is being computed and the kernel is not exposed to DRAM latency.
//   consumer  each lane pulls its 16-byte slots of the NR masks into registers (lane l of a draw's
//             group takes slot j*LPD + l: conflict-free LDS.128), forms every subset intersection
//             of two or more reads (joint_frequencies_combo, HOMOGENOUES_MODELS.py:129-163; one
//             LOP3 ANDs up to three reads), popcounts the lane's words of a subset together through
//             a carry-save adder tree (POPC is the scarce pipe: 16/clk/SM against 64 LOP3), and the
//             partial counts are summed across the group's lanes with xor-shuffles.  Single-read
//             frequencies are row popcounts, precomputed once per read and prefetched with the
//             indices.  The closed-form polynomials (poly_closed.cuh) follow in float64; counts are
//             normalised through an exact x/N lookup table instead of 15 divisions.  For one-group
//             models and recent admixture the counts of 32 draws wait in a small per-warp table and
//             the lookups + polynomial run once per 32 draws, one draw per lane (a tile's polynomial
//             would otherwise be executed by all 8 lanes of each of its 4 draws).
//
// Algorithmic work per draw: G*(2^NR-1)*W word AND+POPC, G*NR*W*8 mask bytes in, 32 B out.
//
// k_small_direct is the unpipelined variant (masks loaded straight into registers) used for rows
// longer than one pass of a lane group (more than 8192 haplotypes).
#pragma once
#include "common.cuh"
#include "poly_closed.cuh"

#ifndef LDPGTA_SMALL_CSA_DEPTH
#define LDPGTA_SMALL_CSA_DEPTH 1  // carry-save levels before POPC (see PopcSum32): one level balances ALU and XU pipes (measured)
#endif
#ifndef LDPGTA_SMALL_MINB
#define LDPGTA_SMALL_MINB 4      // k_small_direct: resident 128-thread blocks per SM the register allocation must allow
#endif

namespace ldpgta {

struct SmallArgs {
    PanelView pv;
    const int32_t *idx;      // [n_draws][NR]
    int64_t n_draws;
    const int64_t *pos;      // optional scatter positions of the outputs
    double *out;             // [*][4]
    uint32_t *counts_out;    // optional [n_draws][G][2^NR]
    double *llr;             // optional [5][llr_stride]: the five LLRs of every draw, written with the likelihoods
    int64_t llr_stride;
    int use_gather4;         // 4 reads per draw: one TMA gather4 per draw instead of four bulk copies
    alignas(64) CUtensorMap tmap[2];   // read-mask tensors of groups 0 and 1 (box = one row)
};

// ---- PTX: mbarrier + bulk async copy (TMA 1-D) ---------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return done != 0;
}
// try_wait suspends the warp in hardware for a bounded time per attempt; a copy that never completes (a bad
// tensor map, a fault) must end in an error, not in a hung GPU, so the spin is bounded.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins)
        if (spins > (1u << 24)) __trap();
}
__device__ __forceinline__ void bulk_copy_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// four rows of a 2-D tensor in one instruction (sm_100 TMA tile::gather4); rows land back to back in shared memory
__device__ __forceinline__ void tma_gather4_g2s(void *dst_smem, const CUtensorMap *tmap, int32_t col, int32_t r0, int32_t r1,
                                                int32_t r2, int32_t r3, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                 ::"r"(smem_u32(dst_smem)), "l"(tmap), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(smem_u32(bar)) : "memory");
}
// per-lane asynchronous 16-byte copies global -> shared (LDGSTS): every lane stages exactly the slots it will read
// back itself, so completion is a per-thread cp.async.wait_group and no barrier or elected producer is involved
__device__ __forceinline__ void cp_async16(void *dst_smem, const void *src_gmem)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- the arithmetic shared by both variants --------------------------------------------------
// m[i][j] = 16-byte slot j of read i's mask held by this lane; adds the lane's partial counts of every
// subset of two or more reads to c[].
template <int NR, int WPL>
__device__ __forceinline__ void small_subset_counts(const uint4 (&m)[NR][WPL / 2], uint32_t *c)
{
    constexpr int NS = 1 << NR, SLOTS = WPL / 2;
#pragma unroll
    for (int s = 3; s < NS; ++s) {
        if ((s & (s - 1)) == 0) continue;               // singletons come from the per-read table
        uint32_t t[2 * WPL];
#pragma unroll
        for (int j = 0; j < SLOTS; ++j) {
            uint32_t x = ~0u, y = ~0u, z = ~0u, w = ~0u;
#pragma unroll
            for (int i = 0; i < NR; ++i)
                if ((s >> i) & 1) { x &= m[i][j].x; y &= m[i][j].y; z &= m[i][j].z; w &= m[i][j].w; }
            t[4 * j] = x; t[4 * j + 1] = y; t[4 * j + 2] = z; t[4 * j + 3] = w;
        }
        c[s] += PopcSum32<2 * WPL, LDPGTA_SMALL_CSA_DEPTH>::run(t);
    }
}

// sums c[] over the LPD lanes of a draw; every lane ends with the totals
template <int NR, int LPD>
__device__ __forceinline__ void small_group_sum(uint32_t *c, bool packed)
{
    constexpr int NS = 1 << NR;
    if (packed) {                         // fewer than 65536 haplotypes: two 16-bit counters per shuffle
#pragma unroll
        for (int s = 0; s < NS; s += 2) {
            uint32_t v = c[s] | (c[s + 1] << 16);
#pragma unroll
            for (int msk = LPD / 2; msk; msk >>= 1) v += __shfl_xor_sync(0xffffffffu, v, msk);
            c[s] = v & 0xffffu;
            c[s + 1] = v >> 16;
        }
    } else {
#pragma unroll
        for (int s = 1; s < NS; ++s)
#pragma unroll
            for (int msk = LPD / 2; msk; msk >>= 1) c[s] += __shfl_xor_sync(0xffffffffu, c[s], msk);
    }
}

// Folds one group's counts into the model state and, after the last group, evaluates the polynomials.
template <int NR, int KIND>
struct SmallModel {
    static constexpr int NS = 1 << NR;
    uint32_t c0[KIND != LDPGTA_HOMOGENEOUS ? NS : 1];        // group 0's counts (two-group models finished in batches)
    double f[KIND == LDPGTA_DISTANT_ADMIXTURE ? NS : 1];

    __device__ __forceinline__ void init()
    {
        if (KIND == LDPGTA_DISTANT_ADMIXTURE)
#pragma unroll
            for (int s = 0; s < NS; ++s) f[s] = 0.0;
        if (KIND != LDPGTA_HOMOGENEOUS) c0[0] = 0;
    }
    // two-group closed forms from the counts of both groups
    static __device__ __forceinline__ void finish_recent(const PanelView &pv, const uint32_t *ca, const uint32_t *cb, double *res)
    {
        const double *norm0 = pv.norm[0], *norm1 = pv.norm[1];
        double fr[NS], gr[NS];
#pragma unroll
        for (int s = 1; s < NS; ++s) { fr[s] = __ldg(norm0 + ca[s]); gr[s] = __ldg(norm1 + cb[s]); }
        closed_recent<NR>(fr, gr, res);
    }
    // distant admixture with two groups: e(S) = p_0 cnt_0 / N_0 + p_1 cnt_1 / N_1 (DISTANT_ADMIXTURE_MODELS.py:136-139) as
    // alpha cnt_0 + beta cnt_1 with alpha = p_0 / N_0, beta = p_1 / N_1: two float64 divisions per draw instead of two per
    // subset (30 at four reads: 0.587 -> 0.508 ms per configs[1]-sized step), the scaling the integer form of the larger draws
    // uses too (poly_general.cuh: Cubic2); e(S) differs from the reference's rounding by at most 2 ulp
    static __device__ __forceinline__ void finish_distant2(const PanelView &pv, const uint32_t *ca, const uint32_t *cb, double *res)
    {
        const double alpha = pv.coef[0] / (double)pv.nhap[0], beta = pv.coef[1] / (double)pv.nhap[1];
        double e[NS];
        e[0] = 0.0;
#pragma unroll
        for (int s = 1; s < NS; ++s) e[s] = alpha * (double)ca[s] + beta * (double)cb[s];
        closed_single<NR>(e, res);
    }
    // returns true when res[] is final
    __device__ __forceinline__ bool fold(const PanelView &pv, int g, const uint32_t *c, double *res)
    {
        if (KIND == LDPGTA_HOMOGENEOUS) {
            const double *norm = pv.norm[0];
            double fr[NS];
#pragma unroll
            for (int s = 1; s < NS; ++s) fr[s] = __ldg(norm + c[s]);             // = c[s] / N exactly
            closed_single<NR>(fr, res);
            return true;
        } else if (KIND == LDPGTA_RECENT_ADMIXTURE) {
            if (g == 0) {
#pragma unroll
                for (int s = 1; s < NS; ++s) c0[s] = c[s];
                return false;
            }
            finish_recent(pv, c0, c, res);
            return true;
        } else {
            // effective_joint_frequency, DISTANT_ADMIXTURE_MODELS.py:136-139: sum_g p_g * cnt_g / N_g, in group order
            const double nh = (double)pv.nhap[g], p = pv.coef[g];
#pragma unroll
            for (int s = 1; s < NS; ++s) f[s] = f[s] + p * (double)c[s] / nh;
            if (g + 1 < pv.G) return false;
            closed_single<NR>(f, res);
            return true;
        }
    }
};

// WITH_LLR: also write the draw's five LLRs (k_mid); k_small leaves them to k_draw_llr (see bootstrap_api.cu: llr_fused)
template <bool WITH_LLR = false>
__device__ __forceinline__ void small_store(const SmallArgs &a, int64_t draw, const double *res)
{
    const int64_t o = a.pos ? a.pos[draw] : draw;
    double2 *dst = reinterpret_cast<double2 *>(a.out + 4 * o);
    dst[0] = make_double2(res[0], res[1]);
    dst[1] = make_double2(res[2], res[3]);
    if (WITH_LLR && a.llr) store_llrs(a.llr, a.llr_stride, o, res);
}

template <int NR>
__device__ __forceinline__ void small_dump(const SmallArgs &a, int64_t draw, int g, const uint32_t *c)
{
    uint32_t *o = a.counts_out + (draw * a.pv.G + g) * (1 << NR);
    o[0] = 0;
#pragma unroll
    for (int s = 1; s < (1 << NR); ++s) o[s] = c[s];
}

// ---- pipelined persistent kernel ---------------------------------------------------------------
template <int NR, int WPL>
struct SmallPipe {
    static constexpr int STAGE_BYTES = 256 * NR * WPL;                       // (32/LPD) draws * NR rows * LPD*WPL words * 8 B
    static constexpr int NWARPS_RAW = (220 * 1024) / STAGE_BYTES;
    static constexpr int NWARPS = NWARPS_RAW > 16 ? 16 : (NWARPS_RAW & ~3);     // register file is allocated per 4 warps
    // one-group models: the counts of 32 draws (8 tiles) wait in a per-warp table so that the lookups and the float64
    // polynomial run once per 32 draws with one draw per lane, instead of once per tile with every draw done by 8 lanes
    static constexpr int BATCH_BYTES = 32 * 16 * 4;
    static constexpr int BARS = NWARPS * (STAGE_BYTES + BATCH_BYTES);         // byte offset of the mbarriers
    static constexpr int SMEM = BARS + NWARPS * 16;
};

// CA = true (experiment, LDPGTA_SMALL_CPASYNC=1): the masks are staged by per-lane cp.async (LDGSTS) instead of TMA.  A lane
// copies the 16-byte slots it later reads into registers (slot j * LPD + lane_in of each of its draw's reads): ~80 SASS
// instructions per tile against ~120 for the TMA route (electing a lane, per-draw coordinates into uniform registers,
// arming the mbarrier), yet the kernel is no faster (the AND / carry-save / POPC core is 650 of ~950 instructions per tile
// and the register file, not the issue rate of the staging code, limits the warps in flight).
template <int NR, int LPD, int WPL, int KIND, bool CA = false>
__global__ void __launch_bounds__(SmallPipe<NR, WPL>::NWARPS * 32, 1) k_small(const __grid_constant__ SmallArgs a)
{
    static_assert(WPL % 2 == 0, "rows are read in 16-byte slots");
    using P = SmallPipe<NR, WPL>;
    constexpr int NS = 1 << NR, SLOTS = WPL / 2, DPW = 32 / LPD, ROWS = DPW * NR;
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);        // warp-uniform for the compiler: barrier and buffer addresses live in uniform registers
    const int lane_in = lane % LPD, dgrp = lane / LPD;
    unsigned char *wbase = smem + (size_t)warp * P::STAGE_BYTES;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + P::BARS) + warp;
    // batched polynomial (one-group models): counts[quad][slot] as uint4, slot = tile-in-batch * DPW + draw-in-tile
    uint4 *batch = reinterpret_cast<uint4 *>(smem + (size_t)P::NWARPS * P::STAGE_BYTES + (size_t)warp * P::BATCH_BYTES);
    constexpr bool BATCHED = KIND == LDPGTA_HOMOGENEOUS;
    constexpr int TPB = LPD;                                                  // tiles per batch: TPB * DPW = 32 draws
    // two-group models (recent admixture, distant admixture of two groups): the same with both groups' counts packed as
    // 16-bit halves (groups of at most 65535 haplotypes)
    const bool batched2 = KIND != LDPGTA_HOMOGENEOUS && a.pv.G == 2 && a.pv.nhap[0] <= 65535u && a.pv.nhap[1] <= 65535u;

    if (lane == 0) mbar_init(bar, 1);
    fence_proxy_async();
    __syncwarp();

    const int G = a.pv.G;
    const int64_t n_tiles = (a.n_draws + DPW - 1) / DPW;
    const int64_t tile0 = (int64_t)blockIdx.x * P::NWARPS + warp, tstride = (int64_t)gridDim.x * P::NWARPS;
    // stages are (tile, group) pairs in order
    const int64_t my_tiles = tile0 < n_tiles ? (n_tiles - tile0 + tstride - 1) / tstride : 0;
    const int64_t n_stages = my_tiles * G;

    // producer lanes: lane l < ROWS owns row (draw l / NR, read l % NR) of every tile
    auto load_idx = [&](int64_t tile) -> int32_t {
        if (lane >= ROWS) return 0;
        int64_t d = tile * DPW + lane / NR;
        if (d >= a.n_draws) d = a.n_draws - 1;                                 // ragged last tile: harmless duplicates
        return __ldg(a.idx + d * NR + lane % NR);
    };
    auto issue = [&](int g_, int32_t ridx) -> uint32_t {
        const int g = KIND == LDPGTA_HOMOGENEOUS ? 0 : g_;
        const int wp = a.pv.wp[g];
        const uint32_t row_bytes = (uint32_t)wp * 8u;
        uint32_t pc = 0;
#ifdef LDPGTA_SMALL_COMPUTE_ONLY
        // measurement build (tools/gather_roof.sh): no copies at all, the arithmetic runs on whatever the buffers hold
        if (lane == 0) mbar_arrive_expect_tx(bar, 0);
        __syncwarp();
        return pc;
#endif
        if (lane == 0) mbar_arrive_expect_tx(bar, row_bytes * ROWS);
        __syncwarp();
        if (NR == 4 && a.use_gather4) {
            // lane 4d holds read 0 of draw d; its neighbours hold reads 1..3
            const int32_t r1 = __shfl_down_sync(0xffffffffu, ridx, 1), r2 = __shfl_down_sync(0xffffffffu, ridx, 2),
                          r3 = __shfl_down_sync(0xffffffffu, ridx, 3);
            if (lane < ROWS && (lane & 3) == 0)
                tma_gather4_g2s(wbase + (size_t)lane * row_bytes, &a.tmap[g], 0, ridx, r1, r2, r3, bar);
        } else if (lane < ROWS) {
            bulk_copy_g2s(wbase + (size_t)lane * row_bytes, a.pv.masks[g] + (int64_t)ridx * wp, row_bytes, bar);
        }
        if (lane < ROWS) pc = __ldg(a.pv.popc[g] + ridx);
        return pc;
    };

    // (tile, group) of the stage being consumed (cur) and of the stage in flight / next to issue (nxt)
    int64_t cur_tile = tile0, nxt_tile = tile0;
    int cur_g = 0, nxt_g = 0;
    auto advance = [&](int64_t &tile, int &g) { if (++g == G) { g = 0; tile += tstride; } };

    // ---- CA staging: every lane handles the NR reads of its own draw ----
    const bool idx_vec = (reinterpret_cast<uintptr_t>(a.idx) & 15) == 0;
    int32_t ca_idx[NR];                      // read rows of this lane's draw in the next stage to issue
    uint32_t ca_pc_nxt[NR];                  // their popcounts (stage in flight)
    auto ca_load_idx = [&](int64_t tile) {
        int64_t d = tile * DPW + dgrp;
        if (d >= a.n_draws) d = a.n_draws - 1;
        if (NR == 4 && idx_vec) {
            const int4 v = __ldg(reinterpret_cast<const int4 *>(a.idx) + d);
            ca_idx[0] = v.x; ca_idx[1] = v.y; ca_idx[2 % NR] = v.z; ca_idx[3 % NR] = v.w;
        } else {
#pragma unroll
            for (int i = 0; i < NR; ++i) ca_idx[i] = __ldg(a.idx + d * NR + i);
        }
    };
    const uint32_t ca_dst0 = smem_u32(wbase) + (uint32_t)(dgrp * NR) * (uint32_t)(LPD * WPL * 8) + (uint32_t)lane_in * 16u;   // exact rows
    auto ca_issue = [&](int g_) {
        const int g = KIND == LDPGTA_HOMOGENEOUS ? 0 : g_;
        const int wp = a.pv.wp[g];
        const uint64_t *masks = a.pv.masks[g];
        const uint32_t *popc = a.pv.popc[g];
        if (wp == LPD * WPL) {
            // rows fill the lane grid exactly: slot j of read i sits j * LPD * 16 bytes after the lane's first slot
#pragma unroll
            for (int i = 0; i < NR; ++i) {
                const char *src = reinterpret_cast<const char *>(masks + (int64_t)ca_idx[i] * (LPD * WPL)) + lane_in * 16;
#pragma unroll
                for (int j = 0; j < SLOTS; ++j)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ca_dst0 + (uint32_t)(i * LPD * WPL * 8 + j * LPD * 16)),
                                 "l"(src + j * LPD * 16) : "memory");
                ca_pc_nxt[i] = __ldg(popc + ca_idx[i]);
            }
        } else {
            unsigned char *st = wbase + (size_t)(dgrp * NR) * wp * 8;
#pragma unroll
            for (int i = 0; i < NR; ++i) {
                const uint64_t *src = masks + (int64_t)ca_idx[i] * wp;
#pragma unroll
                for (int j = 0; j < SLOTS; ++j) {
                    const int w = (j * LPD + lane_in) * 2;
                    if (w < wp) cp_async16(st + ((size_t)i * wp + w) * 8, src + w);
                }
                ca_pc_nxt[i] = __ldg(popc + ca_idx[i]);
            }
        }
        cp_async_commit();
    };

    // prologue: stage 0 in flight, indices of stage 1 loaded
    uint32_t pc_cur = 0;                     // row popcounts of the stage in the buffer
    int32_t idx_nxt = 0;                     // row index of the next stage to issue
    if (CA) {
        if (n_stages > 0) { ca_load_idx(nxt_tile); ca_issue(nxt_g); advance(nxt_tile, nxt_g); }
        if (n_stages > 1 && nxt_g == 0) ca_load_idx(nxt_tile);
    } else {
        if (n_stages > 0) { pc_cur = issue(nxt_g, load_idx(nxt_tile)); advance(nxt_tile, nxt_g); }
        if (n_stages > 1) idx_nxt = load_idx(nxt_tile);
    }

    SmallModel<NR, KIND> model;
    uint32_t parity = 0;
    for (int64_t q = 0; q < n_stages; ++q) {
        const int g = KIND == LDPGTA_HOMOGENEOUS ? 0 : cur_g;
        const int64_t draw = cur_tile * DPW + dgrp;
        const bool writer = draw < a.n_draws && lane_in == 0;
        const int wp = a.pv.wp[g];
        if (g == 0) model.init();
        advance(cur_tile, cur_g);

        // single-read counts of this stage (prefetched with the indices)
        uint32_t c[NS];
#pragma unroll
        for (int s = 0; s < NS; ++s) c[s] = 0;
        if (CA) {
#pragma unroll
            for (int i = 0; i < NR; ++i) if (lane_in == 0) c[1 << i] = ca_pc_nxt[i];
            cp_async_wait_all();
        } else {
#pragma unroll
            for (int i = 0; i < NR; ++i) {
                const uint32_t v = __shfl_sync(0xffffffffu, pc_cur, dgrp * NR + i);
                if (lane_in == 0) c[1 << i] = v;
            }
            mbar_wait(bar, parity);
            parity ^= 1u;
        }

        uint4 m[NR][SLOTS];
        {
            const unsigned char *st = wbase + (size_t)(dgrp * NR) * wp * 8;
            if (wp == LPD * WPL) {                          // rows fill the lane grid exactly: no bounds predicates
#pragma unroll
                for (int j = 0; j < SLOTS; ++j)
#pragma unroll
                    for (int i = 0; i < NR; ++i)
                        m[i][j] = *reinterpret_cast<const uint4 *>(st + ((size_t)i * (LPD * WPL) + (j * LPD + lane_in) * 2) * 8);
            } else {
#pragma unroll
                for (int j = 0; j < SLOTS; ++j) {
                    const int w = (j * LPD + lane_in) * 2;
                    const bool in = w < wp;
#pragma unroll
                    for (int i = 0; i < NR; ++i)
                        m[i][j] = in ? *reinterpret_cast<const uint4 *>(st + ((size_t)i * wp + w) * 8) : make_uint4(0, 0, 0, 0);
                }
            }
        }
        // the buffer is free again: refill it with the next stage, which lands while this one is computed
        __syncwarp();
        if (CA) {
            if (q + 1 < n_stages) {
                ca_issue(nxt_g);                                   // same draw indices for every group of a tile
                advance(nxt_tile, nxt_g);
                if (q + 2 < n_stages && nxt_g == 0) ca_load_idx(nxt_tile);
            }
        } else if (q + 1 < n_stages) {
            fence_proxy_async();
            pc_cur = issue(nxt_g, idx_nxt);
            advance(nxt_tile, nxt_g);
            if (q + 2 < n_stages) idx_nxt = load_idx(nxt_tile);
        }

#ifdef LDPGTA_SMALL_GATHER_ONLY
        // measurement build (tools/gather_roof.sh): the staging pipeline and the register loads without the AND / POPC work --
        // how fast the masks of this draw pattern can be gathered at all
#pragma unroll
        for (int i = 0; i < NR; ++i)
#pragma unroll
            for (int j = 0; j < SLOTS; ++j) c[3] += m[i][j].x ^ m[i][j].y ^ m[i][j].z ^ m[i][j].w;
#else
        small_subset_counts<NR, WPL>(m, c);
#endif
#ifndef LDPGTA_SMALL_NO_PACKED_BATCH
        if (BATCHED && NS >= 8 && a.pv.nhap[0] <= 65535u && !a.counts_out) {
            // one-group model, counts below 2^16: two counters per register from the lane partials on -- through the
            // group's xor-shuffles and the per-warp batch table -- and unpacked once per draw by the lane that finishes it
            uint32_t pk[NS / 2];
            const int k = (int)(q % TPB);
            // as many packed counters as lanes in the group (4 reads, 8 lanes): a transposing reduction -- every step halves
            // the counters a lane carries, 7 shuffles instead of 24 -- leaves counter i summed in lane i, which stores it
            #ifdef LDPGTA_SMALL_BUTTERFLY
            constexpr bool TRANSPOSE = false;                               // A/B build: every counter through the full butterfly
#else
            constexpr bool TRANSPOSE = NS / 2 == LPD && LPD == 8;
#endif
            constexpr int PLANE = TRANSPOSE ? 36 : 32;                      // uint4 stride of a table plane (+16 words: the 32 stores hit 32 banks)
            if (TRANSPOSE) {
#pragma unroll
                for (int t = 0; t < NS / 2; ++t) pk[t] = c[2 * t] + (c[2 * t + 1] << 16);
#pragma unroll
                for (int h = LPD / 2; h; h >>= 1) {
                    const bool up = (lane_in & h) != 0;
#pragma unroll
                    for (int t = 0; t < h; ++t) {
                        const uint32_t send = up ? pk[t] : pk[t + h], keep = up ? pk[t + h] : pk[t];
                        pk[t] = keep + __shfl_xor_sync(0xffffffffu, send, h);
                    }
                }
                const int slot = k * DPW + dgrp;
                reinterpret_cast<uint32_t *>(batch)[(lane_in >> 2) * (PLANE * 4) + slot * 4 + (lane_in & 3)] = pk[0];
            } else {
#pragma unroll
                for (int t = 0; t < NS / 2; ++t) {
                    pk[t] = c[2 * t] + (c[2 * t + 1] << 16);
#pragma unroll
                    for (int msk = LPD / 2; msk; msk >>= 1) pk[t] += __shfl_xor_sync(0xffffffffu, pk[t], msk);
                }
                if (lane_in == 0) {
                    const int slot = k * DPW + dgrp;
#pragma unroll
                    for (int j = 0; j < NS / 8; ++j) batch[j * 32 + slot] = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
                }
            }
            if (k == TPB - 1 || q + 1 == n_stages) {
                __syncwarp();
                uint32_t cc[NS];
#pragma unroll
                for (int j = 0; j < NS / 8; ++j) {
                    const uint4 v = batch[j * PLANE + lane];
                    const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) { cc[8 * j + 2 * e] = w4[e] & 0xffffu; cc[8 * j + 2 * e + 1] = w4[e] >> 16; }
                }
                const int kk = lane / DPW;
                const int64_t draw_k = (tile0 + (q - k + kk) * tstride) * DPW + lane % DPW;
                double res[4];
                if (kk <= k) {
                    model.fold(a.pv, 0, cc, res);
                    if (draw_k < a.n_draws) small_store(a, draw_k, res);
                }
                __syncwarp();
            }
            continue;
        }
#endif
        small_group_sum<NR, LPD>(c, a.pv.nhap[g] <= 65535u);
        if (a.counts_out && writer) small_dump<NR>(a, draw, g, c);
        if (BATCHED) {
            // park this tile's counts; every TPB tiles (or at the end) each lane finishes one draw
            const int k = (int)(q % TPB);
            if (lane_in == 0) {
                const int slot = k * DPW + dgrp;
#pragma unroll
                for (int j = 0; j < NS / 4; ++j) batch[j * 32 + slot] = make_uint4(c[4 * j], c[4 * j + 1], c[4 * j + 2], c[4 * j + 3]);
            }
            if (k == TPB - 1 || q + 1 == n_stages) {
                __syncwarp();
                uint32_t cc[NS];
#pragma unroll
                for (int j = 0; j < NS / 4; ++j) {
                    const uint4 v = batch[j * 32 + lane];
                    cc[4 * j] = v.x; cc[4 * j + 1] = v.y; cc[4 * j + 2] = v.z; cc[4 * j + 3] = v.w;
                }
                const int kk = lane / DPW;                                     // which tile of the batch this lane finishes
                const int64_t tile_k = tile0 + (q - k + kk) * tstride;
                const int64_t draw_k = tile_k * DPW + lane % DPW;
                double res[4];
                if (kk <= k) {
                    model.fold(a.pv, 0, cc, res);
                    if (draw_k < a.n_draws) small_store(a, draw_k, res);
                }
                __syncwarp();
            }
        } else if (KIND != LDPGTA_HOMOGENEOUS && batched2) {
            double res[4];
            if (g == 0) {                                                      // keep group 0's counts
#pragma unroll
                for (int s = 1; s < NS; ++s) model.c0[s] = c[s];
            } else {
                const int k = (int)((q >> 1) % TPB);
                if (lane_in == 0) {
                    const int slot = k * DPW + dgrp;
#pragma unroll
                    for (int j = 0; j < NS / 4; ++j)
                        batch[j * 32 + slot] = make_uint4(model.c0[4 * j] | (c[4 * j] << 16), model.c0[4 * j + 1] | (c[4 * j + 1] << 16),
                                                          model.c0[4 * j + 2] | (c[4 * j + 2] << 16), model.c0[4 * j + 3] | (c[4 * j + 3] << 16));
                }
                if (k == TPB - 1 || q + 1 == n_stages) {
                    __syncwarp();
                    uint32_t ca[NS], cb[NS];
#pragma unroll
                    for (int j = 0; j < NS / 4; ++j) {
                        const uint4 v = batch[j * 32 + lane];
                        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) { ca[4 * j + e] = w[e] & 0xffffu; cb[4 * j + e] = w[e] >> 16; }
                    }
                    const int kk = lane / DPW;
                    const int64_t draw_k = (tile0 + ((q >> 1) - k + kk) * tstride) * DPW + lane % DPW;
                    if (kk <= k) {
                        if (KIND == LDPGTA_RECENT_ADMIXTURE) SmallModel<NR, KIND>::finish_recent(a.pv, ca, cb, res);
                        else SmallModel<NR, KIND>::finish_distant2(a.pv, ca, cb, res);
                        if (draw_k < a.n_draws) small_store(a, draw_k, res);
                    }
                    __syncwarp();
                }
            }
        } else {
            double res[4];
            if (model.fold(a.pv, g, c, res) && writer) small_store(a, draw, res);
        }
    }
}

// ---- two groups in one row --------------------------------------------------------------------------------------
// Recent admixture and distant admixture of two groups at 4 reads per draw gathered their masks group by group: two stages
// per tile of 320-byte rows, which the memory system serves worse than the 640-byte rows of a one-group panel of the same
// total size (gather-only builds: 0.395 against 0.317 ms per configs[1]-sized step, tools/gather_roof.sh).  Here both groups'
// masks of a read sit in ONE row [group 0 | group 1] of a combined table (built once on the device, small.cu: ensure_both),
// so a tile is one stage of the one-group shape <4, 8, 10>.  The intersections are formed on the whole row; the words of a
// subset are popcounted in two carry-save trees, group 0 = slots below `split` (16-byte units), group 1 = the others -- the
// boundary falls inside slot row SJ, where lanes below `lsplit` belong to group 0, so that row enters both trees masked --
// and the two counts travel as the 16-bit halves of one register through the group's shuffles into the batched finish.
struct SmallBothArgs {
    SmallArgs base;          // pv / idx / n_draws / pos / out as in k_small; masks and tensor map of group 0 are not used
    const uint64_t *both;    // [n_reads][80] combined rows
    int lsplit;              // lanes (of the 8 of a draw) whose slot in row SJ belongs to group 0
    alignas(64) CUtensorMap tmap_both;
};

template <int KIND, int SJ>
__global__ void __launch_bounds__(SmallPipe<4, 10>::NWARPS * 32, 1) k_small_both(const __grid_constant__ SmallBothArgs b)
{
    constexpr int NR = 4, LPD = 8, WPL = 10, NS = 16, SLOTS = 5, DPW = 4, ROWS = 16, TPB = LPD;
    using P = SmallPipe<NR, WPL>;
    const SmallArgs &a = b.base;
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane_in = lane % LPD, dgrp = lane / LPD;
    unsigned char *wbase = smem + (size_t)warp * P::STAGE_BYTES;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + P::BARS) + warp;
    uint4 *batch = reinterpret_cast<uint4 *>(smem + (size_t)P::NWARPS * P::STAGE_BYTES + (size_t)warp * P::BATCH_BYTES);
    if (lane == 0) mbar_init(bar, 1);
    fence_proxy_async();
    __syncwarp();

    const int64_t n_tiles = (a.n_draws + DPW - 1) / DPW;
    const int64_t tile0 = (int64_t)blockIdx.x * P::NWARPS + warp, tstride = (int64_t)gridDim.x * P::NWARPS;
    const int64_t n_stages = tile0 < n_tiles ? (n_tiles - tile0 + tstride - 1) / tstride : 0;
    constexpr uint32_t ROW_BYTES = LPD * WPL * 8;

    auto load_idx = [&](int64_t tile) -> int32_t {
        if (lane >= ROWS) return 0;
        int64_t d = tile * DPW + lane / NR;
        if (d >= a.n_draws) d = a.n_draws - 1;
        return __ldg(a.idx + d * NR + lane % NR);
    };
    // returns the row's two popcounts packed (group 0 | group 1 << 16)
    auto issue = [&](int32_t ridx) -> uint32_t {
        if (lane == 0) mbar_arrive_expect_tx(bar, ROW_BYTES * ROWS);
        __syncwarp();
        if (a.use_gather4) {
            const int32_t r1 = __shfl_down_sync(0xffffffffu, ridx, 1), r2 = __shfl_down_sync(0xffffffffu, ridx, 2),
                          r3 = __shfl_down_sync(0xffffffffu, ridx, 3);
            if (lane < ROWS && (lane & 3) == 0) tma_gather4_g2s(wbase + (size_t)lane * ROW_BYTES, &b.tmap_both, 0, ridx, r1, r2, r3, bar);
        } else if (lane < ROWS) {
            bulk_copy_g2s(wbase + (size_t)lane * ROW_BYTES, b.both + (int64_t)ridx * (LPD * WPL), ROW_BYTES, bar);
        }
        return lane < ROWS ? (__ldg(a.pv.popc[0] + ridx) | (__ldg(a.pv.popc[1] + ridx) << 16)) : 0u;
    };

    uint32_t pc_cur = 0;
    int32_t idx_nxt = 0;
    int64_t nxt_tile = tile0;
    if (n_stages > 0) { pc_cur = issue(load_idx(nxt_tile)); nxt_tile += tstride; }
    if (n_stages > 1) idx_nxt = load_idx(nxt_tile);
    const uint32_t mA = lane_in < b.lsplit ? ~0u : 0u;       // slot row SJ: group 0 for these lanes, group 1 for the others

    uint32_t parity = 0;
    for (int64_t q = 0; q < n_stages; ++q) {
        // pk[s] = count of subset s in group 0 | group 1 << 16, this lane's share
        uint32_t pk[NS];
#pragma unroll
        for (int s = 0; s < NS; ++s) pk[s] = 0;
#pragma unroll
        for (int i = 0; i < NR; ++i) {
            const uint32_t v = __shfl_sync(0xffffffffu, pc_cur, dgrp * NR + i);
            if (lane_in == 0) pk[1 << i] = v;
        }
        mbar_wait(bar, parity);
        parity ^= 1u;
        uint4 m[NR][SLOTS];
        {
            const unsigned char *st = wbase + (size_t)(dgrp * NR) * ROW_BYTES;
#pragma unroll
            for (int j = 0; j < SLOTS; ++j)
#pragma unroll
                for (int i = 0; i < NR; ++i)
                    m[i][j] = *reinterpret_cast<const uint4 *>(st + ((size_t)i * (LPD * WPL) + (j * LPD + lane_in) * 2) * 8);
        }
        __syncwarp();
        if (q + 1 < n_stages) {
            fence_proxy_async();
            pc_cur = issue(idx_nxt);
            nxt_tile += tstride;
            if (q + 2 < n_stages) idx_nxt = load_idx(nxt_tile);
        }
#pragma unroll
        for (int s = 3; s < NS; ++s) {
            if ((s & (s - 1)) == 0) continue;
            uint32_t ta[4 * (SJ + 1)], tb[4 * (SLOTS - SJ)];
#pragma unroll
            for (int j = 0; j < SLOTS; ++j) {
                uint32_t x = ~0u, y = ~0u, z = ~0u, w = ~0u;
#pragma unroll
                for (int i = 0; i < NR; ++i)
                    if ((s >> i) & 1) { x &= m[i][j].x; y &= m[i][j].y; z &= m[i][j].z; w &= m[i][j].w; }
                if (j < SJ) { ta[4 * j] = x; ta[4 * j + 1] = y; ta[4 * j + 2] = z; ta[4 * j + 3] = w; }
                else if (j > SJ) { const int k = j - SJ; tb[4 * k] = x; tb[4 * k + 1] = y; tb[4 * k + 2] = z; tb[4 * k + 3] = w; }
                else {
                    ta[4 * j] = x & mA; ta[4 * j + 1] = y & mA; ta[4 * j + 2] = z & mA; ta[4 * j + 3] = w & mA;
                    tb[0] = x & ~mA; tb[1] = y & ~mA; tb[2] = z & ~mA; tb[3] = w & ~mA;
                }
            }
            pk[s] = PopcSum32<4 * (SJ + 1), LDPGTA_SMALL_CSA_DEPTH>::run(ta) + (PopcSum32<4 * (SLOTS - SJ), LDPGTA_SMALL_CSA_DEPTH>::run(tb) << 16);
        }
#pragma unroll
        for (int s = 1; s < NS; ++s)
#pragma unroll
            for (int msk = LPD / 2; msk; msk >>= 1) pk[s] += __shfl_xor_sync(0xffffffffu, pk[s], msk);
        // park the tile's counts; every TPB tiles (or at the end) each lane finishes one draw
        const int k = (int)(q % TPB);
        if (lane_in == 0) {
            const int slot = k * DPW + dgrp;
#pragma unroll
            for (int j = 0; j < NS / 4; ++j) batch[j * 32 + slot] = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
        }
        if (k == TPB - 1 || q + 1 == n_stages) {
            __syncwarp();
            uint32_t ca[NS], cb[NS];
#pragma unroll
            for (int j = 0; j < NS / 4; ++j) {
                const uint4 v = batch[j * 32 + lane];
                const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) { ca[4 * j + e] = w[e] & 0xffffu; cb[4 * j + e] = w[e] >> 16; }
            }
            const int kk = lane / DPW;
            const int64_t draw_k = (tile0 + (q - k + kk) * tstride) * DPW + lane % DPW;
            if (kk <= k) {
                double res[4];
                if (KIND == LDPGTA_RECENT_ADMIXTURE) SmallModel<NR, KIND>::finish_recent(a.pv, ca, cb, res);
                else SmallModel<NR, KIND>::finish_distant2(a.pv, ca, cb, res);
                if (draw_k < a.n_draws) small_store(a, draw_k, res);
            }
            __syncwarp();
        }
    }
}

// ---- window-resident kernel: many draws per window (e.g. 1000 bootstrap subsamples) --------------------------------
// With hundreds of draws per window every read of the window is drawn many times; gathering the masks of each draw again
// costs L2 -> shared-memory traffic and ~120 staging instructions per tile for rows that were on the SM a microsecond ago.
// Here a persistent CTA takes a whole window: ONE bulk copy brings the window's R consecutive read masks (R * 640 bytes) and
// its R popcounts into shared memory, then the 16 warps walk the window's draws in tiles of four, every lane loading its
// slots of its draw's four rows straight from the resident block into registers (the LDS that k_small needs anyway) -- no
// per-tile copies, barriers or popcount loads.  Arithmetic, packed counters, transposing reduction and batched finish are
// those of k_small<4, 8, 10, HOMOGENEOUS>.  Windows are runs of consecutive rows: of the read-mask table itself, or -- for
// windows given as read lists -- of a window-ordered copy of the masks that the plan keeps (bootstrap_api.cu: ensure_wmasks),
// with the sampler's positions inside the window as draw indices.
struct SmallWinArgs {
    SmallArgs base;          // pv / idx [n_draws][4] / out; pos, counts_out, llr are not used
    const uint64_t *masks;   // the table the windows are runs of: the context's read masks, or the plan's window-ordered copy
    const uint32_t *popc;    // its row popcounts
    int local;               // idx holds positions inside the window (window given as a read list) instead of rows of the table
    const int64_t *woff;     // [n_windows + 1] first row of each window
    const int64_t *doff;     // [n_windows + 1] first draw of each window
    int64_t n_windows;
    int rcap;                // rows the shared-memory block holds
};

struct SmallWinSmem {
    static constexpr int NWARPS = 16, BATCH_BYTES = 36 * 16 * 2 + 128;        // two planes of the transposed table (see k_small)
    static constexpr int bytes(int rcap) { return rcap * 640 + ((rcap * 4 + 127) & ~127) + NWARPS * BATCH_BYTES + 64; }
};

static __global__ void __launch_bounds__(SmallWinSmem::NWARPS * 32, 1) k_small_win(const __grid_constant__ SmallWinArgs w)
{
    constexpr int NR = 4, LPD = 8, WPL = 10, NS = 16, SLOTS = 5, DPW = 4, TPB = 8, PLANE = 36, ROW_BYTES = 640;
    const SmallArgs &a = w.base;
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane_in = lane % LPD, dgrp = lane / LPD;
    unsigned char *rows = smem;
    uint32_t *pcs = reinterpret_cast<uint32_t *>(smem + (size_t)w.rcap * ROW_BYTES);
    unsigned char *after = reinterpret_cast<unsigned char *>(pcs) + ((w.rcap * 4 + 127) & ~127);
    uint4 *batch = reinterpret_cast<uint4 *>(after + (size_t)warp * SmallWinSmem::BATCH_BYTES);
    uint64_t *bar = reinterpret_cast<uint64_t *>(after + (size_t)SmallWinSmem::NWARPS * SmallWinSmem::BATCH_BYTES);
    if (threadIdx.x == 0) mbar_init(bar, 1);
    fence_proxy_async();
    __syncthreads();
    const double *norm = a.pv.norm[0];
    const int4 *idx4 = reinterpret_cast<const int4 *>(a.idx);
    uint32_t parity = 0;
    for (int64_t win = blockIdx.x; win < w.n_windows; win += gridDim.x) {
        const int64_t r0 = w.woff[win], d0 = w.doff[win];
        const int R = (int)(w.woff[win + 1] - r0);
        const int64_t D = w.doff[win + 1] - d0;
        // the window's masks: one bulk copy; its popcounts: plain loads
        if (threadIdx.x == 0) {
            fence_proxy_async();                                              // the block was read through the generic proxy until the barrier above
            mbar_arrive_expect_tx(bar, (uint32_t)R * ROW_BYTES);
            bulk_copy_g2s(rows, w.masks + r0 * (LPD * WPL), (uint32_t)R * ROW_BYTES, bar);
        }
        for (int i = threadIdx.x; i < R; i += blockDim.x) pcs[i] = __ldg(w.popc + r0 + i);
        const int64_t n_tiles = (D + DPW - 1) / DPW;
        const int64_t my_tiles = warp < n_tiles ? (n_tiles - warp + SmallWinSmem::NWARPS - 1) / SmallWinSmem::NWARPS : 0;
        // the draw of this lane's group in the warp's first tile (prefetched one tile ahead)
        auto load_draw = [&](int64_t q) -> int4 {
            int64_t d = (warp + q * SmallWinSmem::NWARPS) * DPW + dgrp;
            if (d >= D) d = D - 1;                                            // ragged last tile: harmless duplicates
            return __ldg(idx4 + d0 + d);
        };
        int4 nxt = my_tiles > 0 ? load_draw(0) : make_int4(0, 0, 0, 0);
        mbar_wait(bar, parity);
        parity ^= 1u;
        __syncthreads();                                                      // popcounts visible
        for (int64_t q = 0; q < my_tiles; ++q) {
            const int4 cur = nxt;
            if (q + 1 < my_tiles) nxt = load_draw(q + 1);
            const int sub = w.local ? 0 : (int)r0;
            const int lr[NR] = {cur.x - sub, cur.y - sub, cur.z - sub, cur.w - sub};
            uint32_t c[NS];
#pragma unroll
            for (int s = 0; s < NS; ++s) c[s] = 0;
#pragma unroll
            for (int i = 0; i < NR; ++i) if (lane_in == 0) c[1 << i] = pcs[lr[i]];
            uint4 m[NR][SLOTS];
#pragma unroll
            for (int j = 0; j < SLOTS; ++j)
#pragma unroll
                for (int i = 0; i < NR; ++i)
                    m[i][j] = *reinterpret_cast<const uint4 *>(rows + (size_t)lr[i] * ROW_BYTES + (j * LPD + lane_in) * 16);
            small_subset_counts<NR, WPL>(m, c);
            uint32_t pk[NS / 2];
#pragma unroll
            for (int t = 0; t < NS / 2; ++t) pk[t] = c[2 * t] + (c[2 * t + 1] << 16);
#pragma unroll
            for (int h = LPD / 2; h; h >>= 1) {
                const bool up = (lane_in & h) != 0;
#pragma unroll
                for (int t = 0; t < h; ++t) {
                    const uint32_t send = up ? pk[t] : pk[t + h], keep = up ? pk[t + h] : pk[t];
                    pk[t] = keep + __shfl_xor_sync(0xffffffffu, send, h);
                }
            }
            const int k = (int)(q % TPB);
            reinterpret_cast<uint32_t *>(batch)[(lane_in >> 2) * (PLANE * 4) + (k * DPW + dgrp) * 4 + (lane_in & 3)] = pk[0];
            if (k == TPB - 1 || q + 1 == my_tiles) {
                __syncwarp();
                uint32_t cc[NS];
#pragma unroll
                for (int j = 0; j < NS / 8; ++j) {
                    const uint4 v = batch[j * PLANE + lane];
                    const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) { cc[8 * j + 2 * e] = w4[e] & 0xffffu; cc[8 * j + 2 * e + 1] = w4[e] >> 16; }
                }
                const int kk = lane / DPW;
                const int64_t draw_k = (warp + (q - k + kk) * SmallWinSmem::NWARPS) * DPW + lane % DPW;      // within the window
                if (kk <= k && draw_k < D) {
                    double fr[NS], res[4];
#pragma unroll
                    for (int s = 1; s < NS; ++s) fr[s] = __ldg(norm + cc[s]);
                    closed_single<NR>(fr, res);
                    double2 *dst = reinterpret_cast<double2 *>(a.out + 4 * (d0 + draw_k));
                    dst[0] = make_double2(res[0], res[1]);
                    dst[1] = make_double2(res[2], res[3]);
                }
                __syncwarp();
            }
        }
        __syncthreads();                                                      // every warp is done with the block before it is refilled
    }
}

// ---- direct kernel (no staging): rows longer than one pass of a lane group ---------------------
template <int NR, int LPD, int WPL, int KIND>
__global__ void __launch_bounds__(128, LDPGTA_SMALL_MINB) k_small_direct(const SmallArgs a)
{
    constexpr int NS = 1 << NR, SLOTS = WPL / 2;
    const int lane_in = threadIdx.x % LPD;
    int64_t draw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LPD;
    const bool live = draw < a.n_draws;
    if (!live) draw = a.n_draws - 1;        // keep the lanes in the shuffles
    const bool writer = live && lane_in == 0;

    int32_t r[NR];
#pragma unroll
    for (int i = 0; i < NR; ++i) r[i] = __ldg(a.idx + draw * NR + i);

    SmallModel<NR, KIND> model;
    model.init();
    for (int g = 0; g < a.pv.G; ++g) {
        const int wp = a.pv.wp[g];
        const uint64_t *row[NR];
#pragma unroll
        for (int i = 0; i < NR; ++i) row[i] = a.pv.masks[g] + (int64_t)r[i] * wp;
        uint32_t c[NS];
#pragma unroll
        for (int s = 0; s < NS; ++s) c[s] = 0;
        for (int w0 = 0; w0 < wp; w0 += LPD * WPL) {
            uint4 m[NR][SLOTS];
#pragma unroll
            for (int j = 0; j < SLOTS; ++j) {
                const int w = w0 + (j * LPD + lane_in) * 2;
                const bool in = w < wp;
#pragma unroll
                for (int i = 0; i < NR; ++i) m[i][j] = in ? ldg128(row[i] + w) : make_uint4(0, 0, 0, 0);
            }
            small_subset_counts<NR, WPL>(m, c);
        }
        if (lane_in == 0) {
#pragma unroll
            for (int i = 0; i < NR; ++i) c[1 << i] = __ldg(a.pv.popc[g] + r[i]);
        }
        small_group_sum<NR, LPD>(c, a.pv.nhap[g] <= 65535u);
        if (a.counts_out && writer) small_dump<NR>(a, draw, g, c);
        double res[4];
        if (model.fold(a.pv, g, c, res) && writer) small_store(a, draw, res);
    }
}

}  // namespace ldpgta
