// k_mid.cuh -- draws of 5 (or 6) reads, single-population model: the k_small pipeline with the general
// partition polynomial evaluated from registers.
//
// 16 lanes own a draw (a warp: 2 draws); each lane keeps W64 64-bit words of every read mask in registers
// (lane l takes words j*16 + l), forms the 26 / 57 intersections of two or more reads directly from the
// masks (a LOP3 ANDs up to three of them), popcounts a subset's words through one carry-save level, and the
// 16-bit packed partial counts are added across the 16 lanes with xor-shuffles.  Masks arrive through the same
// single-buffer bulk-copy (TMA 1-D) pipeline as in k_small.  The polynomial is the integer form of
// likelihoods() (HOMOGENOUES_MODELS.py:185-225): exact 64-bit sums over the 15 / 31 two-block and 25 / 90
// three-block partitions, enumerated at compile time, then the reference's few float operations; it runs once per
// 32 draws with one draw per lane (the counts of 16 tiles wait in a per-warp shared-memory table).
// The table kernel (k_general) needs ~6400 warp instructions per 6-read draw; this one ~1300.
// Serves 5 reads of one-group models; from 6 reads on k_lanes is faster (measured).
#pragma once
#include "common.cuh"
#include "k_small.cuh"

#ifndef LDPGTA_MID6_WARPS
#define LDPGTA_MID6_WARPS 12     // measured: 8 -> 2.99e8, 12 -> 3.20e8, 16 (spilling) -> 3.16e8 draws/s
#endif

namespace ldpgta {

template <int NR, int W64>
struct MidPipe {
    static constexpr int STAGE_BYTES = 2 * NR * 16 * W64 * 8;                  // 2 draws * NR rows * 16 lanes * W64 words
    static constexpr int NWARPS_RAW = (200 * 1024) / STAGE_BYTES;
    static constexpr int NWARPS_FIT = NWARPS_RAW > 16 ? 16 : (NWARPS_RAW & ~3);
    // 6 reads keep 60 mask registers and 64 counts live: 12 warps (168 registers each) instead of 16 (128, spilling).
    // (A 7-read instance was measured too: 255 registers with spills, half the speed of the table kernel.)
    static constexpr int NWARPS = (NR == 6 && NWARPS_FIT > LDPGTA_MID6_WARPS) ? LDPGTA_MID6_WARPS : NWARPS_FIT;
    // the counts of 32 draws (16 tiles) wait in a per-warp table; the partition polynomial then runs once per 32 draws,
    // one draw per lane, instead of once per tile with each draw done by 16 lanes
    static constexpr int BATCH_BYTES = 32 * (1 << NR) * 4;
    static constexpr int BARS = NWARPS * (STAGE_BYTES + BATCH_BYTES);
    static constexpr int SMEM = BARS + NWARPS * 16;
};

// exact integer partition sums from the 2^NR counts (index = subset bitmask)
template <int NR>
__device__ __forceinline__ void mid_polynomial(const uint32_t *c, uint32_t nhap, double *res)
{
    constexpr uint32_t FULL = (1u << NR) - 1u, TOP = 1u << (NR - 1);
    u64 two = 0, two_sph = 0, three = 0;
#pragma unroll
    for (uint32_t A = TOP; A < FULL; ++A) {
        const uint32_t rest = FULL ^ A;
        const u64 prod = (u64)c[A] * (u64)c[rest];
        two += prod;
        two_sph += prod * (u64)((1u << __popc(A)) + (1u << __popc(rest)));
        const uint32_t low = rest & (0u - rest);
        u64 inner = 0;
#pragma unroll
        for (uint32_t B = 1; B < FULL; ++B)
            if ((B & rest) == B && (B & low) && B != rest) inner += (u64)c[B] * (u64)c[rest ^ B];
        three += (u64)c[A] * inner;
    }
    // the reference's float operations (see k_general's homogeneous branch)
    const double N = (double)nhap;
    double p3n1 = 1.0;
#pragma unroll
    for (int i = 1; i < NR; ++i) p3n1 *= 3.0;
    const double p3n = 3.0 * p3n1, p2n1 = (double)(1u << (NR - 1));
    const u64 Ff = c[FULL];
    const double d_two = (double)two, d_sph = (double)two_sph, d_three = (double)three;
    res[0] = (double)Ff / N;
    res[1] = (double)Ff / (p2n1 * N) + d_two / p2n1 / (N * N);
    res[2] = (double)(Ff * ((1ull << NR) + 1ull)) / (p3n * N) + d_sph / p3n / (N * N);
    res[3] = (double)Ff / (p3n1 * N) + (2.0 * d_two) / p3n1 / (N * N) + (2.0 * d_three) / p3n1 / (N * N * N);
}

template <int NR, int W64>
__global__ void __launch_bounds__(MidPipe<NR, W64>::NWARPS * 32, 1) k_mid(const __grid_constant__ SmallArgs a)
{
    using P = MidPipe<NR, W64>;
    constexpr int NS = 1 << NR, LPD = 16, DPW = 2, ROWS = DPW * NR, NW32 = 2 * W64;
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane_in = lane % LPD, dgrp = lane / LPD;
    unsigned char *wbase = smem + (size_t)warp * P::STAGE_BYTES;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + P::BARS) + warp;
    uint4 *batch = reinterpret_cast<uint4 *>(smem + (size_t)P::NWARPS * P::STAGE_BYTES + (size_t)warp * P::BATCH_BYTES);
    constexpr int TPB = 32 / DPW;                                             // tiles per polynomial batch

    if (lane == 0) mbar_init(bar, 1);
    fence_proxy_async();
    __syncwarp();

    const int wp = a.pv.wp[0];
    const uint32_t row_bytes = (uint32_t)wp * 8u;
    const int64_t n_tiles = (a.n_draws + DPW - 1) / DPW;
    const int64_t tile0 = (int64_t)blockIdx.x * P::NWARPS + warp, tstride = (int64_t)gridDim.x * P::NWARPS;
    const int64_t my_tiles = tile0 < n_tiles ? (n_tiles - tile0 + tstride - 1) / tstride : 0;

    auto load_idx = [&](int64_t tile) -> int32_t {
        if (lane >= ROWS) return 0;
        int64_t d = tile * DPW + lane / NR;
        if (d >= a.n_draws) d = a.n_draws - 1;
        return __ldg(a.idx + d * NR + lane % NR);
    };
    auto issue = [&](int32_t ridx) -> uint32_t {
        uint32_t pc = 0;
        if (lane == 0) mbar_arrive_expect_tx(bar, row_bytes * ROWS);
        __syncwarp();
        if (lane < ROWS) {
            bulk_copy_g2s(wbase + (size_t)lane * row_bytes, a.pv.masks[0] + (int64_t)ridx * wp, row_bytes, bar);
            pc = __ldg(a.pv.popc[0] + ridx);
        }
        return pc;
    };

    uint32_t pc_cur = 0;
    int32_t idx_nxt = 0;
    if (my_tiles > 0) pc_cur = issue(load_idx(tile0));
    if (my_tiles > 1) idx_nxt = load_idx(tile0 + tstride);

    uint32_t parity = 0;
    for (int64_t q = 0; q < my_tiles; ++q) {
        const int64_t tile = tile0 + q * tstride;
        const int64_t draw = tile * DPW + dgrp;
        const bool writer = draw < a.n_draws && lane_in == 0;

        // packed counters: subset s lives in the (s & 1) half of c2[s >> 1]; per-lane partials stay below 2^16
        uint32_t c2[NS / 2];
#pragma unroll
        for (int s = 0; s < NS / 2; ++s) c2[s] = 0;
#pragma unroll
        for (int i = 0; i < NR; ++i) {
            const uint32_t v = __shfl_sync(0xffffffffu, pc_cur, dgrp * NR + i);
            if (lane_in == 0) c2[(1 << i) >> 1] += v << (16 * ((1 << i) & 1));
        }

        mbar_wait(bar, parity);
        parity ^= 1u;

        uint32_t m[NR][NW32];
        {
            const unsigned char *st = wbase + (size_t)(dgrp * NR) * row_bytes;
#pragma unroll
            for (int j = 0; j < W64; ++j) {
                const int w = j * LPD + lane_in;
                const bool in = w < wp;
#pragma unroll
                for (int i = 0; i < NR; ++i) {
                    const uint2 v = in ? *reinterpret_cast<const uint2 *>(st + ((size_t)i * wp + w) * 8) : make_uint2(0u, 0u);
                    m[i][2 * j] = v.x;
                    m[i][2 * j + 1] = v.y;
                }
            }
        }
        __syncwarp();
        if (q + 1 < my_tiles) {
            fence_proxy_async();
            pc_cur = issue(idx_nxt);
            if (q + 2 < my_tiles) idx_nxt = load_idx(tile + 2 * tstride);
        }

#pragma unroll
        for (int s = 3; s < NS; ++s) {
            if ((s & (s - 1)) == 0) continue;
            uint32_t t[NW32];
#pragma unroll
            for (int k = 0; k < NW32; ++k) {
                uint32_t x = ~0u;
#pragma unroll
                for (int i = 0; i < NR; ++i)
                    if ((s >> i) & 1) x &= m[i][k];
                t[k] = x;
            }
            c2[s >> 1] += PopcSum32<NW32, 1>::run(t) << (16 * (s & 1));
        }
        uint32_t c[NS];
#pragma unroll
        for (int s = 0; s < NS / 2; ++s) {
            uint32_t v = c2[s];
#pragma unroll
            for (int msk = LPD / 2; msk; msk >>= 1) v += __shfl_xor_sync(0xffffffffu, v, msk);
            c[2 * s] = v & 0xffffu;
            c[2 * s + 1] = v >> 16;
        }
        if (a.counts_out && writer) small_dump<NR>(a, draw, 0, c);
        // park this tile's counts: batch[quad][slot], slot = tile-in-batch * DPW + draw-in-tile
        const int k = (int)(q % TPB);
        if (lane_in == 0) {
            const int slot = k * DPW + dgrp;
#pragma unroll
            for (int j = 0; j < NS / 4; ++j) batch[j * 32 + slot] = make_uint4(c[4 * j], c[4 * j + 1], c[4 * j + 2], c[4 * j + 3]);
        }
        if (k == TPB - 1 || q + 1 == my_tiles) {
            __syncwarp();
            uint32_t cc[NS];
#pragma unroll
            for (int j = 0; j < NS / 4; ++j) {
                const uint4 v = batch[j * 32 + lane];
                cc[4 * j] = v.x; cc[4 * j + 1] = v.y; cc[4 * j + 2] = v.z; cc[4 * j + 3] = v.w;
            }
            const int kk = lane / DPW;
            const int64_t draw_k = (tile0 + (q - k + kk) * tstride) * DPW + lane % DPW;
            if (kk <= k) {
                double res[4];
                mid_polynomial<NR>(cc, a.pv.nhap[0], res);
                if (draw_k < a.n_draws) small_store<true>(a, draw_k, res);
            }
            __syncwarp();
        }
    }
}

}  // namespace ldpgta
