// k_lanes.cuh -- draws of 5..12 reads: lanes own the low subsets, registers hold the high-subset counters.
//
// The table kernel (k_general) spends more than half of its instructions outside the AND+POPC loop when a draw
// keeps only one or two warps busy (7..10 reads): building its L/H tables chunk by chunk, block barriers, staging
// stalls.  Here nothing is tabulated before the counts:
//
//   * 2^KL lanes own a draw (a warp: 32 >> KL draws); lane l owns the low subset lo = l of the first KL reads and
//     forms x_w = AND of its low reads' word w on the fly (keep-mask trick: read i enters as (mask | keep_i));
//   * for every one of the 2^KH subsets hi of the other KH = NR - KL reads the lane ANDs x_w with the high reads of
//     hi (compile-time subsets, LOP3 folds up to three inputs) and counts the result; three words at a time go
//     through a carry-save adder, so 3 word-ops cost 2 POPC.64;  F[hi << KL | lo] lives in a register counter;
//   * every mask word is read from shared memory exactly once per lane (a broadcast per draw); the masks arrive by
//     bulk copies (TMA 1-D), and the next tile's copy is issued as soon as the counts are done, so it runs under the
//     polynomial and under the other warps' arithmetic (single buffer: shared memory buys warps, not buffers);
//   * the counters go to a per-draw table in shared memory and the register-blocked partition polynomials of
//     poly_general.cu run on it, one lane group per draw, for the three model classes.
//
// Algorithmic work per draw: G*(2^NR-1)*W word AND+POPC (joint_frequencies_combo, HOMOGENOUES_MODELS.py:129-163),
// polynomials as in k_general (likelihoods(), :185-225 and twins).
#pragma once
#include "common.cuh"
#include "k_small.cuh"
#include "k_general.cuh"
#include "poly_finish.cuh"

namespace ldpgta {

struct LanesSmem {           // per-warp byte offsets / sizes
    int draw_stride;         // 64-bit words between consecutive draws of a buffer (bank-staggered)
    int buf_words;           // words per buffer
    int ftab;                // byte offset of the warp's count tables
    int etab;                // byte offset of the warp's e(S) tables (distant admixture), -1 when absent
    int per_warp;            // bytes per warp (mask buffer + tables + barrier)
};

template <int NR, int KL, typename CT>
__host__ __device__ inline LanesSmem lanes_smem_layout(const PanelView &pv, bool etab)       // etab: room for e(S) of a distant model with three or more groups
{
    constexpr int DPW = 32 >> KL;
    LanesSmem s;
    int words = 0;
    for (int g = 0; g < pv.G; ++g) words += NR * pv.wp[g];
    // consecutive draws start 16 / DPW word-pairs apart (mod 16 words = 128 B) so that the per-draw broadcasts of one
    // warp-wide LDS.64 fall into different banks
    const int stagger = DPW > 1 ? 16 / DPW : 0;
    int stride = (words + 15) & ~15;
    stride += stagger;
    s.draw_stride = stride;
    s.buf_words = stride * DPW;
    s.ftab = s.buf_words * 8;
    int off = s.ftab + DPW * ((pv.G << NR) * (int)sizeof(CT));
    off = (off + 63) & ~63;
    off = (off + 15) & ~15;
    s.etab = -1;
    if (etab) { s.etab = off; off += DPW * (8 << NR); }
    off += 16;                                            // the warp's mbarrier
    s.per_warp = (off + 127) & ~127;
    return s;
}

// sum over the lanes of a draw's group (xor-shuffles below LPD)
template <int LPD>
__device__ __forceinline__ u64 group_sum_u64(u64 v)
{
#pragma unroll
    for (int m = LPD / 2; m; m >>= 1) v += shfl_xor_u64(v, m);
    return v;
}
template <int LPD>
__device__ __forceinline__ double group_sum_f64(double v)
{
#pragma unroll
    for (int m = LPD / 2; m; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    return v;
}

// 12 warps per SM (168 registers); 12 reads keep 64 packed counter registers plus 7 x 3 high words live: 8 warps (255)
template <int NR>
struct LanesWarps { static constexpr int MAX = NR >= 12 ? 8 : 12; };

// CT: count-table entries, uint16_t when every group has at most 65535 haplotypes (the polynomial reads half the bytes)
template <int NR, int KL, typename CT>
__global__ void __launch_bounds__(32 * LanesWarps<NR>::MAX, 1) k_lanes(const __grid_constant__ GeneralArgs a)
{
    constexpr int KH = NR - KL, HS = 1 << KH, LPD = 1 << KL, DPW = 32 / LPD, NS = 1 << NR;
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), nwarps = blockDim.x >> 5;
    const int lo = lane & (LPD - 1), dgrp = lane / LPD;
    const int G = a.pv.G;
    const LanesSmem lay = lanes_smem_layout<NR, KL, CT>(a.pv, a.kind == LDPGTA_DISTANT_ADMIXTURE && a.pv.G > 2);
    unsigned char *wbase = smem + (size_t)warp * lay.per_warp;
    u64 *bufs = reinterpret_cast<u64 *>(wbase);
    CT *ftab = reinterpret_cast<CT *>(wbase + lay.ftab) + (size_t)dgrp * (G << NR);
    uint64_t *bar = reinterpret_cast<uint64_t *>(wbase + lay.per_warp - 16);

    if (lane == 0) mbar_init(bar, 1);
    fence_proxy_async();
    __syncwarp();

    const int64_t n_tiles = (a.n_draws + DPW - 1) / DPW;
    const int64_t tile0 = (int64_t)blockIdx.x * nwarps + warp, tstride = (int64_t)gridDim.x * nwarps;
    const int64_t my_tiles = tile0 < n_tiles ? (n_tiles - tile0 + tstride - 1) / tstride : 0;

    // one bulk copy per (draw, group, read) row of the tile
    uint32_t tile_bytes = 0;
    for (int g = 0; g < G; ++g) tile_bytes += (uint32_t)(DPW * NR * a.pv.wp[g] * 8);
    auto issue = [&](int64_t tile) {
        if (lane == 0) mbar_arrive_expect_tx(bar, tile_bytes);
        __syncwarp();
        int goff = 0;
        for (int g = 0; g < G; ++g) {
            const int wp = a.pv.wp[g];
            for (int r = lane; r < DPW * NR; r += 32) {
                const int d = r / NR, i = r - d * NR;
                int64_t draw = tile * DPW + d;
                if (draw >= a.n_draws) draw = a.n_draws - 1;
                const int32_t ridx = __ldg(a.idx + draw * NR + i);
                bulk_copy_g2s(bufs + (size_t)d * lay.draw_stride + goff + i * wp, a.pv.masks[g] + (int64_t)ridx * wp,
                              (uint32_t)wp * 8u, bar);
            }
            goff += NR * wp;
        }
    };

    u64 keep[KL > 0 ? KL : 1];
#pragma unroll
    for (int i = 0; i < KL; ++i) keep[i] = ((lo >> i) & 1) ? 0ull : ~0ull;

    const PolyScale scale = poly_scale(NR);
    if (my_tiles > 0) issue(tile0);
    uint32_t parity = 0u;

    for (int64_t q = 0; q < my_tiles; ++q) {
        const int64_t tile = tile0 + q * tstride;
        const int64_t draw = tile * DPW + dgrp;
        mbar_wait(bar, parity);
        parity ^= 1u;

        const u64 *M0 = bufs + (size_t)dgrp * lay.draw_stride;
        int goff = 0;
        for (int g = 0; g < G; ++g) {
            const int wp = a.pv.wp[g];
            const u64 *M = M0 + goff;
            // counters: one register per high subset, or two 16-bit counters per register when there are 64 or more
            // (a lane's partial count is at most 64 * wp < 2^16, checked by the launcher)
            constexpr bool PACK = KH >= 6;
            constexpr int NACC = PACK ? HS / 2 : HS;
            uint32_t acc[NACC];
#pragma unroll
            for (int h = 0; h < NACC; ++h) acc[h] = 0;
            // three words of every subset per step: x = the lane's low subset, y = x restricted to the high reads of h
            auto triple = [&](const u64 (&lw)[KL > 0 ? KL : 1][3], const u64 (&hm)[KH][3]) {
                u64 x[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    u64 v = ~0ull;
#pragma unroll
                    for (int i = 0; i < KL; ++i) v &= lw[i][k] | keep[i];
                    x[k] = v;
                }
#pragma unroll
                for (int h = 0; h < HS; ++h) {
                    u64 y[3];
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        u64 v = x[k];
#pragma unroll
                        for (int j = 0; j < KH; ++j)
                            if ((h >> j) & 1) v &= hm[j][k];
                        y[k] = v;
                    }
                    if (PACK && (h & 1)) csa_accumulate<16>(acc[h >> 1], y[0], y[1], y[2]);
                    else csa_accumulate<0>(acc[PACK ? h >> 1 : h], y[0], y[1], y[2]);
                }
            };
            const u64 *row = M;
            int w = 0;
            for (; w + 3 <= wp; w += 3, row += 3) {
                u64 lw[KL > 0 ? KL : 1][3], hm[KH][3];
#pragma unroll
                for (int k = 0; k < 3; ++k) {
#pragma unroll
                    for (int i = 0; i < KL; ++i) lw[i][k] = row[i * wp + k];
#pragma unroll
                    for (int j = 0; j < KH; ++j) hm[j][k] = row[(KL + j) * wp + k];
                }
                triple(lw, hm);
            }
            if (w < wp) {                                          // one or two words left: the missing ones count nothing
                u64 lw[KL > 0 ? KL : 1][3], hm[KH][3];
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const bool in = w + k < wp;
                    const int kk = in ? k : 0;
#pragma unroll
                    for (int i = 0; i < KL; ++i) lw[i][k] = in ? row[i * wp + kk] : 0ull;     // x = 0 (F[0] is reset below)
#pragma unroll
                    for (int j = 0; j < KH; ++j) hm[j][k] = in ? row[(KL + j) * wp + kk] : 0ull;
                }
                triple(lw, hm);
            }
            CT *F = ftab + ((size_t)g << NR);
#pragma unroll
            for (int h = 0; h < HS; ++h) F[(h << KL) | lo] = (CT)(PACK ? (acc[h >> 1] >> (16 * (h & 1))) & 0xffffu : acc[h]);
            goff += NR * wp;
        }
        __syncwarp();                                          // all lanes are done with the masks; tables complete
        if (q + 1 < my_tiles) {                                // refill: the copy lands while the polynomials run
            fence_proxy_async();
            issue(tile + tstride);
        }
        if (lo == 0)
            for (int g = 0; g < G; ++g) ftab[(size_t)g << NR] = 0;                 // the empty subset is not a joint frequency
        __syncwarp();

        const bool live = draw < a.n_draws;
        if (a.counts_out && live) {
            uint32_t *o = a.counts_out + ((draw * G) << NR);
            for (int e = lo; e < (G << NR); e += LPD) o[e] = ftab[e];
        }

        // ---- partition polynomials: the draw's lane group works on the draw's table ------------------------------
        double res[4];
        if (a.kind == LDPGTA_HOMOGENEOUS) {
            u64 two = 0, two_sph = 0;
            u128 three = 0;
            if (a.wide) poly_homog_blocked<CT, u64>(ftab, NR, a.terms, a.n_terms, lo, LPD, two, two_sph, three);
            else poly_homog_blocked<CT, uint32_t>(ftab, NR, a.terms, a.n_terms, lo, LPD, two, two_sph, three, a.runs != 0);
            u64 v[6];
            v[0] = two; v[1] = two_sph;
            split_u128(three, v + 2);
#pragma unroll
            for (int k = 0; k < 6; ++k) v[k] = group_sum_u64<LPD>(v[k]);
            if (lo == 0)
                finish_homogeneous(scale, (double)a.pv.nhap[0], ftab[NS - 1], (double)v[0], (double)v[1], u128_to_double(join_u128(v + 2)), res);
        } else if (a.kind == LDPGTA_RECENT_ADMIXTURE) {
            const CT *F = ftab, *Gt = ftab + NS;
            u64 ff = 0, gg = 0, fg = 0, fg_sph = 0;
            u128 ffg = 0, ggf = 0;
            if (a.wide) poly_recent_blocked<CT, u64>(F, Gt, NR, a.terms, a.n_terms, lo, LPD, ff, gg, fg, fg_sph, ffg, ggf);
            else poly_recent_blocked<CT, uint32_t>(F, Gt, NR, a.terms, a.n_terms, lo, LPD, ff, gg, fg, fg_sph, ffg, ggf, a.runs != 0);
            u64 v[12];
            v[0] = ff; v[1] = gg; v[2] = fg; v[3] = fg_sph;
            split_u128(ffg, v + 4);
            split_u128(ggf, v + 8);
#pragma unroll
            for (int k = 0; k < 12; ++k) v[k] = group_sum_u64<LPD>(v[k]);
            if (lo == 0)
                finish_recent(scale, (double)a.pv.nhap[0], (double)a.pv.nhap[1], F[NS - 1], Gt[NS - 1], (double)v[0], (double)v[1], (double)v[2],
                              (double)v[3], u128_to_double(join_u128(v + 4)), u128_to_double(join_u128(v + 8)), res);
        } else if (G == 2) {
            // distant admixture of two groups as exact integer sums (poly_general.cuh: Cubic2)
            const CT *F = ftab, *Gt = ftab + NS;
            Cubic2 c;
            if (a.wide) poly_distant2_blocked<CT, u64>(F, Gt, NR, a.terms, a.n_terms, lo, LPD, c);
            else poly_distant2_blocked<CT, uint32_t>(F, Gt, NR, a.terms, a.n_terms, lo, LPD, c, a.runs != 0);
            u64 v[22];
            v[0] = c.ff; v[1] = c.gg; v[2] = c.fg; v[3] = c.ff_sph; v[4] = c.gg_sph; v[5] = c.fg_sph;
            split_u128(c.s0, v + 6); split_u128(c.s1, v + 10); split_u128(c.s2, v + 14); split_u128(c.s3, v + 18);
#pragma unroll
            for (int k = 0; k < 22; ++k) v[k] = group_sum_u64<LPD>(v[k]);
            if (lo == 0) {
                const double c0 = a.pv.coef[0], c1 = a.pv.coef[1], N0 = (double)a.pv.nhap[0], N1 = (double)a.pv.nhap[1];
                const double Ef = 0.0 + c0 * (double)F[NS - 1] / N0 + c1 * (double)Gt[NS - 1] / N1;
                finish_distant2(scale, c0 / N0, c1 / N1, Ef, (double)v[0], (double)v[1], (double)v[2], (double)v[3], (double)v[4], (double)v[5],
                                u128_to_double(join_u128(v + 6)), u128_to_double(join_u128(v + 10)), u128_to_double(join_u128(v + 14)),
                                u128_to_double(join_u128(v + 18)), res);
            }
        } else {
            // three or more groups: e(S) = sum_g p_g * cnt_g(S) / N_g (DISTANT_ADMIXTURE_MODELS.py:136-139), tabulated once per draw
            double *etab = reinterpret_cast<double *>(wbase + lay.etab) + ((size_t)dgrp << NR);
            distant_fill_etab<CT>(ftab, a.pv, NR, etab, lo, LPD);
            __syncwarp();
            double two = 0, two_sph = 0, three = 0;
            poly_distant_etab(etab, NR, a.terms, a.n_terms, lo, LPD, two, two_sph, three);
            two = group_sum_f64<LPD>(two); two_sph = group_sum_f64<LPD>(two_sph); three = group_sum_f64<LPD>(three);
            if (lo == 0) finish_distant(scale, etab[NS - 1], two, two_sph, three, res);
        }
        if (lo == 0 && live) {
            const int64_t o = a.pos ? a.pos[draw] : draw;
            double2 *dst = reinterpret_cast<double2 *>(a.out + 4 * o);
            dst[0] = make_double2(res[0], res[1]);
            dst[1] = make_double2(res[2], res[3]);
            if (a.llr) store_llrs(a.llr, a.llr_stride, o, res);
        }
        __syncwarp();                                          // tables are rewritten by the next tile
    }
}

}  // namespace ldpgta
