// k_general.cuh -- draws of 5..16 reads: one CTA per (window, draw).
//
// Phase 1  stage the n sampled read masks of every group in shared memory.
// Phase 2  joint-frequency table F[S] = popcount(AND_{i in S} mask_i) for all 2^n subsets
//          (joint_frequencies_combo, HOMOGENOUES_MODELS.py:129-163), meet-in-the-middle:
//          the reads are split into a low set (kl reads) and a high set (kh = n-kl reads);
//          L[lo][w] and H[hi][w] hold the ANDs of each half for a chunk of wc words, and
//          F[hi<<kl | lo] = sum_w popc(H[hi][w] & L[lo][w]) -- exactly one AND + one POPC
//          per (subset, word), the algorithmic minimum, with no dependent chains.  A warp
//          register-tiles TH his x (32*TL) los; L is laid out [w][lo] so a warp's loads are
//          conflict-free, H[w][hi] is a broadcast.
// Phase 3  the partition polynomials of likelihoods() (HOMOGENOUES_MODELS.py:185-225,
//          RECENT_ADMIXTURE_MODELS.py:198-247, DISTANT_ADMIXTURE_MODELS.py:215-256) straight
//          from the table, without the reference's model tables: sums over unordered
//          2-partitions {A, full\A} and 3-partitions {A,B,C} (A owns the last read, B owns
//          the lowest read outside A).  Integer class sums are exact (64/128-bit).
//
// Algorithmic work per draw: G*(2^n-1)*W word AND+POPC; 2^(n-1)-1 two-block products and
// S(n,3) three-block products.
#pragma once
#include "common.cuh"
#include "poly_general.cuh"
#include "poly_finish.cuh"

namespace ldpgta {

struct GeneralArgs {
    PanelView pv;
    int kind;
    int n, kl, kh;           // reads per draw; low/high split
    int wc;                  // words per table chunk
    int etab;                // distant admixture: e(S) tabulated in shared memory (1) or in a global slice per CTA (2)
    double *etab_scratch;    // the global slices, etab_stride doubles apart
    size_t etab_stride;
    const int32_t *idx;      // [n_draws][n]
    int64_t n_draws;
    const int64_t *pos;      // optional scatter positions
    double *out;             // [*][4]
    uint32_t *counts_out;    // optional [n_draws][G][2^n]
    double *llr;             // optional [5][llr_stride]: the five LLRs of every draw
    int64_t llr_stride;
    const u64 *terms;        // colourings (a_h | b_h << 16 | c_h << 32) of the n-5 high reads with b_h < c_h (poly_general.cuh)
    int64_t n_terms;
    int wide;                // 64-bit convolution accumulators (a group with more than 16383 haplotypes)
    int runs;                // terms lists runs of colourings with a common a_h (poly_general.cuh), integer models only
    unsigned char *table_scratch;   // non-null: the joint-frequency tables live in global memory (L2), one slice per CTA
    size_t table_stride;
};

struct GeneralSmem {         // byte offsets into dynamic shared memory
    int masks, ltab, htab, ftab, etab, red, total;
};

template <typename CT>
__host__ __device__ inline GeneralSmem general_smem_layout(const PanelView &pv, int n, int kl, int kh, int wc, bool table_in_smem = true,
                                                           bool etab = false)
{
    GeneralSmem s;
    int off = 128;                                       // s_idx[32]
    s.masks = off;
    int words = 0;
    for (int g = 0; g < pv.G; ++g) words += n * pv.wp[g];
    off += words * 8;
    s.ltab = off; off += (wc << kl) * 8;
    s.htab = off; off += (wc << kh) * 8;
    off = (off + 63) & ~63;                              // table rows are read with 128-bit loads
    s.ftab = off; if (table_in_smem) off += (int)(sizeof(CT) * pv.G) << n;
    off = (off + 15) & ~15;
    s.etab = -1;
    if (etab) { s.etab = off; off += 8 << n; }           // e(S) of the distant model, one double per subset
    s.red = off; off += 16 * 22 * 8;                     // block sums: up to 22 values per warp (two-group distant model)
    s.total = off;
    return s;
}

// block-wide sum of NV 64-bit values; result valid in thread 0
template <int NV>
__device__ __forceinline__ void block_sum_u64(u64 *v, u64 *red)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = warp_sum_u64(v[k]);
    if (nw == 1) return;
    __syncthreads();
    if (lane == 0)
#pragma unroll
        for (int k = 0; k < NV; ++k) red[warp * NV + k] = v[k];
    __syncthreads();
    if (threadIdx.x == 0)
        for (int w = 1; w < nw; ++w)
#pragma unroll
            for (int k = 0; k < NV; ++k) v[k] += red[w * NV + k];
}

__device__ __forceinline__ void split_u128(u128 x, u64 *limb)
{
#pragma unroll
    for (int k = 0; k < 4; ++k) limb[k] = (u64)(uint32_t)(x >> (32 * k));
}
__device__ __forceinline__ u128 join_u128(const u64 *limb)
{
    u128 x = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) x += (u128)limb[k] << (32 * k);
    return x;
}

// Splits the free reads of `rest` (all but its lowest) between the 32 lanes: the highest
// (up to five) free bits are fixed by the lane id, the others are enumerated by every lane.
struct LaneSplit {
    uint32_t low, x_hi, hi_mask, lo_mask;
    bool active;
};
__device__ __forceinline__ LaneSplit lane_split(uint32_t rest, int lane)
{
    LaneSplit s;
    s.low = rest & (0u - rest);
    uint32_t tmp = rest ^ s.low;
    const int nf = __popc(tmp), nh = nf < 5 ? nf : 5;
    s.x_hi = 0; s.hi_mask = 0;
    for (int k = 0; k < nh; ++k) {
        const uint32_t bit = 1u << (31 - __clz(tmp));
        tmp ^= bit;
        s.hi_mask |= bit;
        if ((lane >> (nh - 1 - k)) & 1) s.x_hi |= bit;
    }
    s.lo_mask = tmp;
    s.active = lane < (1 << nh);
    return s;
}

// NW consecutive 64-bit words from shared memory (16-byte aligned when NW is even): 128-bit loads
template <int NW_>
__device__ __forceinline__ void load_words(const u64 *p, u64 *v, bool in)
{
    if (NW_ % 2 == 0) {
#pragma unroll
        for (int q = 0; q < NW_ / 2; ++q) {
            const ulonglong2 x = in ? reinterpret_cast<const ulonglong2 *>(p)[q] : make_ulonglong2(0ull, 0ull);
            v[2 * q] = x.x;
            v[2 * q + 1] = x.y;
        }
    } else {
#pragma unroll
        for (int q = 0; q < NW_; ++q) v[q] = in ? p[q] : 0ull;
    }
}

// Layout of one word's 2^kl low-subset ANDs.  With TL = 4 a lane owns the low subsets 4 l .. 4 l + 3 (kl = 7: 128 of them);
// stored in subset order the lane's 32 bytes sit 32 bytes apart from its neighbour's and a 128-bit load hits every bank
// twice (1.8e8 conflicts per n=16 launch in ncu, mio_throttle the top stall).  So the first and the second pair of a lane's
// subsets go to two 512-byte halves in which the lanes' 16-byte pieces are contiguous: two conflict-free LDS.128.
template <int TL>
__device__ __forceinline__ int ltab_pos(int id)
{
    return TL == 4 ? ((id >> 1) & 1) * 64 + (id >> 2) * 2 + (id & 1) : id;
}
template <int TL>
__device__ __forceinline__ void load_low(const u64 *word_base, int lane, u64 *v, bool in)
{
    if (TL == 4) {
        const ulonglong2 x = reinterpret_cast<const ulonglong2 *>(word_base)[lane], y = reinterpret_cast<const ulonglong2 *>(word_base + 64)[lane];
        v[0] = x.x; v[1] = x.y; v[2 % TL] = y.x; v[3 % TL] = y.y;
    } else {
        load_words<TL>(word_base + lane * TL, v, in);
    }
}

template <int TH, int TL, typename CT>
__global__ void __launch_bounds__(512, 1) k_general(const GeneralArgs a)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int n = a.n, kl = a.kl, kh = a.kh, G = a.pv.G;
    const int LS = 1 << kl, HS = 1 << kh, NS = 1 << n;
    const GeneralSmem lay = general_smem_layout<CT>(a.pv, n, kl, kh, a.wc, a.table_scratch == nullptr, a.etab == 1);
    int32_t *s_idx = reinterpret_cast<int32_t *>(smem);
    u64 *s_masks = reinterpret_cast<u64 *>(smem + lay.masks);
    u64 *ltab = reinterpret_cast<u64 *>(smem + lay.ltab);
    u64 *htab = reinterpret_cast<u64 *>(smem + lay.htab);
    CT *ftab = a.table_scratch ? reinterpret_cast<CT *>(a.table_scratch + (size_t)blockIdx.x * a.table_stride)
                               : reinterpret_cast<CT *>(smem + lay.ftab);
    u64 *red = reinterpret_cast<u64 *>(smem + lay.red);

    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5, NW = T >> 5;
    const uint32_t full = (uint32_t)NS - 1u, top = 1u << (n - 1);

    for (int64_t draw = blockIdx.x; draw < a.n_draws; draw += gridDim.x) {
        __syncthreads();                                     // previous draw done with shared memory
        if (tid < n) s_idx[tid] = __ldg(a.idx + draw * n + tid);
        __syncthreads();

        // ---- phase 1: stage the masks ---------------------------------------------------
        {
            int goff = 0;
            for (int g = 0; g < G; ++g) {
                const int wp = a.pv.wp[g], slots = wp >> 1;
                for (int e = tid; e < n * slots; e += T) {
                    const int i = e / slots, s = e - i * slots;
                    const uint4 v = ldg128(a.pv.masks[g] + (int64_t)s_idx[i] * wp + 2 * s);
                    *reinterpret_cast<uint4 *>(s_masks + goff + i * wp + 2 * s) = v;
                }
                goff += n * wp;
            }
        }
        __syncthreads();

        // ---- phase 2: joint-frequency tables ---------------------------------------------
        {
            int goff = 0;
            for (int g = 0; g < G; ++g) {
                const int wp = a.pv.wp[g];
                const u64 *M = s_masks + goff;
                CT *F = ftab + ((size_t)g << n);
                for (int w0 = 0; w0 < wp; w0 += a.wc) {
                    const int cw = min(a.wc, wp - w0);
                    if (w0 || g) __syncthreads();            // tables of the previous chunk (or previous group) consumed
                    // branch-free: read i is ANDed in through (mask | keep_i) with keep_i = all ones when i is not in the subset.
                    // When the block is at least as wide as a table, a thread always builds the same subset id, so its
                    // keep masks are loop invariants.
                    constexpr int KL = TL == 1 ? 5 : 7;          // compile-time low split for 5 or more reads
                    if (kl == KL && T >= (1 << KL)) {
                        const int id = tid & ((1 << KL) - 1);
                        u64 keep[KL];
#pragma unroll
                        for (int i = 0; i < KL; ++i) keep[i] = ((id >> i) & 1) ? 0ull : ~0ull;
                        for (int w = tid >> KL; w < cw; w += T >> KL) {
                            u64 v = ~0ull;
#pragma unroll
                            for (int i = 0; i < KL; ++i) v &= M[i * wp + w0 + w] | keep[i];
                            ltab[(w << KL) + ltab_pos<TL>(id)] = v;
                        }
                    } else if (kl == KL) {
                        for (int e = tid; e < (cw << KL); e += T) {
                            const int w = e >> KL, id = e & ((1 << KL) - 1);
                            u64 v = ~0ull;
#pragma unroll
                            for (int i = 0; i < KL; ++i) v &= M[i * wp + w0 + w] | (((id >> i) & 1) ? 0ull : ~0ull);
                            ltab[(w << KL) + ltab_pos<TL>(id)] = v;
                        }
                    } else {
                        for (int e = tid; e < (cw << kl); e += T) {
                            const int w = e >> kl, id = e & (LS - 1);
                            u64 v = ~0ull;
                            for (int i = 0; i < kl; ++i) v &= M[i * wp + w0 + w] | (((id >> i) & 1) ? 0ull : ~0ull);
                            ltab[e] = v;
                        }
                    }
                    if (kh >= 3) {
                        // a thread builds the eight high subsets that differ in the last three reads from their common part:
                        // kh + 4 ANDs for eight entries instead of 8 kh (the build was 8 % of the kernel at 17 reads); the eight
                        // stores of a warp are each 256 contiguous bytes
                        const int qs = kh - 3, Q = 1 << qs;
                        const u64 *Mh = M + (size_t)kl * wp + w0;
                        for (int e = tid; e < (cw << qs); e += T) {
                            const int w = e >> qs, r = e & (Q - 1);
                            u64 b = ~0ull;
                            for (int i = 0; i < qs; ++i) b &= Mh[i * wp + w] | (((r >> i) & 1) ? 0ull : ~0ull);
                            const u64 m0 = Mh[qs * wp + w], m1 = Mh[(qs + 1) * wp + w], m2 = Mh[(qs + 2) * wp + w];
                            u64 *o = htab + (w << kh) + r;
                            const u64 b0 = b & m0, b1 = b & m1;
                            o[0] = b; o[Q] = b0; o[2 * Q] = b1; o[3 * Q] = b0 & m1;
                            o[4 * Q] = b & m2; o[5 * Q] = b0 & m2; o[6 * Q] = b1 & m2; o[7 * Q] = b0 & m1 & m2;
                        }
                    } else
                    for (int e = tid; e < (cw << kh); e += T) {
                        const int w = e >> kh, id = e & (HS - 1);
                        u64 v = ~0ull;
                        for (int i = 0; i < kh; ++i) v &= M[(kl + i) * wp + w0 + w] | (((id >> i) & 1) ? 0ull : ~0ull);
                        htab[e] = v;
                    }
                    __syncthreads();
                    const int ntiles = HS / TH;
                    for (int tile = warp; tile < ntiles; tile += NW) {
                        const int hi0 = tile * TH;
                        uint32_t acc[TH][TL];
#pragma unroll
                        for (int h = 0; h < TH; ++h)
#pragma unroll
                            for (int j = 0; j < TL; ++j) acc[h][j] = 0;
                        // a table in global memory (16-bit entries): the counts of the previous chunks are requested before the
                        // word loop instead of after it, where their L2 latency was 8 % of the kernel at 18 reads (ncu)
                        constexpr bool PRE = TL == 4 && sizeof(CT) == 2;
                        uint2 before[PRE ? TH : 1];
                        const bool pre = PRE && w0 && a.table_scratch != nullptr;
                        if (pre)
#pragma unroll
                            for (int h = 0; h < (PRE ? TH : 1); ++h)
                                before[h] = *reinterpret_cast<const uint2 *>(F + (((size_t)(hi0 + h) << kl) | (size_t)(lane * 4)));
                        // three words at a time: a carry-save adder folds their intersections into a "ones" and a
                        // "twos" word, so three word-ops cost two POPC.64 instead of three (POPC is the scarce pipe)
                        // a lane owns TL consecutive low subsets (lo = lane * TL + j) and the tile TH consecutive high ones, so
                        // both operand sets are fetched with 128-bit shared-memory loads
                        int w = 0;
                        for (; w + 3 <= cw; w += 3) {
                            u64 l[3][TL];
#pragma unroll
                            for (int k = 0; k < 3; ++k) load_low<TL>(ltab + ((w + k) << kl), lane, l[k], lane * TL < LS);
#pragma unroll
                            for (int h2 = 0; h2 < TH; h2 += 2) {
                                u64 hv[3][2];
#pragma unroll
                                for (int k = 0; k < 3; ++k) load_words<(TH > 1 ? 2 : 1)>(htab + ((w + k) << kh) + hi0 + h2, hv[k], true);
#pragma unroll
                                for (int hh = 0; hh < (TH > 1 ? 2 : 1); ++hh)
#pragma unroll
                                    for (int j = 0; j < TL; ++j) {
                                        csa_accumulate<0>(acc[h2 + hh][j], hv[0][hh] & l[0][j], hv[1][hh] & l[1][j], hv[2][hh] & l[2][j]);
                                    }
                            }
                        }
                        for (; w < cw; ++w) {
                            u64 l[TL];
                            load_low<TL>(ltab + (w << kl), lane, l, lane * TL < LS);
#pragma unroll
                            for (int h = 0; h < TH; ++h) {
                                const u64 hv = htab[(w << kh) + hi0 + h];
#pragma unroll
                                for (int j = 0; j < TL; ++j) acc[h][j] += __popcll(hv & l[j]);
                            }
                        }
                        if (TL == 4) {
                            // a lane's four consecutive low subsets are one 8-byte (16-bit entries) or 16-byte (32-bit) group of
                            // a table row: one vector read-modify-write per high subset, lanes side by side = conflict-free
                            // (entry-wise stores put lanes l and l + 16 on one bank: 2.0e8 conflicts per n=16 launch in ncu)
#pragma unroll
                            for (int h = 0; h < TH; ++h) {
                                CT *row = F + (((size_t)(hi0 + h) << kl) | (size_t)(lane * 4));
                                if (sizeof(CT) == 2) {
                                    uint2 v = make_uint2(acc[h][0] | (acc[h][1] << 16), acc[h][2] | (acc[h][3] << 16));
                                    if (pre) { v.x += before[PRE ? h : 0].x; v.y += before[PRE ? h : 0].y; }
                                    else if (w0) { const uint2 o = *reinterpret_cast<const uint2 *>(row); v.x += o.x; v.y += o.y; }   // no carry: counts < 2^16
                                    *reinterpret_cast<uint2 *>(row) = v;
                                } else {
                                    uint4 v = make_uint4(acc[h][0], acc[h][1], acc[h][2], acc[h][3]);
                                    if (w0) { const uint4 o = *reinterpret_cast<const uint4 *>(row); v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
                                    *reinterpret_cast<uint4 *>(row) = v;
                                }
                            }
                        } else {
#pragma unroll
                        for (int h = 0; h < TH; ++h)
#pragma unroll
                            for (int j = 0; j < TL; ++j) {
                                const int fi = ((hi0 + h) << kl) | (lane * TL + j);
                                if (lane * TL + j >= LS) continue;          // fewer than 32 low subsets (n < 5)
                                if (w0 == 0) F[fi] = (CT)acc[h][j];
                                else F[fi] = (CT)(F[fi] + acc[h][j]);
                            }
                        }
                    }
                }
                goff += n * wp;
            }
        }
        __syncthreads();
        if (tid < G) ftab[(size_t)tid << n] = 0;             // the empty subset is not a joint frequency
        __syncthreads();

        if (a.counts_out) {
            uint32_t *o = a.counts_out + ((draw * G) << n);
            for (int e = tid; e < (G << n); e += T) o[e] = ftab[e];
        }

        // ---- phase 3: partition polynomials ----------------------------------------------
        const PolyScale scale = poly_scale(n);
        double res[4];

        if (a.kind == LDPGTA_HOMOGENEOUS) {
            const CT *F = ftab;
            u64 two = 0, two_sph = 0;
            u128 three = 0;
            if (n >= 5) {
                if (a.wide) poly_homog_blocked<CT, u64>(F, n, a.terms, a.n_terms, tid, T, two, two_sph, three);
                else poly_homog_blocked<CT, uint32_t>(F, n, a.terms, a.n_terms, tid, T, two, two_sph, three, a.runs != 0);
            } else
            for (uint32_t A = top + warp; A < full; A += NW) {
                const uint32_t rest = full ^ A;
                const u64 Fa = F[A];
                if (lane == 0) {
                    const u64 prod = Fa * (u64)F[rest];
                    two += prod;
                    two_sph += prod * ((1ull << __popc(A)) + (1ull << __popc(rest)));
                }
                const LaneSplit s = lane_split(rest, lane);
                if (s.active && Fa) {
                    uint32_t x_lo = s.lo_mask;
                    bool go = true;
                    if (s.x_hi == s.hi_mask) {               // x == all free reads would leave C empty
                        go = s.lo_mask != 0;
                        x_lo = (s.lo_mask - 1) & s.lo_mask;
                    }
                    u64 inner = 0;
                    while (go) {
                        const uint32_t B = s.low | s.x_hi | x_lo;
                        inner += (u64)F[B] * (u64)F[rest ^ B];
                        go = x_lo != 0;
                        x_lo = (x_lo - 1) & s.lo_mask;
                    }
                    three += (u128)Fa * inner;
                }
            }
            u64 v[6];
            v[0] = two; v[1] = two_sph;
            split_u128(three, v + 2);
            block_sum_u64<6>(v, red);
            if (tid == 0) {
                finish_homogeneous(scale, (double)a.pv.nhap[0], F[full], (double)v[0], (double)v[1], u128_to_double(join_u128(v + 2)), res);
            }
        } else if (a.kind == LDPGTA_RECENT_ADMIXTURE) {
            const CT *F = ftab, *Gt = ftab + NS;
            u64 ff = 0, gg = 0, fg = 0, fg_sph = 0;
            u128 ffg = 0, ggf = 0;
            if (n >= 5) {
                if (a.wide) poly_recent_blocked<CT, u64>(F, Gt, n, a.terms, a.n_terms, tid, T, ff, gg, fg, fg_sph, ffg, ggf);
                else poly_recent_blocked<CT, uint32_t>(F, Gt, n, a.terms, a.n_terms, tid, T, ff, gg, fg, fg_sph, ffg, ggf, a.runs != 0);
            } else
            for (uint32_t A = top + warp; A < full; A += NW) {
                const uint32_t rest = full ^ A;
                const u64 Fa = F[A], Ga = Gt[A];
                if (lane == 0) {
                    const u64 cross = Fa * (u64)Gt[rest] + Ga * (u64)F[rest];
                    ff += Fa * (u64)F[rest];
                    gg += Ga * (u64)Gt[rest];
                    fg += cross;
                    fg_sph += cross * ((1ull << __popc(A)) + (1ull << __popc(rest)));
                }
                const LaneSplit s = lane_split(rest, lane);
                if (s.active && (Fa | Ga)) {
                    uint32_t x_lo = s.lo_mask;
                    bool go = true;
                    if (s.x_hi == s.hi_mask) {
                        go = s.lo_mask != 0;
                        x_lo = (s.lo_mask - 1) & s.lo_mask;
                    }
                    u64 in_ff = 0, in_gg = 0, in_fg = 0;
                    while (go) {
                        const uint32_t B = s.low | s.x_hi | x_lo, C = rest ^ B;
                        const u64 fb = F[B], fc = F[C], gb = Gt[B], gc = Gt[C];
                        in_ff += fb * fc;
                        in_gg += gb * gc;
                        in_fg += fb * gc + gb * fc;
                        go = x_lo != 0;
                        x_lo = (x_lo - 1) & s.lo_mask;
                    }
                    ffg += (u128)Ga * in_ff + (u128)Fa * in_fg;
                    ggf += (u128)Fa * in_gg + (u128)Ga * in_fg;
                }
            }
            u64 v[12];
            v[0] = ff; v[1] = gg; v[2] = fg; v[3] = fg_sph;
            split_u128(ffg, v + 4);
            split_u128(ggf, v + 8);
            block_sum_u64<12>(v, red);
            if (tid == 0) {
                finish_recent(scale, (double)a.pv.nhap[0], (double)a.pv.nhap[1], F[full], Gt[full], (double)v[0], (double)v[1], (double)v[2],
                              (double)v[3], u128_to_double(join_u128(v + 4)), u128_to_double(join_u128(v + 8)), res);
            }
        } else if (G == 2 && n >= 5) {
            // distant admixture of two groups: e(S) is linear in the two count tables, so the sums over partitions are exact
            // integer sums (poly_general.cuh: Cubic2), scaled by powers of p_g / N_g at the end
            const CT *F = ftab, *Gt = ftab + NS;
            Cubic2 c;
            if (a.wide) poly_distant2_blocked<CT, u64>(F, Gt, n, a.terms, a.n_terms, tid, T, c);
            else poly_distant2_blocked<CT, uint32_t>(F, Gt, n, a.terms, a.n_terms, tid, T, c, a.runs != 0);
            u64 v[22];
            v[0] = c.ff; v[1] = c.gg; v[2] = c.fg; v[3] = c.ff_sph; v[4] = c.gg_sph; v[5] = c.fg_sph;
            split_u128(c.s0, v + 6); split_u128(c.s1, v + 10); split_u128(c.s2, v + 14); split_u128(c.s3, v + 18);
            block_sum_u64<22>(v, red);
            if (tid == 0) {
                const double c0 = a.pv.coef[0], c1 = a.pv.coef[1], N0 = (double)a.pv.nhap[0], N1 = (double)a.pv.nhap[1];
                const double Ef = 0.0 + c0 * (double)F[full] / N0 + c1 * (double)Gt[full] / N1;      // as DISTANT_ADMIXTURE_MODELS.py:136-139
                finish_distant2(scale, c0 / N0, c1 / N1, Ef, (double)v[0], (double)v[1], (double)v[2], (double)v[3], (double)v[4], (double)v[5],
                                u128_to_double(join_u128(v + 6)), u128_to_double(join_u128(v + 10)), u128_to_double(join_u128(v + 14)),
                                u128_to_double(join_u128(v + 18)), res);
            }
        } else {
            // distant admixture of three or more groups (and forced 2..4 reads): e(S) = sum_g p_g * cnt_g(S) / N_g
            // (DISTANT_ADMIXTURE_MODELS.py:136-139)
            double cg[MAXG];
            for (int g = 0; g < G; ++g) cg[g] = a.pv.coef[g];
            auto E = [&](uint32_t S) {
                double e = 0.0;
                for (int g = 0; g < G; ++g) e = e + cg[g] * (double)ftab[((size_t)g << n) + S] / (double)a.pv.nhap[g];
                return e;
            };
            double two = 0, two_sph = 0, three = 0;
            double *etab = a.etab == 1 ? reinterpret_cast<double *>(smem + lay.etab)
                         : a.etab == 2 ? a.etab_scratch + (size_t)blockIdx.x * a.etab_stride : nullptr;
            if (n >= 5) {                                    // the launcher always provides a table for 5 or more reads
                // e(S) in shared memory is bank-swizzled (poly_general.cuh).  Not in a global slice: L2 has no banks to
                // conflict on and the eight addresses of a row cost 64-bit arithmetic there (17 reads: -17 %), nor below 12
                // reads, where the fill outweighs the polynomial.
                const bool swz = a.etab == 1 && n >= 12;
                distant_fill_etab<CT>(ftab, a.pv, n, etab, tid, T, swz);
                __syncthreads();
                poly_distant_etab(etab, n, a.terms, a.n_terms, tid, T, two, two_sph, three, swz);
            } else
            for (uint32_t A = top + warp; A < full; A += NW) {
                const uint32_t rest = full ^ A;
                const double Ea = E(A);
                if (lane == 0) {
                    const double prod = Ea * E(rest);
                    two += prod;
                    two_sph += prod * (double)((1ull << __popc(A)) + (1ull << __popc(rest)));
                }
                const LaneSplit s = lane_split(rest, lane);
                if (s.active && Ea != 0.0) {
                    uint32_t x_lo = s.lo_mask;
                    bool go = true;
                    if (s.x_hi == s.hi_mask) {
                        go = s.lo_mask != 0;
                        x_lo = (s.lo_mask - 1) & s.lo_mask;
                    }
                    double inner = 0;
                    while (go) {
                        const uint32_t B = s.low | s.x_hi | x_lo;
                        inner += E(B) * E(rest ^ B);
                        go = x_lo != 0;
                        x_lo = (x_lo - 1) & s.lo_mask;
                    }
                    three += Ea * inner;
                }
            }
            two = warp_sum_f64(two); two_sph = warp_sum_f64(two_sph); three = warp_sum_f64(three);
            double *redd = reinterpret_cast<double *>(red);
            if (NW > 1) {
                __syncthreads();
                if (lane == 0) { redd[warp * 3] = two; redd[warp * 3 + 1] = two_sph; redd[warp * 3 + 2] = three; }
                __syncthreads();
                if (tid == 0)
                    for (int w = 1; w < NW; ++w) { two += redd[w * 3]; two_sph += redd[w * 3 + 1]; three += redd[w * 3 + 2]; }
            }
            if (tid == 0) {
                finish_distant(scale, E(full), two, two_sph, three, res);
            }
        }
        if (tid == 0) {
            const int64_t o = a.pos ? a.pos[draw] : draw;
            double2 *dst = reinterpret_cast<double2 *>(a.out + 4 * o);
            dst[0] = make_double2(res[0], res[1]);
            dst[1] = make_double2(res[2], res[3]);
            if (a.llr) store_llrs(a.llr, a.llr_stride, o, res);
        }
    }
}

}  // namespace ldpgta
