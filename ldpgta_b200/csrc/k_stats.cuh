// k_stats.cuh -- mask construction and the reductions that follow the bootstrap loop.
#pragma once
#include "common.cuh"

namespace ldpgta {

// Pads caller rows [n_rows][words] to [n_rows][wp] and complements REF alleles inside the
// group's N bits (hap ^ (2^N - 1), HOMOGENOUES_MODELS.py:87,94).
static __global__ void k_pack_rows(const uint64_t *__restrict__ src, const uint8_t *__restrict__ complement,
                            uint64_t *__restrict__ dst, int64_t n_rows, int words, int wp, uint32_t nhap)
{
    const int64_t total = n_rows * wp;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / wp;
        const int w = (int)(e - r * wp);
        u64 v = 0;
        if (w < words) {
            v = src[r * words + w];
            const int bits = (int)nhap - 64 * w;                 // valid bits in this word
            const u64 valid = bits >= 64 ? ~0ull : ((1ull << bits) - 1ull);
            if (complement && complement[r]) v = ~v;
            v &= valid;
        }
        dst[e] = v;
    }
}

// allele row a = panel row legend_idx[a], complemented inside the group's N bits for REF alleles
// (build_hap_dict, HOMOGENOUES_MODELS.py:84-94, as a gather from the resident panel)
static __global__ void k_gather_rows(const uint64_t *__restrict__ panel, const int64_t *__restrict__ legend_idx,
                              const uint8_t *__restrict__ complement, uint64_t *__restrict__ dst, int64_t n_alleles, int wp,
                              uint32_t nhap)
{
    const int64_t total = n_alleles * wp;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t a = e / wp;
        const int w = (int)(e - a * wp);
        u64 v = panel[legend_idx[a] * wp + w];
        if (complement && complement[a]) {
            const int bits = (int)nhap - 64 * w;
            const u64 valid = bits >= 64 ? ~0ull : (bits > 0 ? ((1ull << bits) - 1ull) : 0ull);
            v = ~v & valid;
        }
        dst[e] = v;
    }
}

// mask[r] = AND of allele rows allele_idx[off[r] .. off[r+1])  (build_intrenal_hap_dict,
// HOMOGENOUES_MODELS.py:112-123, once per read instead of once per draw)
static __global__ void k_build_read_masks(const uint64_t *__restrict__ rows, const int32_t *__restrict__ allele_idx,
                                   const int64_t *__restrict__ off, int64_t n_reads, int wp, uint64_t *__restrict__ out)
{
    const int slots = wp >> 1;
    const int64_t total = n_reads * slots;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / slots;
        const int s = (int)(e - r * slots);
        uint4 v = make_uint4(~0u, ~0u, ~0u, ~0u);
        const int64_t b = off[r], eend = off[r + 1];
        if (b == eend) v = make_uint4(0, 0, 0, 0);
        for (int64_t k = b; k < eend; ++k) {
            const uint4 m = ldg128(rows + (int64_t)allele_idx[k] * wp + 2 * s);
            v.x &= m.x; v.y &= m.y; v.z &= m.z; v.w &= m.w;
        }
        *reinterpret_cast<uint4 *>(out + r * wp + 2 * s) = v;
    }
}

// One call of model.get_likelihoods(alleles) (HOMOGENOUES_MODELS.py:272-286): build_intrenal_hap_dict for the n
// haplotypes of the call, their popcounts and the identity draw 0..n-1 in ONE launch (one block per group); the inputs
// come straight from page-locked host memory the device can read (no staging copies for a few hundred bytes).
struct SingleArgs {
    const uint64_t *rows[MAXG];     // allele rows [n_alleles][wp[g]]
    uint64_t *masks[MAXG];          // out: [n][wp[g]] scratch read masks
    uint32_t *popc[MAXG];           // out: [n]
    int wp[MAXG];
    const int32_t *allele_idx;      // device-visible host memory
    const int64_t *off;             // [n + 1]
    int32_t *draw;                  // out: [n] = 0..n-1
    int n;
};

static __global__ void k_prepare_single(const SingleArgs a)
{
    const int g = blockIdx.x, wp = a.wp[g], slots = wp >> 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int e = threadIdx.x; e < a.n * slots; e += blockDim.x) {
        const int r = e / slots, s_ = e - r * slots;
        uint4 v = make_uint4(~0u, ~0u, ~0u, ~0u);
        const int64_t b = a.off[r], eend = a.off[r + 1];
        for (int64_t k = b; k < eend; ++k) {
            const uint4 m = ldg128(a.rows[g] + (int64_t)a.allele_idx[k] * wp + 2 * s_);
            v.x &= m.x; v.y &= m.y; v.z &= m.z; v.w &= m.w;
        }
        *reinterpret_cast<uint4 *>(a.masks[g] + (int64_t)r * wp + 2 * s_) = v;
    }
    if (g == 0 && threadIdx.x < a.n) a.draw[threadIdx.x] = threadIdx.x;
    __syncthreads();
    for (int r = warp; r < a.n; r += nw) {
        uint32_t c = 0;
        for (int w = lane; w < wp; w += 32) c += __popcll(a.masks[g][(int64_t)r * wp + w]);
#pragma unroll
        for (int m = 16; m; m >>= 1) c += __shfl_xor_sync(0xffffffffu, c, m);
        if (lane == 0) a.popc[g][r] = c;
    }
}

// out[i] = rows[list[i]] (and the rows' popcounts): the read masks in the order of the plan's window read lists, so that every
// window is a run of consecutive rows (k_small_win)
static __global__ void k_gather_listed_rows(const uint64_t *__restrict__ rows, const uint32_t *__restrict__ popc, const int32_t *__restrict__ list,
                                            int64_t n, int wp, uint64_t *__restrict__ out, uint32_t *__restrict__ out_popc)
{
    const int slots = wp >> 1;
    const int64_t total = n * slots;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = e / slots;
        const int s_ = (int)(e - i * slots);
        const int64_t r = list[i];
        *reinterpret_cast<uint4 *>(out + i * wp + 2 * s_) = ldg128(rows + r * wp + 2 * s_);
        if (s_ == 0) out_popc[i] = popc[r];
    }
}

// norm[x] = x / nhap with IEEE division: joint_frequencies_combo(normalize=True) divides every count
// by N (HOMOGENOUES_MODELS.py:160-161); counts are integers in 0..N, so a table is exact and turns
// 15 float64 divisions per draw into 15 cached loads.
static __global__ void k_norm_table(double *__restrict__ out, uint32_t nhap)
{
    for (uint32_t x = blockIdx.x * blockDim.x + threadIdx.x; x <= nhap; x += gridDim.x * blockDim.x)
        out[x] = (double)x / (double)nhap;
}

// popcount of every row (the single-read joint frequencies), one warp per row
static __global__ void k_row_popc(const uint64_t *__restrict__ rows, int64_t n_rows, int wp, uint32_t *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n_rows; r += warps) {
        uint32_t c = 0;
        for (int w = lane; w < wp; w += 32) c += __popcll(rows[r * wp + w]);
#pragma unroll
        for (int m = 16; m; m >>= 1) c += __shfl_xor_sync(0xffffffffu, c, m);
        if (lane == 0) out[r] = c;
    }
}

// Read scores (build_score_dict, ANEUPLOIDY_TEST.py:151-189), one warp per read.  A read's SNPs whose
// effective allele frequency sum_g alpha_g * popcount(hap_g) lies in [0.01, 0.99] qualify (:173-175); over
// the 2^k choices of (hap, hap ^ ones) for the k qualifying SNPs the score counts those whose haplotype
// frequency  sum_g w * popcount(AND of the chosen rows in group g)  lies in [min_HF, 1 - min_HF] (:181-184).
// `w` is what the reference actually applies to every group: the LAST group's alpha (its generators are
// consumed after the comprehension variable has moved on).  The stored allele rows are hap or its
// complement depending on the observed base; the set {hap, hap ^ ones} is the same either way.
struct ScoreArgs {
    const uint64_t *rows[MAXG];     // allele rows [n_alleles][wp[g]]
    int wp[MAXG];
    uint32_t nhap[MAXG];
    double alpha[MAXG];
    int G;
    const uint8_t *complement;      // 1 = the stored row is the complement of the panel row
    const int32_t *allele_idx;
    const int64_t *off;
    int64_t n_reads;
    double w, min_HF, max_HF;
    int32_t *out;                   // score per read; -k = the read has k > 12 qualifying SNPs (left to k_read_scores_long)
    const int64_t *long_items;      // k_read_scores_long: (read, chunk of 4096 combinations) pairs
    int64_t n_long_items;
};

constexpr int SCORE_MAXQ = 24;      // qualifying SNPs per read the device enumerates (2^24 combinations)
constexpr int SCORE_SHORT = 12;     // up to here one warp does the whole read

// qualifying SNPs of read r (:171-178): allele rows whose effective frequency lies in [0.01, 0.99]; returns their number
// (the list holds the first SCORE_MAXQ)
__device__ __forceinline__ int score_qualifying(const ScoreArgs &a, int64_t r, int lane, int32_t *q)
{
    const int64_t b = a.off[r], e = a.off[r + 1];
    int k = 0;
    for (int64_t t = b; t < e; ++t) {
        const int32_t ai = a.allele_idx[t];
        double freq = 0.0;
        for (int g = 0; g < a.G; ++g) {
            uint32_t c = 0;
            const uint64_t *row = a.rows[g] + (int64_t)ai * a.wp[g];
            for (int w = lane; w < a.wp[g]; w += 32) c += __popcll(row[w]);
#pragma unroll
            for (int m = 16; m; m >>= 1) c += __shfl_xor_sync(0xffffffffu, c, m);
            if (a.complement[ai]) c = a.nhap[g] - c;                       // popcount of the panel row itself
            freq = freq + a.alpha[g] * (double)c;
        }
        if (0.01 <= freq && freq <= 0.99) {
            if (k < SCORE_MAXQ) q[k] = ai;
            ++k;
        }
    }
    return k;
}

// number of combinations `choice` in [c0, c1) of (row, complemented row) per qualifying SNP whose haplotype frequency
// lies in [min_HF, max_HF] (:181-184)
__device__ __forceinline__ int32_t score_range(const ScoreArgs &a, const int32_t *q, int k, uint32_t c0, uint32_t c1, int lane)
{
    int32_t score = 0;
    for (uint32_t choice = c0; choice < c1; ++choice) {
        double freq = 0.0;
        for (int g = 0; g < a.G; ++g) {
            const int wp = a.wp[g];
            uint32_t c = 0;
            for (int w = lane; w < wp; w += 32) {
                const int bits = (int)a.nhap[g] - 64 * w;
                const u64 valid = bits >= 64 ? ~0ull : (bits > 0 ? ((1ull << bits) - 1ull) : 0ull);
                u64 v = valid;
                for (int i = 0; i < k; ++i) {
                    const u64 x = a.rows[g][(int64_t)q[i] * wp + w];
                    v &= ((choice >> i) & 1) ? ~x : x;
                }
                c += __popcll(v);
            }
#pragma unroll
            for (int m = 16; m; m >>= 1) c += __shfl_xor_sync(0xffffffffu, c, m);
            freq = freq + a.w * (double)c;
        }
        score += (a.min_HF <= freq && freq <= a.max_HF) ? 1 : 0;
    }
    return score;
}

static __global__ void k_read_scores(const ScoreArgs a)
{
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < a.n_reads; r += warps) {
        int32_t q[SCORE_MAXQ];
        const int k = score_qualifying(a, r, lane, q);
        int32_t score = 0;
        if (k > SCORE_SHORT) score = -k;                      // long read: the host schedules k_read_scores_long for it
        else if (k > 0) score = score_range(a, q, k, 0u, 1u << k, lane);
        if (lane == 0) a.out[r] = score;
    }
}

// reads with 13..24 qualifying SNPs: one warp per (read, block of 4096 combinations), counts added atomically (integers)
static __global__ void k_read_scores_long(const ScoreArgs a)
{
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t it = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; it < a.n_long_items; it += warps) {
        const int64_t r = a.long_items[2 * it], chunk = a.long_items[2 * it + 1];
        int32_t q[SCORE_MAXQ];
        const int k = score_qualifying(a, r, lane, q);
        const uint32_t c0 = (uint32_t)chunk << 12;
        const int32_t part = score_range(a, q, k, c0, c0 + 4096u, lane);
        if (lane == 0 && part) atomicAdd(a.out + r, part);
    }
}

// pairs of statistics() (ANEUPLOIDY_TEST.py:298): indices into (monosomy, disomy, SPH, BPH)
static __device__ __constant__ int c_pair_num[5] = {3, 1, 3, 1, 2};
static __device__ __constant__ int c_pair_den[5] = {2, 0, 1, 2, 0};

// the five LLRs of one draw from v = (monosomy, disomy, SPH, BPH), pairs as above with compile-time indices (indexing v[]
// through the constant tables would put it in local memory)
__device__ __forceinline__ void five_llrs(const double (&v)[4], double (&x)[5])
{
    x[0] = llr_of(v[3], v[2]);
    x[1] = llr_of(v[1], v[0]);
    x[2] = llr_of(v[3], v[1]);
    x[3] = llr_of(v[1], v[2]);
    x[4] = llr_of(v[2], v[0]);
}

// error-free accumulation (Neumaier) so that window means match the reference's
// exact-rational statistics.mean/variance (ANEUPLOIDY_TEST.py:51-56) to the last bits
struct Comp {
    double s, c;
    __device__ __forceinline__ void add(double x)
    {
        const double t = s + x;
        c += (fabs(s) >= fabs(x)) ? ((s - t) + x) : ((x - t) + s);
        s = t;
    }
    __device__ __forceinline__ double value() const { return s + c; }
};

__device__ __forceinline__ double warp_sum_comp(Comp a)
{
#pragma unroll
    for (int m = 16; m; m >>= 1) {
        const double os = __shfl_xor_sync(0xffffffffu, a.s, m), oc = __shfl_xor_sync(0xffffffffu, a.c, m);
        a.add(os);
        a.c += oc;
    }
    return a.value();
}

// ---- statistics() on the device: likelihoods -> LLRs -> window mean / variance -> segment partial sums ---------------
// A segment is one (embryo, chromosome): a run of consecutive windows whose LLRs statistics() (ANEUPLOIDY_TEST.py:289-325)
// reduces together.
//   k_draw_llr      one thread per draw: the five LLRs (:78-90, :298), [5][n_draws].  float64 log and division: this is
//                   the arithmetic floor of the reductions (5 logs per draw).
//                   (the draw kernels of the resident pipeline write the LLRs themselves: store_llrs, common.cuh)
//   k_window_stats  one warp per window: mean and sample variance of the five LLR series (:51-56, :309)
//   k_segment_partials  one block per (segment, pair): totals and the column sums of zip(*A) (:73) over the segment's
//                   windows, fixed summation order (no atomics), optionally finished into mean / std (:65-76)
static __global__ void k_draw_llr(const double *__restrict__ lik, int64_t n_draws, double *__restrict__ out)
{
    for (int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; d < n_draws; d += (int64_t)gridDim.x * blockDim.x) {
        const double2 a = __ldg(reinterpret_cast<const double2 *>(lik) + 2 * d), b = __ldg(reinterpret_cast<const double2 *>(lik) + 2 * d + 1);
        const double v[4] = {a.x, a.y, b.x, b.y};
        double x[5];
        five_llrs(v, x);
#pragma unroll
        for (int p = 0; p < 5; ++p) out[p * n_draws + d] = x[p];
    }
}

struct StatsArgs {
    double *llr;               // [5][n_draws] per-draw LLRs: read, or written here when lik is given
    const double *lik;         // optional [n_draws][4]: the LLRs are formed here (one pass less than k_draw_llr + reading them back)
    const int64_t *doff;       // [n_windows + 1] draws of window w = [doff[w], doff[w+1])
    int64_t n_windows, n_draws;
    double *mean_var;          // [5][n_windows][2]
    double *wsum;              // [5][n_windows] sum of the window's LLRs (feeds the chromosome totals)
};

// A window of at most 32 draws on one warp, one draw per lane: the five LLRs (formed from the likelihoods, or read), their
// plain tree sums (5 roundings), mean and sample variance; writes mean_var / wsum and leaves the lane's LLRs in x[] (0 for
// lanes without a draw), the LLR sum of pair `lane` in mine and its mean in mean_l (lanes 0..4).
__device__ __forceinline__ void window_le32(const StatsArgs &a, int64_t w, int64_t b, int64_t e, int lane, bool write_llr,
                                            double (&x)[5], double &mine, double &mean_l)
{
    const double cnt = (double)(e - b);
    const bool live = b + lane < e;
    double sum[5], sq[5];
    if (a.lik) {
        double v[4] = {0.0, 0.0, 0.0, 0.0};
        if (live) {
            const double2 u0 = __ldg(reinterpret_cast<const double2 *>(a.lik) + 2 * (b + lane)),
                          u1 = __ldg(reinterpret_cast<const double2 *>(a.lik) + 2 * (b + lane) + 1);
            v[0] = u0.x; v[1] = u0.y; v[2] = u1.x; v[3] = u1.y;
        }
        five_llrs(v, x);                                   // lanes without a draw: LLR(0, 0) = 0
#pragma unroll
        for (int p = 0; p < 5; ++p)
            if (live && write_llr) a.llr[(int64_t)p * a.n_draws + b + lane] = x[p];
    } else {
#pragma unroll
        for (int p = 0; p < 5; ++p) x[p] = live ? a.llr[(int64_t)p * a.n_draws + b + lane] : 0.0;
    }
#pragma unroll
    for (int p = 0; p < 5; ++p) sum[p] = warp_sum_f64(x[p]);
    // the five means are ONE float64 division executed by lanes 0..4 side by side (a division is ~40 instructions
    // whether one lane or all need it); they travel back by shuffle
    mine = sum[0];
#pragma unroll
    for (int p = 1; p < 5; ++p) mine = lane == p ? sum[p] : mine;
    mean_l = mine / cnt;
#pragma unroll
    for (int p = 0; p < 5; ++p) {
        const double m = __shfl_sync(0xffffffffu, mean_l, p);
        const double d = live ? x[p] - m : 0.0;
        sq[p] = warp_sum_f64(d * d);
    }
    double ss = sq[0];
#pragma unroll
    for (int p = 1; p < 5; ++p) ss = lane == p ? sq[p] : ss;
    if (lane < 5) {
        double *mv = a.mean_var + 2 * ((int64_t)lane * a.n_windows + w);
        mv[0] = mean_l;
        mv[1] = (e - b > 1) ? ss / (cnt - 1.0) : nan("");
        a.wsum[(int64_t)lane * a.n_windows + w] = mine;
    }
}

// One warp per window (and all five LLR pairs): mean = sum / count, variance from the squared deviations of a second
// pass -- registers for windows of at most 32 draws, L1/L2 otherwise; sums of more than 32 values are Neumaier-compensated
// (statistics.mean / statistics.variance of ANEUPLOIDY_TEST.py:51-56 are exact-rational).
static __global__ void k_window_stats(const StatsArgs a)
{
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < a.n_windows; w += warps) {
        const int64_t b = a.doff[w], e = a.doff[w + 1];
        const double cnt = (double)(e - b);
        if (e - b <= 32) {
            double x[5], mine, mean_l;
            window_le32(a, w, b, e, lane, true, x, mine, mean_l);
        } else {
#pragma unroll 1
            for (int p = 0; p < 5; ++p) {
                double *x = a.llr + (int64_t)p * a.n_draws;
                Comp acc = {0.0, 0.0};
                for (int64_t i = b + lane; i < e; i += 32) {
                    double v;
                    if (a.lik) {
                        const double y = a.lik[4 * i + c_pair_num[p]], z = a.lik[4 * i + c_pair_den[p]];
                        v = llr_of(y, z);
                        x[i] = v;                                   // read back below by the lane that wrote it
                    } else v = x[i];
                    acc.add(v);
                }
                const double s = warp_sum_comp(acc);
                const double m = s / cnt;
                Comp q = {0.0, 0.0};
                for (int64_t i = b + lane; i < e; i += 32) { const double d = x[i] - m; q.add(d * d); }
                const double ss = warp_sum_comp(q);
                if (lane == 0) {
                    double *mv = a.mean_var + 2 * ((int64_t)p * a.n_windows + w);
                    mv[0] = m;
                    mv[1] = ss / (cnt - 1.0);
                    a.wsum[(int64_t)p * a.n_windows + w] = s;
                }
            }
        }
    }
}

// mean_and_std_of_mean_of_rnd_var (ANEUPLOIDY_TEST.py:65-76) + fraction_of_negative_LLRs (:313) of one (segment, pair) by
// one warp: q = its partials.  The operations and their order are those of ldpgta_finish_chromosome (the squared
// deviations are added column by column); lanes only fetch the columns.
__device__ __forceinline__ void segment_finish(const double *q, int ncols, int first_draws, double *out3, int lane)
{
    // q may have been written by other blocks of this launch (k_segment_from_chunks): read it from L2
    const double M = __ldcg(q + 1), N = (double)first_draws;
    const double mu = __ldcg(q) / N;
    double ss = 0.0;
    for (int i0 = 0; i0 < ncols; i0 += 32) {
        const double mine = i0 + lane < ncols ? __ldcg(q + 3 + i0 + lane) : 0.0;
        const int cnt = min(32, ncols - i0);
        for (int j = 0; j < cnt; ++j) {
            const double d = __shfl_sync(0xffffffffu, mine, j) - mu;
            ss += d * d;
        }
    }
    if (lane == 0) {
        out3[0] = mu / M;
        out3[1] = sqrt(ss / (N - 1.0)) / M;
        out3[2] = __ldcg(q + 2) / M;
    }
}

// partials[s][p][0..ncp+3) of segment s = windows [seg_off[s], seg_off[s+1]) and LLR pair p, one block per (s, p):
//   [0] sum over windows of the window sums, [1] number of windows, [2] number of windows with a negative mean,
//   [3 + i] sum over windows of the LLR of draw i, for the columns i every window of the chromosome has (zip truncation,
//   ANEUPLOIDY_TEST.py:73; the other columns are zero).
// SEG_ROWS x SEG_COLS threads: row r adds the windows r, r + SEG_ROWS, ... of a column (four loads in flight), the rows are
// then combined in row order; every sum is Neumaier-compensated and its order is fixed by the launch shape, so results are
// bit-reproducible.  With stats != nullptr the block also finishes the (segment, pair): see segment_finish.
#ifndef LDPGTA_SEG_SPLIT
#define LDPGTA_SEG_SPLIT 1     // blocks per (segment, pair); measured on configs[1]: 1 -> 21 us, 4 -> 30 us (the launch is latency-bound)
#endif
constexpr int SEG_ROWS = 32, SEG_COLS = 32, SEG_SPLIT = LDPGTA_SEG_SPLIT;

struct SegmentArgs {
    const double *llr;         // [5][n_draws]
    const int64_t *doff;       // [n_windows + 1]
    const double *wsum;        // [5][n_windows]
    const double *mean_var;    // [5][n_windows][2]
    const int64_t *seg_off;
    const int32_t *seg_cols, *seg_first;
    int64_t n_windows, n_draws;
    int ncp;
    double *partials;          // [n_segments][5][ncp + 3]
    double *stats;             // optional [n_segments][5][3]
    double *split;             // scratch [n_segments][5][SEG_SPLIT][ncp + 3]: a segment's windows are dealt to SEG_SPLIT blocks
    unsigned int *arrived;     // [n_segments][5] zero on entry, zero again on exit
};

static __global__ void __launch_bounds__(SEG_ROWS * SEG_COLS) k_segment_partials(const SegmentArgs a)
{
    __shared__ double sh_s[SEG_ROWS][SEG_COLS + 1], sh_c[SEG_ROWS][SEG_COLS + 1];
    __shared__ bool sh_last;
    const int s = blockIdx.x, p = blockIdx.y, z = blockIdx.z, stride = a.ncp + 3;
    const int64_t seg_lo = a.seg_off[s], seg_hi = a.seg_off[s + 1];
    // block z of SEG_SPLIT takes a contiguous share of the segment's windows
    const int64_t share = (seg_hi - seg_lo + SEG_SPLIT - 1) / SEG_SPLIT;
    const int64_t lo = min(seg_lo + (int64_t)z * share, seg_hi), hi = min(lo + share, seg_hi);
    const int ncols = a.seg_cols[s];
    const int cx = threadIdx.x % SEG_COLS, row = threadIdx.x / SEG_COLS;
    double *mine = a.split + (((int64_t)s * 5 + p) * SEG_SPLIT + z) * stride;
    const double *x = a.llr + (int64_t)p * a.n_draws;
    // head: totals over the windows (thread t takes windows t, t + 1024, ...)
    {
        Comp tot = {0.0, 0.0};
        double neg = 0.0;
        for (int64_t w = lo + threadIdx.x; w < hi; w += SEG_ROWS * SEG_COLS) {
            tot.add(a.wsum[(int64_t)p * a.n_windows + w]);
            neg += a.mean_var[2 * ((int64_t)p * a.n_windows + w)] < 0.0 ? 1.0 : 0.0;
        }
#pragma unroll
        for (int m = 16; m; m >>= 1) {
            const double os = __shfl_xor_sync(0xffffffffu, tot.s, m), oc = __shfl_xor_sync(0xffffffffu, tot.c, m);
            tot.add(os);
            tot.c += oc;
            neg += __shfl_xor_sync(0xffffffffu, neg, m);
        }
        if (cx == 0) { sh_s[row][0] = tot.s; sh_c[row][0] = tot.c; sh_s[row][1] = neg; }
        __syncthreads();
        if (threadIdx.x == 0) {
            Comp t = {0.0, 0.0};
            double n = 0.0;
            for (int r = 0; r < SEG_ROWS; ++r) { t.add(sh_s[r][0]); t.c += sh_c[r][0]; n += sh_s[r][1]; }
            mine[0] = t.value();
            mine[1] = (double)(hi - lo);
            mine[2] = n;
        }
        __syncthreads();
    }
    // columns
    for (int col0 = 0; col0 < a.ncp; col0 += SEG_COLS) {
        const int col = col0 + cx;
        Comp acc = {0.0, 0.0};
        if (col < ncols) {
            int64_t w = lo + row;
            for (; w + 7 * SEG_ROWS < hi; w += 8 * SEG_ROWS) {                       // eight loads in flight
                double v[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) v[k] = x[a.doff[w + k * SEG_ROWS] + col];
#pragma unroll
                for (int k = 0; k < 8; ++k) acc.add(v[k]);
            }
            for (; w < hi; w += SEG_ROWS) acc.add(x[a.doff[w] + col]);
        }
        sh_s[row][cx] = acc.s; sh_c[row][cx] = acc.c;
        __syncthreads();
        if (row == 0 && col < a.ncp) {
            Comp t = {0.0, 0.0};
#pragma unroll
            for (int r = 0; r < SEG_ROWS; ++r) { t.add(sh_s[r][cx]); t.c += sh_c[r][cx]; }
            mine[3 + col] = t.value();
        }
        __syncthreads();
    }
    // the last of the segment's SEG_SPLIT blocks adds the shares in share order (whichever block that is, the sum is the same)
    __threadfence();
    if (threadIdx.x == 0) {
        const unsigned int seen = atomicAdd(a.arrived + s * 5 + p, 1u);
        sh_last = seen == SEG_SPLIT - 1;
        if (sh_last) a.arrived[s * 5 + p] = 0u;                       // ready for the next launch
    }
    __syncthreads();
    if (!sh_last) return;
    __threadfence();
    double *out = a.partials + ((int64_t)s * 5 + p) * stride;
    const double *all = a.split + ((int64_t)s * 5 + p) * SEG_SPLIT * stride;
    for (int col = threadIdx.x; col < stride; col += blockDim.x) {
        Comp t = {0.0, 0.0};
#pragma unroll
        for (int k = 0; k < SEG_SPLIT; ++k) t.add(__ldcg(all + (int64_t)k * stride + col));
        out[col] = t.value();
    }
    __syncthreads();
    if (a.stats && threadIdx.x < 32) {
        __threadfence_block();
        segment_finish(out, ncols, a.seg_first[s], a.stats + 3 * ((int64_t)s * 5 + p), threadIdx.x);
    }
}

// the finish alone, for partials that were summed across GPUs first: one warp per (segment, pair).  out[s][p][3].
static __global__ void k_finish_segments(const double *__restrict__ partials, const int32_t *__restrict__ seg_cols,
                                         const int32_t *__restrict__ seg_first, int n_segments, int ncp, double *__restrict__ out)
{
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (t >= n_segments * 5) return;
    segment_finish(partials + (int64_t)t * (ncp + 3), seg_cols[t / 5], seg_first[t / 5], out + 3 * (int64_t)t, threadIdx.x & 31);
}

// ---- the same reductions with the column sums formed next to the window statistics --------------------------------------
// k_segment_partials walks the whole LLR matrix (40 bytes per draw) a second time for the column sums of zip(*A).  Here a
// block takes a *chunk* of up to CH_WIN consecutive windows of ONE segment (chunk table made by ldpgta_set_windows) and adds
// the chunk's LLRs column by column while they are in registers:
//   colpart[chunk][pair][0] = sum of the window sums, [1] = windows with a negative mean, [2 + i] = sum of column i (i < ncp).
// k_segment_from_chunks then adds (ncp + 2) * 8 bytes per (chunk, pair) instead of reading every LLR, and the LLR matrix need
// not be written at all when nobody asked for it.  Summation orders are fixed by the launch shape (bit-reproducible).
//   k_window_stats_chunks      no window has more than 32 draws: one window per warp (window_le32), CH_WIN warps
//   k_window_stats_block<DPT, NT>  windows of up to NT * DPT draws: the block takes its windows one after the other, thread t
//                              owns draws t, t + NT, ... of each (LLRs in registers for the second pass and the column sums)
constexpr int CH_WIN = 8;

struct ChunkStatsArgs {
    StatsArgs st;
    const int64_t *chunk_win;  // [n_chunks + 1] first window of each chunk (a chunk never straddles a segment)
    int64_t n_chunks;
    double *colpart;           // [n_chunks][5][ncp + 2]
    int ncp;
    int write_llr;
};

static __global__ void __launch_bounds__(CH_WIN * 32, 4) k_window_stats_chunks(const ChunkStatsArgs a)
{
    __shared__ double xs[CH_WIN][5][32];
    __shared__ double tot[CH_WIN][5][2];
    const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5, cstride = a.ncp + 2;
    for (int64_t ch = blockIdx.x; ch < a.n_chunks; ch += gridDim.x) {
        const int64_t w0 = a.chunk_win[ch];
        const int nw = (int)(a.chunk_win[ch + 1] - w0);
        if (wi < nw) {
            const int64_t w = w0 + wi, b = a.st.doff[w], e = a.st.doff[w + 1];
            double x[5], mine, mean_l;
            window_le32(a.st, w, b, e, lane, a.write_llr != 0, x, mine, mean_l);
#pragma unroll
            for (int p = 0; p < 5; ++p) xs[wi][p][lane] = x[p];
            if (lane < 5) { tot[wi][lane][0] = mine; tot[wi][lane][1] = mean_l < 0.0 ? 1.0 : 0.0; }
        }
        __syncthreads();
        for (int t = threadIdx.x; t < 5 * cstride; t += CH_WIN * 32) {
            const int p = t / cstride, j = t % cstride;
            double s = 0.0;
            if (j >= 2) for (int k = 0; k < nw; ++k) s += xs[k][p][j - 2];
            else        for (int k = 0; k < nw; ++k) s += tot[k][p][j];
            a.colpart[(ch * 5 + p) * cstride + j] = s;
        }
        __syncthreads();
    }
}

// the same without shared memory or block barriers: ONE WARP takes the windows of a chunk one after the other and keeps the
// column sums in registers (lane i = column i)
static __global__ void __launch_bounds__(256, 4) k_window_stats_warpchunks(const ChunkStatsArgs a)
{
    const int lane = threadIdx.x & 31, cstride = a.ncp + 2;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t ch = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; ch < a.n_chunks; ch += warps) {
        const int64_t w0 = a.chunk_win[ch];
        const int nw = (int)(a.chunk_win[ch + 1] - w0);
        double col[5] = {0.0, 0.0, 0.0, 0.0, 0.0}, tsum = 0.0, tneg = 0.0;
        for (int k = 0; k < nw; ++k) {
            const int64_t w = w0 + k, b = a.st.doff[w], e = a.st.doff[w + 1];
            double x[5], mine, mean_l;
            window_le32(a.st, w, b, e, lane, a.write_llr != 0, x, mine, mean_l);
#pragma unroll
            for (int p = 0; p < 5; ++p) col[p] += x[p];
            tsum += mine;                                    // lanes 0..4: pair `lane`
            tneg += mean_l < 0.0 ? 1.0 : 0.0;
        }
        double *out = a.colpart + ch * 5 * cstride;
#pragma unroll
        for (int p = 0; p < 5; ++p)
            if (lane < a.ncp) out[p * cstride + 2 + lane] = col[p];
        if (lane < 5) { out[lane * cstride] = tsum; out[lane * cstride + 1] = tneg; }
    }
}

template <int DPT, int NT, int MINB>
static __global__ void __launch_bounds__(NT, MINB) k_window_stats_block(const ChunkStatsArgs a)
{
    constexpr int NW = NT / 32;
    __shared__ double red[NW][5], red2[NW][5];
    const int t = threadIdx.x, lane = t & 31, wi = t >> 5, cstride = a.ncp + 2;
    for (int64_t ch = blockIdx.x; ch < a.n_chunks; ch += gridDim.x) {
        const int64_t w0 = a.chunk_win[ch];
        const int nw = (int)(a.chunk_win[ch + 1] - w0);
        double col[5][DPT];
#pragma unroll
        for (int p = 0; p < 5; ++p)
#pragma unroll
            for (int j = 0; j < DPT; ++j) col[p][j] = 0.0;
        Comp tsum = {0.0, 0.0};          // threads 0..4: sum of the window sums and negative means of pair t
        double tneg = 0.0;
        // the likelihoods of the next window are requested before this window's arithmetic and barriers
        double2 nx[DPT][2];
        int64_t nb = a.st.doff[w0], ne = a.st.doff[w0 + 1];
        auto fetch = [&](int64_t b_, int64_t e_) {
#pragma unroll
            for (int j = 0; j < DPT; ++j) {
                const int64_t i = b_ + t + NT * j;
                if (a.st.lik && i < e_) {
                    nx[j][0] = __ldg(reinterpret_cast<const double2 *>(a.st.lik) + 2 * i);
                    nx[j][1] = __ldg(reinterpret_cast<const double2 *>(a.st.lik) + 2 * i + 1);
                } else nx[j][0] = nx[j][1] = make_double2(0.0, 0.0);
            }
        };
        fetch(nb, ne);
        for (int k = 0; k < nw; ++k) {
            const int64_t w = w0 + k, b = nb, e = ne;
            const double cnt = (double)(e - b);
            double x[5][DPT];
            Comp acc[5];
            double2 cur[DPT][2];
#pragma unroll
            for (int j = 0; j < DPT; ++j) { cur[j][0] = nx[j][0]; cur[j][1] = nx[j][1]; }
            if (k + 1 < nw) { nb = e; ne = a.st.doff[w + 2]; fetch(nb, ne); }
#pragma unroll
            for (int p = 0; p < 5; ++p) acc[p] = {0.0, 0.0};
#pragma unroll
            for (int j = 0; j < DPT; ++j) {
                const int64_t i = b + t + NT * j;
                const bool live = i < e;
                if (a.st.lik) {
                    const double v[4] = {cur[j][0].x, cur[j][0].y, cur[j][1].x, cur[j][1].y};      // zeros without a draw: LLR(0, 0) = 0
                    double xj[5];
                    five_llrs(v, xj);
#pragma unroll
                    for (int p = 0; p < 5; ++p) {
                        x[p][j] = xj[p];
                        if (live && a.write_llr) a.st.llr[(int64_t)p * a.st.n_draws + i] = xj[p];
                    }
                } else {
#pragma unroll
                    for (int p = 0; p < 5; ++p) x[p][j] = live ? a.st.llr[(int64_t)p * a.st.n_draws + i] : 0.0;
                }
#pragma unroll
                for (int p = 0; p < 5; ++p) acc[p].add(x[p][j]);
            }
#pragma unroll
            for (int p = 0; p < 5; ++p) {
                const double v = warp_sum_comp(acc[p]);
                if (lane == 0) red[wi][p] = v;
            }
            __syncthreads();
            // every warp combines the warps' sums itself (lanes 0..4, pair `lane`): no second barrier before the deviations
            double sum_l = 0.0, mean_l = 0.0;
            if (lane < 5) {
                Comp c = {0.0, 0.0};
#pragma unroll
                for (int r = 0; r < NW; ++r) c.add(red[r][lane]);
                sum_l = c.value();
                mean_l = sum_l / cnt;
            }
#pragma unroll
            for (int p = 0; p < 5; ++p) {
                const double m = __shfl_sync(0xffffffffu, mean_l, p);
                Comp q = {0.0, 0.0};
#pragma unroll
                for (int j = 0; j < DPT; ++j) {
                    const double d = (b + t + NT * j < e) ? x[p][j] - m : 0.0;
                    q.add(d * d);
                    col[p][j] += x[p][j];
                }
                const double v = warp_sum_comp(q);
                if (lane == 0) red2[wi][p] = v;
            }
            __syncthreads();
            if (t < 5) {
                Comp c = {0.0, 0.0};
#pragma unroll
                for (int r = 0; r < NW; ++r) c.add(red2[r][t]);
                double *mv = a.st.mean_var + 2 * ((int64_t)t * a.st.n_windows + w);
                mv[0] = mean_l;
                mv[1] = (e - b > 1) ? c.value() / (cnt - 1.0) : nan("");
                a.st.wsum[(int64_t)t * a.st.n_windows + w] = sum_l;
                tsum.add(sum_l);
                tneg += mean_l < 0.0 ? 1.0 : 0.0;
            }
        }
        double *out = a.colpart + ch * 5 * cstride;
#pragma unroll
        for (int p = 0; p < 5; ++p)
#pragma unroll
            for (int j = 0; j < DPT; ++j)
                if (t + NT * j < a.ncp) out[p * cstride + 2 + t + NT * j] = col[p][j];
        if (t < 5) { out[t * cstride] = tsum.value(); out[t * cstride + 1] = tneg; }
    }
}

struct SegChunkArgs {
    const double *colpart;     // [n_chunks][5][ncp + 2]
    const int64_t *cseg_off;   // [n_segments + 1] chunks of segment s = [cseg_off[s], cseg_off[s+1])
    const int64_t *seg_off;
    const int32_t *seg_cols, *seg_first;
    int ncp;
    double *partials;          // [n_segments][5][ncp + 3]
    double *stats;             // optional [n_segments][5][3]
    unsigned int *arrived;     // [n_segments][5] zero on entry, zero again on exit
};
constexpr int SC_ROWS = 16, SC_COLS = 32;

// one block per (segment, pair, 32 entries of the chunk rows): SC_ROWS x SC_COLS threads, row r adds the chunks r,
// r + SC_ROWS, ... of its entry (Neumaier-compensated, four loads in flight), the rows are combined in row order; the last
// block of a (segment, pair) to finish also runs segment_finish.  Layout of partials as k_segment_partials.
static __global__ void __launch_bounds__(SC_ROWS * SC_COLS) k_segment_from_chunks(const SegChunkArgs a)
{
    __shared__ double sh_s[SC_ROWS][SC_COLS + 1], sh_c[SC_ROWS][SC_COLS + 1];
    __shared__ bool sh_last;
    const int s = blockIdx.x, p = blockIdx.y, stride = a.ncp + 3, cstride = a.ncp + 2;
    const int64_t lo = a.cseg_off[s], hi = a.cseg_off[s + 1];
    const int ncols = a.seg_cols[s];
    const int cx = threadIdx.x % SC_COLS, row = threadIdx.x / SC_COLS;
    const int j = blockIdx.z * SC_COLS + cx;                 // entry of the chunk rows
    const int64_t CS = 5 * (int64_t)cstride;
    Comp acc = {0.0, 0.0};
    if (j < cstride) {
        const double *x = a.colpart + (int64_t)p * cstride + j;
        int64_t c = lo + row;
        for (; c + 3 * SC_ROWS < hi; c += 4 * SC_ROWS) {
            double v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) v[k] = __ldg(x + (c + k * SC_ROWS) * CS);
#pragma unroll
            for (int k = 0; k < 4; ++k) acc.add(v[k]);
        }
        for (; c < hi; c += SC_ROWS) acc.add(__ldg(x + c * CS));
    }
    sh_s[row][cx] = acc.s; sh_c[row][cx] = acc.c;
    __syncthreads();
    double *out = a.partials + ((int64_t)s * 5 + p) * stride;
    if (row == 0 && j < cstride) {
        Comp t = {0.0, 0.0};
#pragma unroll
        for (int r = 0; r < SC_ROWS; ++r) { t.add(sh_s[r][cx]); t.c += sh_c[r][cx]; }
        const double v = t.value();
        if (j == 0) out[0] = v;
        else if (j == 1) { out[2] = v; out[1] = (double)(a.seg_off[s + 1] - a.seg_off[s]); }
        else out[1 + j] = (j - 2 < ncols) ? v : 0.0;
    }
    if (!a.stats) return;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int seen = atomicAdd(a.arrived + s * 5 + p, 1u);
        sh_last = seen == gridDim.z - 1;
        if (sh_last) a.arrived[s * 5 + p] = 0u;                       // ready for the next launch
    }
    __syncthreads();
    if (!sh_last) return;
    __threadfence();
    if (threadIdx.x < 32) segment_finish(out, ncols, a.seg_first[s], a.stats + 3 * ((int64_t)s * 5 + p), threadIdx.x);
}

// binning() of PLOT_PANEL.py:121-144 / BALANCED_ROC_CURVE.py:131-154: one block per (series, bin).  A bin is a run
// of windows [b0, b1); mean_and_std_of_mean_of_rnd_var (PLOT_PANEL.py:37-48) on the windows' LLR rows:
//   N = draws of the bin's FIRST window, M = b1 - b0, mu = (sum of all values) / N,
//   std = sqrt(sum_i (column_i - mu)^2 / (N - 1)) / M over the columns every window has (zip truncation), mean = mu / M.
// x[n_series][n_draws]; out[n_series][n_bins][2] = (mean, std).
static __global__ void k_bin_stats(const double *__restrict__ x, int64_t n_draws, const int64_t *__restrict__ woff,
                            const int64_t *__restrict__ ranges, int64_t n_bins, double *__restrict__ out)
{
    __shared__ double sh_s[32], sh_c[32];
    __shared__ int64_t sh_min[32];
    __shared__ double sh_mu;
    __shared__ int64_t sh_cols;
    const int64_t bin = blockIdx.x;
    const double *xs = x + (int64_t)blockIdx.y * n_draws;
    const int64_t b0 = ranges[2 * bin], b1 = ranges[2 * bin + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int64_t lo = woff[b0], hi = woff[b1];
    Comp acc = {0.0, 0.0};
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) acc.add(xs[i]);
    int64_t mn = INT64_MAX;
    for (int64_t w = b0 + threadIdx.x; w < b1; w += blockDim.x) mn = min(mn, woff[w + 1] - woff[w]);
#pragma unroll
    for (int m = 16; m; m >>= 1) {
        const double os = __shfl_xor_sync(0xffffffffu, acc.s, m), oc = __shfl_xor_sync(0xffffffffu, acc.c, m);
        acc.add(os);
        acc.c += oc;
        mn = min(mn, (int64_t)__shfl_xor_sync(0xffffffffu, (long long)mn, m));
    }
    if (lane == 0) { sh_s[warp] = acc.s; sh_c[warp] = acc.c; sh_min[warp] = mn; }
    __syncthreads();
    const double N = (double)(woff[b0 + 1] - woff[b0]), M = (double)(b1 - b0);
    if (threadIdx.x == 0) {
        Comp t = {0.0, 0.0};
        int64_t m = INT64_MAX;
        for (int w = 0; w < nwarps; ++w) { t.add(sh_s[w]); t.c += sh_c[w]; m = min(m, sh_min[w]); }
        sh_mu = t.value() / N;
        sh_cols = m;
    }
    __syncthreads();
    const double mu = sh_mu;
    const int64_t cols = sh_cols;
    Comp ss = {0.0, 0.0};
    for (int64_t i = threadIdx.x; i < cols; i += blockDim.x) {
        Comp col = {0.0, 0.0};
        for (int64_t w = b0; w < b1; ++w) col.add(xs[woff[w] + i]);
        const double d = col.value() - mu;
        ss.add(d * d);
    }
#pragma unroll
    for (int m = 16; m; m >>= 1) {
        const double os = __shfl_xor_sync(0xffffffffu, ss.s, m), oc = __shfl_xor_sync(0xffffffffu, ss.c, m);
        ss.add(os);
        ss.c += oc;
    }
    __syncthreads();
    if (lane == 0) { sh_s[warp] = ss.s; sh_c[warp] = ss.c; }
    __syncthreads();
    if (threadIdx.x == 0) {
        Comp t = {0.0, 0.0};
        for (int w = 0; w < nwarps; ++w) { t.add(sh_s[w]); t.c += sh_c[w]; }
        double *o = out + 2 * ((int64_t)blockIdx.y * n_bins + bin);
        o[0] = mu / M;
        o[1] = sqrt(t.value() / (N - 1.0)) / M;      // N = 1 divides by zero, as the reference does
    }
}

// validation of a batch's read indices on the device (the host path used to scan them before the first copy):
// an index outside [0, n_reads) is replaced by 0 and reported through *flag, the caller then fails the whole call.
static __global__ void k_sanitize_idx(int32_t *__restrict__ idx, int64_t count, uint32_t n_reads, unsigned int *__restrict__ flag)
{
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t quads = ((reinterpret_cast<uintptr_t>(idx) & 15) == 0) ? count / 4 : 0;
    bool bad = false;
    for (int64_t i = tid; i < quads; i += stride) {
        int4 v = reinterpret_cast<int4 *>(idx)[i];
        const bool b = (uint32_t)v.x >= n_reads || (uint32_t)v.y >= n_reads || (uint32_t)v.z >= n_reads || (uint32_t)v.w >= n_reads;
        if (b) {
            v.x = (uint32_t)v.x >= n_reads ? 0 : v.x;
            v.y = (uint32_t)v.y >= n_reads ? 0 : v.y;
            v.z = (uint32_t)v.z >= n_reads ? 0 : v.z;
            v.w = (uint32_t)v.w >= n_reads ? 0 : v.w;
            reinterpret_cast<int4 *>(idx)[i] = v;
            bad = true;
        }
    }
    for (int64_t i = quads * 4 + tid; i < count; i += stride)
        if ((uint32_t)idx[i] >= n_reads) { idx[i] = 0; bad = true; }
    if (bad) atomicOr(flag, 1u);
}

// ---- pick_reads (ANEUPLOIDY_TEST.py:226-233) on the device ------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011): a counter-based generator, so draw d of a call is a pure function of (seed, d).
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t *out)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// One thread per draw: K distinct reads of the draw's window, uniformly over K-subsets (Floyd's algorithm; for
// i = 0..K-1: t uniform in [0, R-K+i], taken unless already chosen, else R-K+i), written as read-mask rows.
// Draw d belongs to the window j with draw_off[j] <= d < draw_off[j+1] of the draw-size class being sampled; that is window
// w = cls_win[j] of the sample (identity when cls_win is null), whose reads are win_reads[win_off[w] ...] (or simply the
// rows win_off[w] ... when win_reads is null: read masks stored window by window).
// The j-th random integer of draw d is mulhi64(x, bound) with x = 64 bits of Philox(counter = (d, j/2), key = seed).
struct SampleArgs {
    const int64_t *draw_off;     // [n_windows + 1] draw offsets of the class
    const int32_t *win_of_draw;  // optional [n_draws]: class window of every draw (saves the search)
    const int64_t *win_off;      // read-list offsets of the sample's windows
    const int32_t *win_reads;    // or nullptr
    const int32_t *cls_win;      // or nullptr
    int64_t n_windows, n_draws;
    uint64_t seed;
    const uint64_t *seed_dev;    // overrides seed when set (a captured launch reads the seed of the replay)
    int32_t *idx_out;            // [n_draws][K] rows of the read-mask table (optional when pick_out is all the draw kernel reads)
    int32_t *pick_out;           // optional [n_draws][K]: positions inside the window's read list (k_small_win on read lists)
};

template <int K>
static __global__ void k_sample_draws(const SampleArgs a)
{
    const uint64_t seed = a.seed_dev ? *a.seed_dev : a.seed;
    for (int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; d < a.n_draws; d += (int64_t)gridDim.x * blockDim.x) {
        int64_t lo = 0, hi = a.n_windows;                    // upper_bound(draw_off, d) - 1
        if (a.win_of_draw) lo = a.win_of_draw[d];
        else
            while (hi - lo > 1) {
                const int64_t mid = (lo + hi) >> 1;
                if (a.draw_off[mid] <= d) lo = mid; else hi = mid;
            }
        const int64_t w = a.cls_win ? a.cls_win[lo] : lo;
        const int64_t base = a.win_off[w];
        const uint32_t R = (uint32_t)(a.win_off[w + 1] - base);
        uint32_t pick[K];
        uint32_t rnd[4];
#pragma unroll
        for (int i = 0; i < K; ++i) {
            if ((i & 1) == 0)
                philox4x32_10((uint32_t)d, (uint32_t)((uint64_t)d >> 32), (uint32_t)(i >> 1), 0u, (uint32_t)seed, (uint32_t)(seed >> 32), rnd);
            const uint64_t x = (i & 1) ? ((uint64_t)rnd[3] << 32 | rnd[2]) : ((uint64_t)rnd[1] << 32 | rnd[0]);
            const uint32_t top = R - K + i;
            uint32_t t = (uint32_t)__umul64hi(x, (uint64_t)top + 1ull);
            bool dup = false;
#pragma unroll
            for (int j = 0; j < i; ++j) dup |= pick[j] == t;
            pick[i] = dup ? top : t;
        }
        if (a.idx_out) {
#pragma unroll
            for (int i = 0; i < K; ++i) a.idx_out[d * K + i] = a.win_reads ? a.win_reads[base + pick[i]] : (int32_t)(base + pick[i]);
        }
        if (a.pick_out) {
#pragma unroll
            for (int i = 0; i < K; ++i) a.pick_out[d * K + i] = (int32_t)pick[i];
        }
    }
}

}  // namespace ldpgta
