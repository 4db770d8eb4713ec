// poly_general.cuh -- the partition polynomials of likelihoods() for 5..16 reads, evaluated from
// the joint-frequency table(s) in shared memory (HOMOGENOUES_MODELS.py:185-225,
// RECENT_ADMIXTURE_MODELS.py:198-247, DISTANT_ADMIXTURE_MODELS.py:215-256; weights of
// MAKE_STATISTICAL_MODEL.py:24-102 in closed form).
//
// The three-block sum  S3 = sum over unordered {A,B,C} of F[A] F[B] F[C]  has S(n,3) ~ 3^n/6 terms
// (7.1 M at n = 16) and is the expensive part.  Let A be the block that owns the last read and
// U' the other n-1 reads: S3 = 1/2 * sum over colourings (a,b,c) of U' of G[a] F[b] F[c] with
// G[a] = F[last | a] and F[empty] = 0.  U' is split into its 4 low reads (Lo) and h = n-5 high
// reads (Hi); a table row is the 16 entries that share the high part.  For one colouring
// (a_h,b_h,c_h) of Hi all 81 colourings of Lo are a trilinear form of three rows,
//      T(X,Y,Z) = sum_a X[a] * (Y*Z)[Lo\a],   (Y*Z)[s] = sum_{b subset s} Y[b] Z[s\b],
// evaluated from registers with compile-time indices: 97 multiply-adds for 81 terms, no subset
// enumeration in the inner loop.  The colourings of Hi with b_h < c_h (the swap gives the same
// value) come from a list built once per h on the host; b_h = c_h = empty is the one extra
// term with weight 1/2.  Two-block sums are plain coalesced passes over the table.
// The routines do not depend on the table kernel's tile shape: they are compiled once in
// poly_general.cu and linked (relocatable device code) into every k_general instance.
#pragma once
#include "common.cuh"

namespace ldpgta {

// per-thread partials: two, two_sph (64-bit) and three (128-bit); the caller sums them across the block;
// ACC = uint32_t when 16 * N^2 < 2^32 (every group has at most 16383 haplotypes), else u64.
//
// runs = true: `terms` is a list of RUNS of colourings instead (general.cu: ensure_run_terms).  T(X,Y,Z) is linear in the
// convolution, and the colourings that share a_h share X, so their convolutions can be summed first and dotted with X once:
// a run a_h | b_base << 16 | l << 32 stands for the colourings b_h = b_base | x over the subsets x of l (up to three of the
// lowest reads outside a_h), and costs 81 IMADs per colouring plus one X row and 16 wide multiply-adds per run instead of
// per colouring -- the polynomial is bound by instruction issue, and this is 1.4x fewer instructions.  The 32-bit sums hold
// 2^|l| convolutions, which is why the launcher bounds |l| by the panel size.
template <typename CT, typename ACC>
__device__ void poly_homog_blocked(const CT *F, int n, const u64 *terms, int64_t n_terms, int tid, int T,
                                   u64 &two, u64 &two_sph, u128 &three, bool runs = false);
template <typename CT, typename ACC>
__device__ void poly_recent_blocked(const CT *F, const CT *G, int n, const u64 *terms, int64_t n_terms, int tid, int T,
                                    u64 &ff, u64 &gg, u64 &fg, u64 &fg_sph, u128 &ffg, u128 &ggf, bool runs = false);
// Distant admixture of TWO groups as an integer form.  e(S) = alpha F[S] + beta G[S] (alpha = p_0 / N_0, beta = p_1 / N_1) is
// linear in the two count tables, so the sums of products of e over the partitions expand into the exact integer sums that the
// recent-admixture routine already forms, plus the two pure ones:
//     sum e_A e_B e_C = alpha^3 S0 + alpha^2 beta S1 + alpha beta^2 S2 + beta^3 S3,   S_k = triples with k factors from G
//     sum e_A e_R     = alpha^2 ff + alpha beta fg + beta^2 gg                       (and the same with the SPH weights)
// -- no e(S) table, no float64 convolutions (they ran at a quarter of the IMAD rate and, from 15 reads, out of L2), and the
// rounding of 3^(n-1)/2 floating-point additions is replaced by four conversions.  Same arguments as poly_recent_blocked.
struct Cubic2 {
    u64 ff, gg, fg, ff_sph, gg_sph, fg_sph;
    u128 s0, s1, s2, s3;
};
template <typename CT, typename ACC>
__device__ void poly_distant2_blocked(const CT *F, const CT *G, int n, const u64 *terms, int64_t n_terms, int tid, int T, Cubic2 &out,
                                      bool runs = false);

// Distant admixture of three or more groups, with the effective frequencies tabulated once per draw: etab[S] = sum_g p_g * cnt_g(S) / N_g
// (DISTANT_ADMIXTURE_MODELS.py:136-139) for all 2^n subsets, so that the polynomial reads doubles instead of redoing the
// conversions and divisions at every use (an entry is used ~1.5^n times).  fill: threads tid of T, the caller
// synchronises them before poly_distant_etab.
//
// Bank swizzle of e(S) (k_general, when the table is in shared memory).  The threads of a warp take consecutive colourings of the list, which
// share a_h and run through the subsets b_h of the reads outside it, so their Y rows (and Z rows) differ in the lowest free
// high reads.  A row of e(S) is 128 bytes -- all 32 banks -- so in subset order the eight 128-bit loads of a quarter warp
// hit the same bank group: 7.9 wavefronts per load over the real lists (tools/bank_conflicts.py), and the load/store unit,
// not the DMULs, bounds the polynomial (ncu at 13 reads: 85 % short_scoreboard).  Swizzled, the 16-byte unit u of row r sits
// at u ^ fold3(r): flipping any read moves the unit to another bank group, 1.7 wavefronts per load (12-14 reads: +15 %,
// +26 %, +50 % draws/s).  The same swizzle on the integer tables (32-byte rows, 3.5 -> 1.7 wavefronts) cost more in address arithmetic than
// it saved -- their polynomials are bound by instruction issue, not by the load/store unit -- and was dropped.
__host__ __device__ __forceinline__ uint32_t fold3(uint32_t x)
{
    x ^= x >> 12; x ^= x >> 6; x ^= x >> 3;
    return x & 7u;
}
__host__ __device__ __forceinline__ uint32_t etab_swizzle(uint32_t S) { return S ^ (fold3(S >> 4) << 1); }
template <typename CT>
__device__ void distant_fill_etab(const CT *ftab, const PanelView &pv, int n, double *etab, int tid, int T, bool swz = false);
__device__ void poly_distant_etab(const double *etab, int n, const u64 *terms, int64_t n_terms, int tid, int T,
                                  double &two, double &two_sph, double &three, bool swz = false);

}  // namespace ldpgta
