/*
 * oracle/oracle_c.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the likelihood path on packed 64-bit rows, used only by
 * tests/, __graft_entry__.smoke() and bench.py as the CHECKER for sizes where
 * the CPython oracle (oracle/ldpgta_oracle.py) is too slow (n = 10..16 reads per
 * draw, thousands of draws).  Parity status: PINNED -- tests/test_oracle_c.py
 * checks it against the golden vectors made by the unmodified reference
 * (tests/golden/) and against the CPython oracle.
 *
 * What it restates (paths into /root/reference):
 *   joint counts     HOMOGENOUES_MODELS.py:129-163  (every subset, AND + popcount)
 *   general n        HOMOGENOUES_MODELS.py:185-225, RECENT_ADMIXTURE_MODELS.py:198-247,
 *                    DISTANT_ADMIXTURE_MODELS.py:215-256 with the model tables of
 *                    MAKE_STATISTICAL_MODEL.py:24-102 written as sums over set
 *                    partitions (weights 3,6,6 / 2^n+1, 2^|A|+2^|B| / 2,2 over 3^n, 3^n, 2^n)
 *   n = 2,3,4        the closed forms :227-270 (normalise first, then the same partition sums)
 *   LLR              ANEUPLOIDY_TEST.py:78-90
 * Integer class sums are exact (unsigned __int128); the few final operations are
 * carried in long double and rounded once to double.
 *
 * Build: gcc -O2 -shared -fPIC -o oracle/liboracle.so oracle/oracle_c.c -lm   (see oracle/Makefile)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

typedef unsigned __int128 u128;

/* counts[S] for S = 0..2^n-1 (counts[0] = 0), masks = n rows of W words */
void oracle_joint_counts(const uint64_t *masks, int n, int W, uint32_t *counts)
{
    size_t size = (size_t)1 << n;
    uint64_t *inter = (uint64_t *)malloc(size * (size_t)W * sizeof(uint64_t));
    counts[0] = 0;
    for (size_t s = 1; s < size; ++s) {
        int low = __builtin_ctzll(s);
        size_t rest = s & (s - 1);
        const uint64_t *m = masks + (size_t)low * W;
        uint64_t *dst = inter + s * W;
        uint32_t c = 0;
        if (rest == 0) {
            for (int w = 0; w < W; ++w) { dst[w] = m[w]; c += (uint32_t)__builtin_popcountll(dst[w]); }
        } else {
            const uint64_t *p = inter + rest * W;
            for (int w = 0; w < W; ++w) { dst[w] = p[w] & m[w]; c += (uint32_t)__builtin_popcountll(dst[w]); }
        }
        counts[s] = c;
    }
    free(inter);
}

static long double ipow(long double b, int e) { long double r = 1; while (e-- > 0) r *= b; return r; }

/* homogeneous, integer counts (general formula, any n >= 2) */
void oracle_lik_homogeneous(const uint32_t *F, int n, uint32_t N, double *out)
{
    uint32_t full = ((uint32_t)1 << n) - 1, top = (uint32_t)1 << (n - 1);
    u128 two = 0, two_sph = 0, three = 0;
    for (uint32_t a = top; a < full; ++a) {              /* a owns the last read */
        uint32_t rest = full ^ a;
        u128 prod = (u128)F[a] * F[rest];
        two += prod;
        two_sph += prod * (((u128)1 << __builtin_popcount(a)) + ((u128)1 << __builtin_popcount(rest)));
        uint32_t low = rest & (~rest + 1);
        u128 inner = 0;
        for (uint32_t b = (rest - 1) & rest; b; b = (b - 1) & rest)
            if (b & low) inner += (u128)F[b] * F[rest ^ b];
        three += (u128)F[a] * inner;
    }
    long double Nl = N, f = (long double)F[full] / Nl;
    long double t2 = (long double)two / (Nl * Nl), t2s = (long double)two_sph / (Nl * Nl);
    long double t3 = (long double)three / (Nl * Nl * Nl);
    out[0] = (double)f;
    out[1] = (double)((2 * f + 2 * t2) / ipow(2, n));
    out[2] = (double)(((ipow(2, n) + 1) * f + t2s) / ipow(3, n));
    out[3] = (double)((3 * f + 6 * t2 + 6 * t3) / ipow(3, n));
}

/* distant admixture / generic float frequencies e[S] */
void oracle_lik_float(const double *e, int n, double *out)
{
    uint32_t full = ((uint32_t)1 << n) - 1, top = (uint32_t)1 << (n - 1);
    long double two = 0, two_sph = 0, three = 0;
    for (uint32_t a = top; a < full; ++a) {
        uint32_t rest = full ^ a;
        long double prod = (long double)e[a] * e[rest];
        two += prod;
        two_sph += prod * (ipow(2, __builtin_popcount(a)) + ipow(2, __builtin_popcount(rest)));
        uint32_t low = rest & (~rest + 1);
        long double inner = 0;
        for (uint32_t b = (rest - 1) & rest; b; b = (b - 1) & rest)
            if (b & low) inner += (long double)e[b] * e[rest ^ b];
        three += (long double)e[a] * inner;
    }
    long double f = e[full];
    out[0] = (double)f;
    out[1] = (double)((2 * f + 2 * two) / ipow(2, n));
    out[2] = (double)(((ipow(2, n) + 1) * f + two_sph) / ipow(3, n));
    out[3] = (double)((3 * f + 6 * two + 6 * three) / ipow(3, n));
}

/* recent admixture: F counts in group 0 (M haplotypes), G in group 1 (N haplotypes) */
void oracle_lik_recent(const uint32_t *F, const uint32_t *G, int n, uint32_t M, uint32_t N, double *out)
{
    uint32_t full = ((uint32_t)1 << n) - 1, top = (uint32_t)1 << (n - 1);
    long double Ml = M, Nl = N;
    u128 ff = 0, gg = 0, fg = 0, fg_sph = 0, ffg = 0, ggf = 0;
    for (uint32_t a = top; a < full; ++a) {
        uint32_t rest = full ^ a;
        u128 cross = (u128)F[a] * G[rest] + (u128)G[a] * F[rest];
        ff += (u128)F[a] * F[rest];
        gg += (u128)G[a] * G[rest];
        fg += cross;
        fg_sph += cross * (((u128)1 << __builtin_popcount(a)) + ((u128)1 << __builtin_popcount(rest)));
        uint32_t low = rest & (~rest + 1);
        u128 in_ff = 0, in_gg = 0, in_fg = 0;
        for (uint32_t b = (rest - 1) & rest; b; b = (b - 1) & rest)
            if (b & low) {
                uint32_t c = rest ^ b;
                in_ff += (u128)F[b] * F[c];
                in_gg += (u128)G[b] * G[c];
                in_fg += (u128)F[b] * G[c] + (u128)G[b] * F[c];
            }
        /* unordered {a,b,c}: two parents' worth of F and one of G, and vice versa */
        ffg += (u128)G[a] * in_ff + (u128)F[a] * in_fg;
        ggf += (u128)F[a] * in_gg + (u128)G[a] * in_fg;
    }
    long double s = (long double)F[full] / Ml + (long double)G[full] / Nl;
    long double cross = (long double)fg / (Ml * Nl), cross_sph = (long double)fg_sph / (Ml * Nl);
    long double two_bph = (long double)ff / (Ml * Ml) + (long double)gg / (Nl * Nl) + 2 * cross;
    long double three = (long double)ffg / (Ml * Ml * Nl) + (long double)ggf / (Ml * Nl * Nl);
    out[0] = (double)(s / 2);
    out[1] = (double)((s + cross) / ipow(2, n));
    out[2] = (double)(((ipow(2, n) + 1) * s + cross_sph) / (2 * ipow(3, n)));
    out[3] = (double)((3 * s / 2 + two_bph + three) / ipow(3, n));
}

/*
 * A batch of draws on packed read masks, the shape the CUDA path consumes.
 *   masks[g]      pointers to [n_reads][row_words] rows of group g
 *   read_idx      concatenated read indices, draw d owns [offsets[d], offsets[d+1])
 *   kind          0 homogeneous, 1 recent admixture, 2 distant admixture
 *   proportions   kind 2 only, one per group
 *   out           [n_draws][4] = monosomy, disomy, SPH, BPH
 * The reference's dispatch is kept: n in {2,3,4} normalises the counts first
 * (closed forms), any other n uses integer counts.
 */
int oracle_batch(const uint64_t *const *masks, int n_groups, const uint32_t *n_hap, int row_words,
                 const int32_t *read_idx, const int64_t *offsets, int64_t n_draws, int kind,
                 const double *proportions, double *out)
{
    for (int64_t d = 0; d < n_draws; ++d) {
        int n = (int)(offsets[d + 1] - offsets[d]);
        if (n < 2 || n > 20) return -1;
        size_t size = (size_t)1 << n;
        uint32_t *cnt = (uint32_t *)malloc(size * n_groups * sizeof(uint32_t));
        uint64_t *rows = (uint64_t *)malloc((size_t)n * row_words * sizeof(uint64_t));
        for (int g = 0; g < n_groups; ++g) {
            for (int i = 0; i < n; ++i)
                memcpy(rows + (size_t)i * row_words, masks[g] + (size_t)read_idx[offsets[d] + i] * row_words,
                       row_words * sizeof(uint64_t));
            oracle_joint_counts(rows, n, row_words, cnt + g * size);
        }
        double *o = out + 4 * d;
        if (kind == 0 && n > 4) {
            oracle_lik_homogeneous(cnt, n, n_hap[0], o);
        } else if (kind == 1 && n > 4) {
            oracle_lik_recent(cnt, cnt + size, n, n_hap[0], n_hap[1], o);
        } else {
            /* float frequencies: closed forms (all kinds) and distant admixture */
            double *e = (double *)malloc(size * sizeof(double));
            double *e2 = (double *)malloc(size * sizeof(double));
            for (size_t s = 1; s < size; ++s) {
                if (kind == 2) {
                    double acc = 0;
                    for (int g = 0; g < n_groups; ++g) acc += proportions[g] * cnt[g * size + s] / n_hap[g];
                    e[s] = acc;
                } else {
                    e[s] = (double)cnt[s] / n_hap[0];
                    if (kind == 1) e2[s] = (double)cnt[size + s] / n_hap[1];
                }
            }
            if (kind == 1) {
                /* closed recent forms = the recent formula on normalised floats */
                uint32_t full = (uint32_t)size - 1, top = (uint32_t)1 << (n - 1);
                long double cross = 0, cross_sph = 0, two_bph = 0, three = 0;
                for (uint32_t a = top; a < full; ++a) {
                    uint32_t rest = full ^ a;
                    long double c = (long double)e[a] * e2[rest] + (long double)e2[a] * e[rest];
                    cross += c;
                    cross_sph += c * (ipow(2, __builtin_popcount(a)) + ipow(2, __builtin_popcount(rest)));
                    two_bph += (long double)e[a] * e[rest] + (long double)e2[a] * e2[rest] + 2 * c;
                    uint32_t low = rest & (~rest + 1);
                    for (uint32_t b = (rest - 1) & rest; b; b = (b - 1) & rest)
                        if (b & low) {
                            uint32_t cc = rest ^ b;
                            long double in_ff = (long double)e[b] * e[cc], in_gg = (long double)e2[b] * e2[cc];
                            long double in_fg = (long double)e[b] * e2[cc] + (long double)e2[b] * e[cc];
                            three += e2[a] * in_ff + e[a] * in_fg + e[a] * in_gg + e2[a] * in_fg;
                        }
                }
                long double s = (long double)e[full] + e2[full];
                o[0] = (double)(s / 2);
                o[1] = (double)((s + cross) / ipow(2, n));
                o[2] = (double)(((ipow(2, n) + 1) * s + cross_sph) / (2 * ipow(3, n)));
                o[3] = (double)((3 * s / 2 + two_bph + three) / ipow(3, n));
            } else {
                oracle_lik_float(e, n, o);
            }
            free(e); free(e2);
        }
        free(cnt); free(rows);
    }
    return 0;
}

/* ANEUPLOIDY_TEST.py:78-90 */
double oracle_llr(double y, double x)
{
    if (x != 0 && y != 0) return log(y / x);
    if (x != 0) return -1.23456789;
    if (y != 0) return +1.23456789;
    return 0;
}
