// poly_closed.cuh -- likelihood polynomials for draws of 2, 3 and 4 reads on normalised
// frequencies (the reference's likelihoods2/3/4: HOMOGENOUES_MODELS.py:227-270,
// DISTANT_ADMIXTURE_MODELS.py:258-298, RECENT_ADMIXTURE_MODELS.py:249-312).
//
// The single-population forms are evaluated in the operation order CPython uses for the
// reference's expressions (left-to-right, no fused multiply-add: this translation unit is
// compiled with -fmad=false), so they agree with the reference to the last bit.  The
// two-population forms are evaluated as sums over set partitions; they agree to ~1 ulp.
#pragma once
#include "common.cuh"

namespace ldpgta {

// f[S], S = 1 .. 2^NR - 1; out = monosomy, disomy, SPH, BPH
template <int NR>
__device__ __forceinline__ void closed_single(const double *f, double *out)
{
    if (NR == 2) {
        const double a = f[1], b = f[2], ab = f[3];
        out[0] = ab;
        out[1] = (ab + a * b) / 2.0;
        out[2] = (5.0 * ab + 4.0 * a * b) / 9.0;
        out[3] = (ab + 2.0 * a * b) / 3.0;
    } else if (NR == 3) {
        const double a = f[1], b = f[2], ab = f[3], c = f[4], ac = f[5], bc = f[6], abc = f[7];
        out[0] = abc;
        out[1] = (abc + ab * c + ac * b + bc * a) / 4.0;
        out[2] = abc / 3.0 + 2.0 * (ab * c + ac * b + bc * a) / 9.0;
        out[3] = (abc + 2.0 * (ab * c + ac * b + bc * a + a * b * c)) / 9.0;
    } else {
        const double a = f[1], b = f[2], c = f[4], d = f[8];
        const double ab = f[3], ac = f[5], ad = f[9], bc = f[6], bd = f[10], cd = f[12];
        const double abc = f[7], abd = f[11], acd = f[13], bcd = f[14], abcd = f[15];
        out[0] = abcd;
        out[1] = (abcd + abc * d + bcd * a + acd * b + abd * c + ab * cd + ad * bc + ac * bd) / 8.0;
        out[2] = (17.0 * abcd + 10.0 * (abc * d + bcd * a + acd * b + abd * c) + 8.0 * (ab * cd + ad * bc + ac * bd)) / 81.0;
        out[3] = (abcd + 2.0 * (ab * c * d + a * bd * c + a * bc * d + ac * b * d + a * b * cd + ad * b * c + abc * d +
                                a * bcd + acd * b + abd * c + ab * cd + ad * bc + ac * bd)) / 27.0;
    }
}

// Two populations, each parent drawn from one of them: f in group 0, g in group 1.
template <int NR>
__device__ __forceinline__ void closed_recent(const double *f, const double *g, double *out)
{
    constexpr int FULL = (1 << NR) - 1, TOP = 1 << (NR - 1);
    double cross = 0, cross_sph = 0, two_bph = 0, three = 0;
#pragma unroll
    for (int a = TOP; a < FULL; ++a) {
        const int rest = FULL ^ a;
        const double c = f[a] * g[rest] + g[a] * f[rest];
        cross += c;
        cross_sph += c * (double)((1 << __popc(a)) + (1 << __popc(rest)));
        two_bph += f[a] * f[rest] + g[a] * g[rest] + 2.0 * c;
        const int low = rest & -rest;
#pragma unroll
        for (int b = 1; b < FULL; ++b) {
            if ((b & rest) == b && (b & low) && b != rest) {
                const int cc = rest ^ b;
                const double in_ff = f[b] * f[cc], in_gg = g[b] * g[cc], in_fg = f[b] * g[cc] + g[b] * f[cc];
                three += g[a] * in_ff + f[a] * in_fg + f[a] * in_gg + g[a] * in_fg;
            }
        }
    }
    const double s = f[FULL] + g[FULL];
    double p3 = 1.0;
#pragma unroll
    for (int i = 0; i < NR; ++i) p3 *= 3.0;
    out[0] = s / 2.0;
    out[1] = (s + cross) / (double)(1 << NR);
    out[2] = ((double)((1 << NR) + 1) * s + cross_sph) / (2.0 * p3);
    out[3] = (3.0 * s / 2.0 + two_bph + three) / p3;
}

}  // namespace ldpgta
