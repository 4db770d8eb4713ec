// poly_finish.cuh -- the last few float operations of likelihoods() (HOMOGENOUES_MODELS.py:196-222,
// RECENT_ADMIXTURE_MODELS.py:209-245, DISTANT_ADMIXTURE_MODELS.py:226-254) from the exact partition sums,
// in the reference's operation order.  Shared by k_general and k_lanes.
#pragma once
#include "common.cuh"

namespace ldpgta {

struct PolyScale {
    double p2n1, p2n, p3n1, p3n;     // 2^(n-1), 2^n, 3^(n-1), 3^n
    u64 sph_full;                    // 2^n + 1
};
__device__ __forceinline__ PolyScale poly_scale(int n)
{
    PolyScale s;
    s.p2n1 = (double)(1u << (n - 1));
    s.p3n1 = 1.0;
    for (int i = 1; i < n; ++i) s.p3n1 *= 3.0;
    s.p3n = 3.0 * s.p3n1;
    s.p2n = 2.0 * s.p2n1;
    s.sph_full = (1ull << n) + 1ull;
    return s;
}

// two = sum {A,B} F[A]F[B]; two_sph = the same weighted by 2^|A| + 2^|B|; three = sum {A,B,C} F[A]F[B]F[C]
__device__ __forceinline__ void finish_homogeneous(const PolyScale &s, double N, u64 Ff, double two, double two_sph, double three,
                                                   double *res)
{
    res[0] = (double)Ff / N;
    res[1] = (double)Ff / (s.p2n1 * N) + two / s.p2n1 / (N * N);
    res[2] = (double)(Ff * s.sph_full) / (s.p3n * N) + two_sph / s.p3n / (N * N);
    res[3] = (double)Ff / (s.p3n1 * N) + (2.0 * two) / s.p3n1 / (N * N) + (2.0 * three) / s.p3n1 / (N * N * N);
}

__device__ __forceinline__ void finish_recent(const PolyScale &s, double M, double N, u64 Ff, u64 Gf, double ff, double gg, double fg,
                                              double fg_sph, double ffg, double ggf, double *res)
{
    const double t = (double)Ff / M + (double)Gf / N;
    res[0] = t / 2.0;
    res[1] = t / s.p2n + fg / s.p2n1 / (2.0 * M * N);
    res[2] = t * (double)s.sph_full / (2.0 * s.p3n) + fg_sph / s.p3n / (2.0 * M * N);
    res[3] = t / (2.0 * s.p3n1)
           + (ff / (M * M) + gg / (N * N) + 2.0 * fg / (M * N)) * 2.0 / s.p3n1 / 6.0
           + (ffg / M + ggf / N) * 2.0 / s.p3n1 / (6.0 * M * N);
}

__device__ __forceinline__ void finish_distant(const PolyScale &s, double Ef, double two, double two_sph, double three, double *res)
{
    res[0] = Ef;
    res[1] = Ef / s.p2n1 + two / s.p2n1;
    res[2] = (double)s.sph_full / s.p3n * Ef + two_sph / s.p3n;
    res[3] = Ef / s.p3n1 + 2.0 * two / s.p3n1 + 2.0 * three / s.p3n1;
}

// two-group distant admixture from the integer sums of poly_distant2_blocked (poly_general.cuh); Ef = e(full)
__device__ __forceinline__ void finish_distant2(const PolyScale &s, double alpha, double beta, double Ef, double ff, double gg, double fg,
                                                double ff_sph, double gg_sph, double fg_sph, double s0, double s1, double s2, double s3,
                                                double *res)
{
    const double a2 = alpha * alpha, ab = alpha * beta, b2 = beta * beta;
    const double two = a2 * ff + ab * fg + b2 * gg;
    const double two_sph = a2 * ff_sph + ab * fg_sph + b2 * gg_sph;
    const double three = a2 * alpha * s0 + a2 * beta * s1 + alpha * b2 * s2 + b2 * beta * s3;
    finish_distant(s, Ef, two, two_sph, three, res);
}

}  // namespace ldpgta
