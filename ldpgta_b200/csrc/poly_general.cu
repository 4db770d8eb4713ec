// poly_general.cu -- definitions of the blocked partition polynomials declared in poly_general.cuh.
// Compiled with -maxrregcount=128 so that a 512-thread k_general block can call them.
#include "common.cuh"
#include "poly_general.cuh"

namespace ldpgta {

// 128-bit table loads.  The tables live in shared memory (k_lanes, k_general up to 16 reads) or in an L2-resident global
// slice (17-18 reads, two-group models at 16); through a generic pointer every row costs generic LD.E.128s, so the
// callers' pointer is classified once (SPACE: 0 generic, 1 shared) and shared tables are read with LDS.128.
template <int SPACE>
__device__ __forceinline__ uint4 load128(const void *p)
{
    if (SPACE == 1) {
        uint4 v;
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                     : "r"((uint32_t)__cvta_generic_to_shared(p)));
        return v;
    }
    return *reinterpret_cast<const uint4 *>(p);
}

template <typename CT, int SPACE>
__device__ __forceinline__ void load_row16(const CT *tab, uint32_t row, uint32_t *v);

template <typename CT, int SPACE>
struct RowLoader;

template <int SPACE>
struct RowLoader<uint32_t, SPACE> {
  static __device__ __forceinline__ void run(const uint32_t *tab, uint32_t row, uint32_t *v)
{
    const uint4 *p = reinterpret_cast<const uint4 *>(tab + ((size_t)row << 4));
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint4 x = load128<SPACE>(p + q);
        v[4 * q] = x.x; v[4 * q + 1] = x.y; v[4 * q + 2] = x.z; v[4 * q + 3] = x.w;
    }
}
};

template <int SPACE>
struct RowLoader<uint16_t, SPACE> {
  static __device__ __forceinline__ void run(const uint16_t *tab, uint32_t row, uint32_t *v)
{
    const uint4 *p = reinterpret_cast<const uint4 *>(tab + ((size_t)row << 4));
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const uint4 x = load128<SPACE>(p + q);
        const uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) { v[8 * q + 2 * e] = w[e] & 0xffffu; v[8 * q + 2 * e + 1] = w[e] >> 16; }
    }
}
};

template <typename CT, int SPACE>
__device__ __forceinline__ void load_row16(const CT *tab, uint32_t row, uint32_t *v)
{
    RowLoader<CT, SPACE>::run(tab, row, v);
}

// (Y*Z)[s] for the 16 subsets s of Lo; ACC = uint32_t when 16 * max(Y) * max(Z) < 2^32, else u64
template <typename ACC>
__device__ __forceinline__ void conv16(const uint32_t *Y, const uint32_t *Z, ACC *out)
{
#pragma unroll
    for (int s = 0; s < 16; ++s) {
        ACC acc = 0;
#pragma unroll
        for (int b = 0; b < 16; ++b)
            if ((b & s) == b) acc += (ACC)Y[b] * (ACC)Z[s ^ b];
        out[s] = acc;
    }
}

// out[s] += (Y*Z)[s]
template <typename ACC>
__device__ __forceinline__ void conv16_add(const uint32_t *Y, const uint32_t *Z, ACC *out)
{
#pragma unroll
    for (int s = 0; s < 16; ++s) {
        ACC acc = out[s];
#pragma unroll
        for (int b = 0; b < 16; ++b)
            if ((b & s) == b) acc += (ACC)Y[b] * (ACC)Z[s ^ b];
        out[s] = acc;
    }
}

template <typename ACC>
__device__ __forceinline__ u64 dot16(const uint32_t *X, const ACC *conv)
{
    u64 acc = 0;
#pragma unroll
    for (int a = 0; a < 16; ++a) acc += (u64)X[a] * (u64)conv[15 ^ a];
    return acc;
}

// ---- homogeneous ------------------------------------------------------------------------------
// returns per-thread partials: two, two_sph (u64) and three (u64; the caller sums them in 128 bits)
template <typename CT, typename ACC, int SPACE, bool RUNS>
__device__ __forceinline__ void poly_homog_impl(const CT *F, int n, const u64 *terms, int64_t n_terms, int tid, int T,
                                                u64 &two_out, u64 &two_sph_out, u128 &three_out)
{
    const int h = n - 5;
    const uint32_t top_row = 1u << h, usub = (1u << (n - 1)) - 1u, top = 1u << (n - 1);
    // local accumulators: the outputs are references into the caller's frame, and updating them in the loops costs a
    // generic load and store per term (the compiler cannot rule out aliasing with the table)
    u64 two = 0, two_sph = 0;
    u128 three = 0;
    for (uint32_t a = tid; a < usub; a += T) {                  // a != U': the other block is not empty
        const u64 prod = (u64)F[top | a] * (u64)F[usub ^ a];
        const int pa = __popc(a);
        two += prod;
        two_sph += prod * ((2ull << pa) + (1ull << (n - 1 - pa)));
    }
    u64 t_next = tid < n_terms ? terms[tid] : 0;                 // the colouring list is read one iteration ahead
    if (RUNS) {
        const uint32_t fullh = top_row - 1u;
        for (int64_t e = tid; e < n_terms; e += T) {
            const u64 t = t_next;
            if (e + T < n_terms) t_next = terms[e + T];
            const uint32_t ra = (uint32_t)(t & 0xffffu), base = (uint32_t)(t >> 16) & 0xffffu, l = (uint32_t)(t >> 32) & 0xffffu;
            const uint32_t rest = fullh ^ ra;
            uint32_t Y[16], Z[16];
            ACC c[16];
#pragma unroll
            for (int s = 0; s < 16; ++s) c[s] = 0;
            // the colourings of the run: b_h = base | x over the subsets x of l, c_h the others.  A fixed trip count with the
            // subsets formed from the single reads of l: the row addresses do not depend on the previous iteration, so the
            // loads are scheduled ahead of the products (a loop over x = (x - 1) & l exposed every load's latency)
            const uint32_t l0 = l & (0u - l), l12 = l ^ l0, l1 = l12 & (0u - l12), l2 = l12 ^ l1;
            const int cnt = 1 << __popc(l);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (i < cnt) {
                    const uint32_t rb = base | ((i & 1) ? l0 : 0u) | ((i & 2) ? l1 : 0u) | ((i & 4) ? l2 : 0u);
                    load_row16<CT, SPACE>(F, rb, Y);
                    load_row16<CT, SPACE>(F, rest ^ rb, Z);
                    conv16_add<ACC>(Y, Z, c);
                }
            }
            load_row16<CT, SPACE>(F, top_row + ra, Y);           // X, once per run
            three += dot16<ACC>(Y, c);
        }
    } else
    for (int64_t e = tid; e < n_terms; e += T) {
        const u64 t = t_next;
        if (e + T < n_terms) t_next = terms[e + T];
        uint32_t X[16], Y[16], Z[16];
        ACC c[16];
        load_row16<CT, SPACE>(F, top_row + (uint32_t)(t & 0xffffu), X);
        load_row16<CT, SPACE>(F, (uint32_t)(t >> 16) & 0xffffu, Y);
        load_row16<CT, SPACE>(F, (uint32_t)(t >> 32) & 0xffffu, Z);
        conv16<ACC>(Y, Z, c);
        three += dot16<ACC>(X, c);
    }
    if (tid == 0) {                                             // b_h = c_h = empty: both blocks inside Lo, weight 1/2
        uint32_t X[16], Y[16];
        ACC c[16];
        load_row16<CT, SPACE>(F, top_row + (top_row - 1u), X);
        load_row16<CT, SPACE>(F, 0u, Y);
        conv16<ACC>(Y, Y, c);
        three += dot16<ACC>(X, c) >> 1;
    }
    two_out = two; two_sph_out = two_sph; three_out = three;
}

template <typename CT, typename ACC>
__device__ void poly_homog_blocked(const CT *F, int n, const u64 *terms, int64_t n_terms, int tid, int T,
                                   u64 &two, u64 &two_sph, u128 &three, bool runs)
{
    if constexpr (sizeof(ACC) == 4) {                            // the launcher asks for runs with 32-bit sums only
        if (runs) {
            if (__isShared(F)) poly_homog_impl<CT, ACC, 1, true>(F, n, terms, n_terms, tid, T, two, two_sph, three);
            else poly_homog_impl<CT, ACC, 0, true>(F, n, terms, n_terms, tid, T, two, two_sph, three);
            return;
        }
    }
    if (__isShared(F)) poly_homog_impl<CT, ACC, 1, false>(F, n, terms, n_terms, tid, T, two, two_sph, three);
    else poly_homog_impl<CT, ACC, 0, false>(F, n, terms, n_terms, tid, T, two, two_sph, three);
}

// ---- recent admixture: F = group 0, G = group 1 ------------------------------------------------
// FULL: also the pure triples (S0 = FFF, S3 = GGG) and the SPH-weighted ff and gg, for the two-group distant model; ffg = S1
// (two factors from F), ggf = S2
template <typename CT, typename ACC, int SPACE, bool RUNS, bool FULL>
__device__ __forceinline__ void poly_recent_impl(const CT *F, const CT *G, int n, const u64 *terms, int64_t n_terms, int tid,
                                                 int T, Cubic2 &out)
{
    const int h = n - 5;
    const uint32_t top_row = 1u << h, usub = (1u << (n - 1)) - 1u, top = 1u << (n - 1);
    u64 ff = 0, gg = 0, fg = 0, fg_sph = 0, ff_sph = 0, gg_sph = 0;   // local accumulators (see poly_homog_blocked)
    u128 ffg = 0, ggf = 0, fff = 0, ggg = 0;
    for (uint32_t a = tid; a < usub; a += T) {
        const uint32_t A = top | a, R = usub ^ a;
        const u64 fa = F[A], ga = G[A], fr = F[R], gr = G[R];
        const u64 cross = fa * gr + ga * fr;
        const int pa = __popc(a);
        const u64 w = (2ull << pa) + (1ull << (n - 1 - pa));
        ff += fa * fr;
        gg += ga * gr;
        fg += cross;
        fg_sph += cross * w;
        if (FULL) { ff_sph += fa * fr * w; gg_sph += ga * gr * w; }
    }
    auto triple = [&](uint32_t ra, uint32_t rb, uint32_t rc, bool special) {
        // unordered {A,B,C}, A owns the last read: two parents from F and one from G, and vice versa.  Only one Y row and
        // one Z row are live at a time (a row is 16 registers; all six at once do not fit 128 registers with the products).
        uint32_t Xf[16], Xg[16], Y[16], Z[16];
        ACC c[16];
        u64 s_ffg, s_ggf, s_fff = 0, s_ggg = 0;
        load_row16<CT, SPACE>(F, top_row + ra, Xf); load_row16<CT, SPACE>(G, top_row + ra, Xg);
        // the two mixed convolutions Yf*Zg and Yg*Zf enter both sums with the same partner (Xf for ffg, Xg for ggf), so they
        // are accumulated into one array and dotted once: 4 convolutions + 4 dot products per colouring instead of 4 + 6
        // (two more row loads; 32 * max(F) * max(G) stays below 2^32 whenever ACC is 32-bit: groups of at most 11585)
        load_row16<CT, SPACE>(F, rb, Y); load_row16<CT, SPACE>(F, rc, Z);
        conv16<ACC>(Y, Z, c); s_ffg = dot16<ACC>(Xg, c);                                          // Yf * Zf
        if (FULL) s_fff = dot16<ACC>(Xf, c);
        load_row16<CT, SPACE>(G, rc, Z);
        conv16<ACC>(Y, Z, c);                                                                     // Yf * Zg
        load_row16<CT, SPACE>(G, rb, Y); load_row16<CT, SPACE>(F, rc, Z);
        conv16_add<ACC>(Y, Z, c); s_ffg += dot16<ACC>(Xf, c); s_ggf = dot16<ACC>(Xg, c);          // + Yg * Zf
        load_row16<CT, SPACE>(G, rc, Z);
        conv16<ACC>(Y, Z, c); s_ggf += dot16<ACC>(Xf, c);                                         // Yg * Zg
        if (FULL) s_ggg = dot16<ACC>(Xg, c);
        if (special) { s_ffg >>= 1; s_ggf >>= 1; s_fff >>= 1; s_ggg >>= 1; }
        ffg += s_ffg;
        ggf += s_ggf;
        if (FULL) { fff += s_fff; ggg += s_ggg; }
    };
    u64 t_next = tid < n_terms ? terms[tid] : 0;
    if (RUNS) {
        // runs of colourings that share a_h: three convolution sums (Yf*Zf, Yf*Zg + Yg*Zf, Yg*Zg) and four dot products per run
        const uint32_t fullh = top_row - 1u;
        for (int64_t e = tid; e < n_terms; e += T) {
            const u64 t = t_next;
            if (e + T < n_terms) t_next = terms[e + T];
            const uint32_t ra = (uint32_t)(t & 0xffffu), base = (uint32_t)(t >> 16) & 0xffffu, l = (uint32_t)(t >> 32) & 0xffffu;
            const uint32_t rest = fullh ^ ra;
            uint32_t Y[16], Z[16];
            ACC cff[16], cmix[16], cgg[16];
#pragma unroll
            for (int s = 0; s < 16; ++s) { cff[s] = 0; cmix[s] = 0; cgg[s] = 0; }
            const uint32_t l0 = l & (0u - l), l12 = l ^ l0, l1 = l12 & (0u - l12), l2 = l12 ^ l1;
            const int cnt = 1 << __popc(l);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (i < cnt) {
                    const uint32_t rb = base | ((i & 1) ? l0 : 0u) | ((i & 2) ? l1 : 0u) | ((i & 4) ? l2 : 0u), rc = rest ^ rb;
                    load_row16<CT, SPACE>(F, rb, Y); load_row16<CT, SPACE>(F, rc, Z);
                    conv16_add<ACC>(Y, Z, cff);                                                   // Yf * Zf
                    load_row16<CT, SPACE>(G, rc, Z);
                    conv16_add<ACC>(Y, Z, cmix);                                                  // Yf * Zg
                    load_row16<CT, SPACE>(G, rb, Y);
                    conv16_add<ACC>(Y, Z, cgg);                                                   // Yg * Zg
                    load_row16<CT, SPACE>(F, rc, Z);
                    conv16_add<ACC>(Y, Z, cmix);                                                  // Yg * Zf
                }
            }
            load_row16<CT, SPACE>(F, top_row + ra, Y); load_row16<CT, SPACE>(G, top_row + ra, Z); // Xf, Xg
            ffg += dot16<ACC>(Z, cff);
            ffg += dot16<ACC>(Y, cmix);
            ggf += dot16<ACC>(Z, cmix);
            ggf += dot16<ACC>(Y, cgg);
            if (FULL) { fff += dot16<ACC>(Y, cff); ggg += dot16<ACC>(Z, cgg); }
        }
    } else
    for (int64_t e = tid; e < n_terms; e += T) {
        const u64 t = t_next;
        if (e + T < n_terms) t_next = terms[e + T];
        triple((uint32_t)(t & 0xffffu), (uint32_t)(t >> 16) & 0xffffu, (uint32_t)(t >> 32) & 0xffffu, false);
    }
    if (tid == 0) triple(top_row - 1u, 0u, 0u, true);
    out.ff = ff; out.gg = gg; out.fg = fg; out.fg_sph = fg_sph; out.ff_sph = ff_sph; out.gg_sph = gg_sph;
    out.s0 = fff; out.s1 = ffg; out.s2 = ggf; out.s3 = ggg;
}

template <typename CT, typename ACC, bool FULL>
__device__ __forceinline__ void poly_recent_dispatch(const CT *F, const CT *G, int n, const u64 *terms, int64_t n_terms, int tid, int T,
                                                     Cubic2 &out, bool runs)
{
    const bool sh = __isShared(F) && __isShared(G);
    if constexpr (sizeof(ACC) == 4) {
        if (runs) {
            if (sh) poly_recent_impl<CT, ACC, 1, true, FULL>(F, G, n, terms, n_terms, tid, T, out);
            else poly_recent_impl<CT, ACC, 0, true, FULL>(F, G, n, terms, n_terms, tid, T, out);
            return;
        }
    }
    if (sh) poly_recent_impl<CT, ACC, 1, false, FULL>(F, G, n, terms, n_terms, tid, T, out);
    else poly_recent_impl<CT, ACC, 0, false, FULL>(F, G, n, terms, n_terms, tid, T, out);
}

template <typename CT, typename ACC>
__device__ void poly_recent_blocked(const CT *F, const CT *G, int n, const u64 *terms, int64_t n_terms, int tid, int T,
                                    u64 &ff, u64 &gg, u64 &fg, u64 &fg_sph, u128 &ffg, u128 &ggf, bool runs)
{
    Cubic2 c;
    poly_recent_dispatch<CT, ACC, false>(F, G, n, terms, n_terms, tid, T, c, runs);
    ff = c.ff; gg = c.gg; fg = c.fg; fg_sph = c.fg_sph; ffg = c.s1; ggf = c.s2;
}

template <typename CT, typename ACC>
__device__ void poly_distant2_blocked(const CT *F, const CT *G, int n, const u64 *terms, int64_t n_terms, int tid, int T, Cubic2 &out,
                                      bool runs)
{
    poly_recent_dispatch<CT, ACC, true>(F, G, n, terms, n_terms, tid, T, out, runs);
}

// ---- distant admixture from a tabulated e(S) ------------------------------------------------------
template <typename CT>
__device__ void distant_fill_etab(const CT *ftab, const PanelView &pv, int n, double *etab, int tid, int T, bool swz)
{
    const int G = pv.G;
    const uint32_t NS = 1u << n;
    for (uint32_t S = tid; S < NS; S += T) {
        double e = 0.0;
        for (int g = 0; g < G; ++g) e = e + pv.coef[g] * (double)ftab[((size_t)g << n) + S] / (double)pv.nhap[g];
        etab[swz ? etab_swizzle(S) : S] = e;
    }
}

template <int SPACE, bool SWZ>
__device__ __forceinline__ void poly_distant_etab_impl(const double *E, int n, const u64 *terms, int64_t n_terms, int tid, int T,
                                                       double &two_out, double &two_sph_out, double &three_out)
{
    const int h = n - 5;
    const uint32_t top_row = 1u << h, usub = (1u << (n - 1)) - 1u, top = 1u << (n - 1);
    auto at = [&](uint32_t S) { return E[SWZ ? etab_swizzle(S) : S]; };
    auto row = [&](uint32_t r, double *v) {
        const double *p = E + ((size_t)r << 4);
        const uint32_t m = SWZ ? fold3(r) : 0u;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const uint4 x = load128<SPACE>(p + 2 * (q ^ m));
            v[2 * q] = __hiloint2double((int)x.y, (int)x.x);
            v[2 * q + 1] = __hiloint2double((int)x.w, (int)x.z);
        }
    };
    double two = 0, two_sph = 0, three = 0;
    for (uint32_t a = tid; a < usub; a += T) {
        const double prod = at(top | a) * at(usub ^ a);
        const int pa = __popc(a);
        two += prod;
        two_sph += prod * (double)((2ull << pa) + (1ull << (n - 1 - pa)));
    }
    auto triple = [&](uint32_t ra, uint32_t rb, uint32_t rc) {
        double Y[16], Z[16], acc = 0.0;                      // the X row is read where it is used (three rows do not fit)
        row(rb, Y); row(rc, Z);
        const double *X = E + ((size_t)(top_row + ra) << 4);
        const uint32_t mx = SWZ ? fold3(top_row + ra) << 1 : 0u;
#pragma unroll
        for (int s = 0; s < 16; ++s) {
            double cv = 0.0;
#pragma unroll
            for (int b = 0; b < 16; ++b)
                if ((b & s) == b) cv += Y[b] * Z[s ^ b];
            acc += X[(15 ^ s) ^ mx] * cv;
        }
        return acc;
    };
    u64 t_next = tid < n_terms ? terms[tid] : 0;
    for (int64_t e = tid; e < n_terms; e += T) {
        const u64 t = t_next;
        if (e + T < n_terms) t_next = terms[e + T];
        three += triple((uint32_t)(t & 0xffffu), (uint32_t)(t >> 16) & 0xffffu, (uint32_t)(t >> 32) & 0xffffu);
    }
    if (tid == 0) three += 0.5 * triple(top_row - 1u, 0u, 0u);
    two_out = two; two_sph_out = two_sph; three_out = three;
}

__device__ void poly_distant_etab(const double *etab, int n, const u64 *terms, int64_t n_terms, int tid, int T,
                                  double &two, double &two_sph, double &three, bool swz)
{
    if (swz) {
        if (__isShared(etab)) poly_distant_etab_impl<1, true>(etab, n, terms, n_terms, tid, T, two, two_sph, three);
        else poly_distant_etab_impl<0, true>(etab, n, terms, n_terms, tid, T, two, two_sph, three);
    } else {
        if (__isShared(etab)) poly_distant_etab_impl<1, false>(etab, n, terms, n_terms, tid, T, two, two_sph, three);
        else poly_distant_etab_impl<0, false>(etab, n, terms, n_terms, tid, T, two, two_sph, three);
    }
}

#define LDPGTA_INST(CT)                                                                                              \
    template __device__ void poly_homog_blocked<CT, uint32_t>(const CT *, int, const u64 *, int64_t, int, int, u64 &, u64 &, u128 &, bool); \
    template __device__ void poly_homog_blocked<CT, u64>(const CT *, int, const u64 *, int64_t, int, int, u64 &, u64 &, u128 &, bool);      \
    template __device__ void poly_recent_blocked<CT, uint32_t>(const CT *, const CT *, int, const u64 *, int64_t, int, int, u64 &, u64 &, u64 &, u64 &, u128 &, u128 &, bool); \
    template __device__ void poly_recent_blocked<CT, u64>(const CT *, const CT *, int, const u64 *, int64_t, int, int, u64 &, u64 &, u64 &, u64 &, u128 &, u128 &, bool);      \
    template __device__ void poly_distant2_blocked<CT, uint32_t>(const CT *, const CT *, int, const u64 *, int64_t, int, int, Cubic2 &, bool); \
    template __device__ void poly_distant2_blocked<CT, u64>(const CT *, const CT *, int, const u64 *, int64_t, int, int, Cubic2 &, bool);      \
    template __device__ void distant_fill_etab<CT>(const CT *, const PanelView &, int, double *, int, int, bool);
LDPGTA_INST(uint16_t)
LDPGTA_INST(uint32_t)
#undef LDPGTA_INST

}  // namespace ldpgta
