// lanes.cu -- launcher of k_lanes (k_lanes.cuh): draws of 7..12 reads, and 5..6 reads of two-group models.
#include <algorithm>
#include <stdlib.h>

#include "common.cuh"
#include "k_lanes.cuh"

namespace ldpgta {

template <int NR, int KL, typename CT>
static int launch_lanes_ct(ldpgta_ctx *c, const GeneralArgs &a)
{
    constexpr int DPW = 32 >> KL;
    static const int max_warps = [] { const char *e = getenv("LDPGTA_LANES_WARPS"); return e ? std::max(1, atoi(e)) : 12; }();
    const LanesSmem lay = lanes_smem_layout<NR, KL, CT>(a.pv, a.kind == LDPGTA_DISTANT_ADMIXTURE && a.pv.G > 2);
    const int fit = (c->max_smem_optin - 1024) / lay.per_warp;
    if (fit < 2) return 1;                                   // rows too long for two warps per SM: the table kernel takes it
    const int nwarps = std::min(std::min(max_warps, LanesWarps<NR>::MAX), fit);
    const size_t smem = (size_t)nwarps * lay.per_warp;
    auto kern = k_lanes<NR, KL, CT>;
    LDPGTA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t tiles = (a.n_draws + DPW - 1) / DPW;
    const unsigned blocks = (unsigned)std::max<int64_t>(1, std::min<int64_t>((tiles + nwarps - 1) / nwarps, c->sm_count));
    kern<<<blocks, 32 * nwarps, smem, c->stream>>>(a);
    c->launches++;
    LDPGTA_CUDA(cudaGetLastError());
    return 0;
}

// 16-bit count tables when every group fits (the usual case), else 32-bit
template <int NR, int KL>
static int launch_lanes_t(ldpgta_ctx *c, const GeneralArgs &a)
{
    static const bool force_u32 = getenv("LDPGTA_PLAN_U32") != nullptr;
    bool small = !force_u32;
    for (int g = 0; g < c->G; ++g) small = small && c->nhap[g] <= 65535u;
    return small ? launch_lanes_ct<NR, KL, uint16_t>(c, a) : launch_lanes_ct<NR, KL, uint32_t>(c, a);
}

int launch_lanes(ldpgta_ctx *c, GeneralArgs a)
{
    if (getenv("LDPGTA_NO_LANES")) return 1;
    // measured against the table kernel (tools/sweep_sizes.sh with LDPGTA_LANES_MAXN): ahead up to 12 reads for one group and
    // for recent admixture, up to 10 for distant admixture, whose float64 polynomial is too long for one warp beyond that
    static const int max_n_env = [] { const char *e = getenv("LDPGTA_LANES_MAXN"); return e ? atoi(e) : 0; }();      // tuning override
    // (two-group distant admixture runs the integer polynomial of recent admixture: poly_general.cuh, Cubic2)
    const bool float_poly = a.kind == LDPGTA_DISTANT_ADMIXTURE && c->G > 2;
    const int max_n = max_n_env ? std::min(max_n_env, 12) : float_poly ? 10 : 12;
    if (a.n < 5 || a.n > max_n) return 1;
    for (int g = 0; g < c->G; ++g)
        if (c->wp[g] * 64 > 65535) return 1;           // 16-bit packed partial counts, row bytes of one bulk copy
    const int h = a.n - 5;
    if (int rc = ensure_hi_terms(c, h)) return rc;
    a.terms = reinterpret_cast<const u64 *>(c->hi_terms[h]);
    a.n_terms = c->n_hi_terms[h];
    a.wide = 0;
    for (int g = 0; g < c->G; ++g) a.wide |= c->nhap[g] > (c->G == 2 ? 11585u : 16383u);   // 32-bit convolution sums: 16 (two groups: 32) products of two counts
    // 12 reads: the colourings in runs of four that share the X row (poly_general.cuh), when the 32-bit convolution sums hold
    // four convolutions (LDPGTA_PLAN_NO_RUNS=1 for comparison: +1.5 % one group, +3 % two groups; at 10 and 11 reads the 32
    // lanes of a draw have too few runs to share and lose 3-8 %)
    a.runs = 0;
    if (a.n >= 12 && !float_poly && !a.wide && !getenv("LDPGTA_PLAN_NO_RUNS")) {
        uint64_t nmax = 0;
        for (int g = 0; g < c->G; ++g) nmax = std::max<uint64_t>(nmax, c->nhap[g]);
        const uint64_t per_conv = (c->G == 2 ? 32u : 16u) * nmax * nmax;
        int bits = 0;
        while (bits < 2 && (per_conv << (bits + 1)) <= 0xffffffffull) ++bits;
        if (bits >= 1) {
            if (int rc = ensure_run_terms(c, h, bits)) return rc;
            a.terms = reinterpret_cast<const u64 *>(c->run_terms[h][bits]);
            a.n_terms = c->n_run_terms[h][bits];
            a.runs = 1;
        }
    }
    a.kl = a.kh = a.wc = 0; a.etab = 0; a.etab_scratch = nullptr; a.etab_stride = 0;
    a.table_scratch = nullptr; a.table_stride = 0;
    static const int alt = [] { const char *e = getenv("LDPGTA_LANES_ALT"); return e ? atoi(e) : 0; }();     // tuning: the other low/high split (measured slower: tools/sweep_sizes.sh)
    if (alt)
        switch (a.n) {
        case 6: return launch_lanes_t<6, 2>(c, a);
        case 7: return launch_lanes_t<7, 4>(c, a);
        case 8: return launch_lanes_t<8, 3>(c, a);
        case 9: return launch_lanes_t<9, 4>(c, a);
        case 10: return launch_lanes_t<10, 4>(c, a);
        case 11: return launch_lanes_t<11, 4>(c, a);
        }
    switch (a.n) {
    case 5: return launch_lanes_t<5, 3>(c, a);
    case 6: return launch_lanes_t<6, 3>(c, a);
    case 7: return launch_lanes_t<7, 3>(c, a);
    case 8: return launch_lanes_t<8, 4>(c, a);
    case 9: return launch_lanes_t<9, 4>(c, a);
    case 10: return launch_lanes_t<10, 5>(c, a);
    case 11: return launch_lanes_t<11, 5>(c, a);
    case 12: return launch_lanes_t<12, 5>(c, a);
    }
    return 1;
}

}  // namespace ldpgta
