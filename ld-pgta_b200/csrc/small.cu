// small.cu -- kernels and launchers for draws of 2..4 reads (k_small.cuh).
#include <algorithm>

#include "common.cuh"
#include "k_small.cuh"

namespace ldpgta {

template <int NR, int LPD, int WPL, int KIND>
static int launch_small_pipe(ldpgta_ctx *c, const SmallArgs &a)
{
    using P = SmallPipe<NR, WPL>;
    auto kern = k_small<NR, LPD, WPL, KIND>;
    LDPGTA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, P::SMEM));
    constexpr int DPW = 32 / LPD;
    const int64_t tiles = (a.n_draws + DPW - 1) / DPW;
    const int64_t want = (tiles + P::NWARPS - 1) / P::NWARPS;
    const unsigned blocks = (unsigned)std::min<int64_t>(want, c->sm_count);      // persistent: one CTA per SM
    kern<<<blocks, P::NWARPS * 32, P::SMEM, c->stream>>>(a);
    c->launches++;
    LDPGTA_CUDA(cudaGetLastError());
    return 0;
}

template <int NR, int KIND>
static int launch_small_k(ldpgta_ctx *c, const SmallArgs &a)
{
    int wmax = 0;
    for (int g = 0; g < c->G; ++g) wmax = std::max(wmax, c->wp[g]);
    // lanes per draw x words per lane: the smallest tile that covers a row in one pass
    if (wmax <= 16) return launch_small_pipe<NR, 4, 4, KIND>(c, a);
    if (wmax <= 32) return launch_small_pipe<NR, 8, 4, KIND>(c, a);
    if (wmax <= 40) return launch_small_pipe<NR, 4, 10, KIND>(c, a);
    if (wmax <= 64) return launch_small_pipe<NR, 8, 8, KIND>(c, a);
    if (wmax <= 80) return launch_small_pipe<NR, 8, 10, KIND>(c, a);
    if (wmax <= 128) return launch_small_pipe<NR, 16, 8, KIND>(c, a);
    // longer rows (more than 8192 haplotypes): unpipelined kernel looping over 256-word chunks
    const int threads = 128;
    const int64_t blocks = (a.n_draws * 32 + threads - 1) / threads;
    if (blocks > 0x7fffffffLL) return fail(LDPGTA_E_UNSUPPORTED, "batch too large for one launch");
    k_small_direct<NR, 32, 8, KIND><<<(unsigned)blocks, threads, 0, c->stream>>>(a);
    c->launches++;
    LDPGTA_CUDA(cudaGetLastError());
    return 0;
}

template <int NR>
static int launch_small_n(ldpgta_ctx *c, const SmallArgs &a, int kind)
{
    if (kind == LDPGTA_HOMOGENEOUS) return launch_small_k<NR, LDPGTA_HOMOGENEOUS>(c, a);
    if (kind == LDPGTA_RECENT_ADMIXTURE) return launch_small_k<NR, LDPGTA_RECENT_ADMIXTURE>(c, a);
    return launch_small_k<NR, LDPGTA_DISTANT_ADMIXTURE>(c, a);
}

int launch_small(ldpgta_ctx *c, const SmallArgs &a, int n, int kind)
{
    if (n == 2) return launch_small_n<2>(c, a, kind);
    if (n == 3) return launch_small_n<3>(c, a, kind);
    return launch_small_n<4>(c, a, kind);
}

}  // namespace ldpgta
