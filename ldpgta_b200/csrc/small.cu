// small.cu -- kernels and launchers for draws of 2..4 reads (k_small.cuh).
#include <algorithm>
#include <stdlib.h>

#include "common.cuh"
#include "k_small.cuh"
#include "k_mid.cuh"

namespace ldpgta {

template <int NR, int LPD, int WPL, int KIND>
static int launch_small_pipe(ldpgta_ctx *c, SmallArgs &a)
{
    using P = SmallPipe<NR, WPL>;
    // staging by TMA (gather4 / bulk copies; default) or by per-lane cp.async (LDPGTA_SMALL_CPASYNC=1): measured equal for one
    // group (0.3748 vs 0.3752 ms per configs[1] step), 12 % slower for two groups -- see k_small.cuh and DESIGN.md
    static const bool use_ca = [] { const char *e = getenv("LDPGTA_SMALL_CPASYNC"); return e && *e && *e != '0'; }();
    auto kern = use_ca ? k_small<NR, LPD, WPL, KIND, true> : k_small<NR, LPD, WPL, KIND, false>;
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
static int launch_small_k(ldpgta_ctx *c, SmallArgs &a)
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
static int launch_small_n(ldpgta_ctx *c, SmallArgs &a, int kind)
{
    if (kind == LDPGTA_HOMOGENEOUS) return launch_small_k<NR, LDPGTA_HOMOGENEOUS>(c, a);
    if (kind == LDPGTA_RECENT_ADMIXTURE) return launch_small_k<NR, LDPGTA_RECENT_ADMIXTURE>(c, a);
    return launch_small_k<NR, LDPGTA_DISTANT_ADMIXTURE>(c, a);
}

// Tensor map over group g's read masks: rank 2, [n_reads][wp] uint64, box = {wp, 1} (tile::gather4 fetches four
// such boxes).  Encoded through the driver entry point so that the library does not link libcuda.
int make_read_mask_tmap(ldpgta_ctx *c, int g)
{
    c->tmap_ok[g] = false;
    if (!c->read_masks[g] || c->n_reads_g[g] <= 0 || c->wp[g] > 256) return 0;
    typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_fn encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
            cudaGetLastError();
            return 0;
        }
        encode = (encode_fn)fn;
    }
    const cuuint64_t gdim[2] = {(cuuint64_t)c->wp[g], (cuuint64_t)c->n_reads_g[g]};
    const cuuint64_t gstride[1] = {(cuuint64_t)c->wp[g] * 8};
    const cuuint32_t box[2] = {(cuuint32_t)c->wp[g], 1};
    const cuuint32_t estride[2] = {1, 1};
    const CUresult r = encode(&c->tmap[g], CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, (void *)c->read_masks[g], gdim, gstride, box, estride,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    c->tmap_ok[g] = (r == CUDA_SUCCESS);
    return 0;
}

template <int NR, int W64>
static int launch_mid_t(ldpgta_ctx *c, SmallArgs &a)
{
    using P = MidPipe<NR, W64>;
    auto kern = k_mid<NR, W64>;
    LDPGTA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, P::SMEM));
    const int64_t tiles = (a.n_draws + 1) / 2;
    const int64_t want = (tiles + P::NWARPS - 1) / P::NWARPS;
    const unsigned blocks = (unsigned)std::min<int64_t>(want, c->sm_count);
    kern<<<blocks, P::NWARPS * 32, P::SMEM, c->stream>>>(a);
    c->launches++;
    LDPGTA_CUDA(cudaGetLastError());
    return 0;
}

// 5 or 6 reads, one population of at most 65535 haplotypes and rows of at most 80 words: the register kernel.
// Returns 1 when the draw shape is outside that envelope (the caller then uses the table kernel).
int launch_mid(ldpgta_ctx *c, SmallArgs &a, int n)
{
    a.use_gather4 = 0;
    if (c->G != 1 || c->nhap[0] > 65535u || c->wp[0] > 80 || getenv("LDPGTA_NO_MID")) return 1;
    const int wp = c->wp[0];
    if (n == 5) return wp <= 32 ? launch_mid_t<5, 2>(c, a) : wp <= 48 ? launch_mid_t<5, 3>(c, a) : launch_mid_t<5, 5>(c, a);
    if (n == 6) return wp <= 32 ? launch_mid_t<6, 2>(c, a) : wp <= 48 ? launch_mid_t<6, 3>(c, a) : launch_mid_t<6, 5>(c, a);
    return 1;
}

int launch_small(ldpgta_ctx *c, SmallArgs &a, int n, int kind)
{
    a.use_gather4 = 0;
    if (n == 4 && !getenv("LDPGTA_NO_GATHER4") && a.pv.masks[0] == c->read_masks[0]) {
        bool ok = true;
        // each draw's four rows land at draw * 4 * row_bytes in the warp's buffer: TMA needs that 128-byte aligned
        for (int g = 0; g < c->G && g < 2; ++g) { ok = ok && c->tmap_ok[g] && (c->wp[g] % 4 == 0); a.tmap[g] = c->tmap[g]; }
        a.use_gather4 = (ok && c->G <= 2) ? 1 : 0;
    }
    if (n == 2) return launch_small_n<2>(c, a, kind);
    if (n == 3) return launch_small_n<3>(c, a, kind);
    return launch_small_n<4>(c, a, kind);
}

}  // namespace ldpgta
