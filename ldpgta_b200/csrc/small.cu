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

// ---- combined rows of a two-group panel (k_small_both) ----------------------------------------------------------
static __global__ void k_interleave_rows(const uint64_t *__restrict__ m0, const uint64_t *__restrict__ m1, int wp0, int wp1, int64_t n_reads,
                                         uint64_t *__restrict__ out)
{
    // one warp per read, 80 words per row: [group 0 | group 1 | zero]
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n_reads; r += warps)
        for (int w = lane; w < 80; w += 32)
            out[r * 80 + w] = w < wp0 ? m0[r * wp0 + w] : w < wp0 + wp1 ? m1[r * wp1 + (w - wp0)] : 0ull;
}

static int ensure_both(ldpgta_ctx *c)
{
    if (c->both_valid) return 0;
    const int64_t n = c->n_reads;
    const size_t need = std::max<size_t>((size_t)n * 640, 16);
    if (c->masks_both_bytes < need) {
        if (c->masks_both) { LDPGTA_CUDA(cudaStreamSynchronize(c->stream)); LDPGTA_CUDA(cudaFree(c->masks_both)); c->masks_both = nullptr; c->masks_both_bytes = 0; }
        const cudaError_t e = cudaMalloc((void **)&c->masks_both, need);
        if (e != cudaSuccess) { cudaGetLastError(); return 1; }           // no room for the second copy: the two-stage kernel runs
        c->masks_both_bytes = need;
    }
    k_interleave_rows<<<(unsigned)std::min<int64_t>((n * 32 + 255) / 256 + 1, (int64_t)c->sm_count * 16), 256, 0, c->stream>>>(
        c->read_masks[0], c->read_masks[1], c->wp[0], c->wp[1], n, c->masks_both);
    c->launches++;
    LDPGTA_CUDA(cudaGetLastError());
    // tensor map [n_reads][80] uint64, box = one row (tile::gather4 fetches four)
    c->both_tmap_ok = false;
    typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (n > 0 && cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && fn) {
        const cuuint64_t gdim[2] = {80, (cuuint64_t)n};
        const cuuint64_t gstride[1] = {640};
        const cuuint32_t box[2] = {80, 1};
        const cuuint32_t estride[2] = {1, 1};
        const CUresult r = ((encode_fn)fn)(&c->tmap_both, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, (void *)c->masks_both, gdim, gstride, box, estride,
                                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        c->both_tmap_ok = r == CUDA_SUCCESS;
    } else {
        cudaGetLastError();
    }
    c->both_valid = true;
    return 0;
}

// 4 reads per draw, two groups whose rows together fill the 80-word tile and whose boundary lies in the third slot row
// (group 0 of 32..47 words, e.g. 2504 + 2504 haplotypes): one gather per read instead of one per read and group.
// Returns 1 when the shape is outside that envelope.
template <int KIND>
static int launch_small_both(ldpgta_ctx *c, SmallArgs &a)
{
    static const bool off = [] { const char *e = getenv("LDPGTA_SMALL_NO_BOTH"); return e && *e && *e != '0'; }();
    const int wp0 = c->wp[0], wp1 = c->wp[1];
    if (off || c->G != 2 || wp0 + wp1 > 80 || wp0 + wp1 <= 64 || wp0 < 32 || wp0 >= 48 || c->nhap[0] > 65535u || c->nhap[1] > 65535u ||
        a.counts_out || a.pv.masks[0] != c->read_masks[0] || a.pv.masks[1] != c->read_masks[1] || a.n_draws < 4096)
        return 1;
    if (int rc = ensure_both(c)) return rc;
    using P = SmallPipe<4, 10>;
    SmallBothArgs b;
    b.base = a;
    b.base.use_gather4 = (c->both_tmap_ok && !getenv("LDPGTA_NO_GATHER4")) ? 1 : 0;
    b.both = c->masks_both;
    b.lsplit = wp0 / 2 - 16;                                   // slot row 2 holds slots 16..23
    b.tmap_both = c->tmap_both;
    auto kern = k_small_both<KIND, 2>;
    LDPGTA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, P::SMEM));
    const int64_t tiles = (a.n_draws + 3) / 4;
    const unsigned blocks = (unsigned)std::min<int64_t>((tiles + P::NWARPS - 1) / P::NWARPS, c->sm_count);
    kern<<<blocks, P::NWARPS * 32, P::SMEM, c->stream>>>(b);
    c->launches++;
    LDPGTA_CUDA(cudaGetLastError());
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

// the context's panel and the plan's largest window fit the window-resident kernel: one group of at most 65535 haplotypes in
// rows of exactly 80 words, 4 .. 4096 rows that fit the shared-memory block
bool small_windows_ok(const ldpgta_ctx *c, int64_t max_rows)
{
    return c->G == 1 && c->wp[0] == 80 && c->nhap[0] <= 65535u && max_rows >= 4 && max_rows <= 4096 &&
           SmallWinSmem::bytes((int)max_rows) <= c->max_smem_optin;
}

// Window-resident kernel for plans with many draws per window (k_small.cuh: k_small_win).  Returns 1 when the plan is outside
// its envelope (the caller launches k_small): one group of at most 65535 haplotypes in rows of exactly 80 words, 4 reads per
// draw, windows that are runs of consecutive rows and fit the shared-memory block.
int launch_small_windows(ldpgta_ctx *c, const int32_t *d_idx, int64_t n_draws, double *d_out, const int64_t *d_woff, const int64_t *d_doff,
                         int64_t n_windows, int64_t max_rows, const uint64_t *masks, const uint32_t *popc, bool local)
{
    if (!small_windows_ok(c, max_rows) || n_windows < 1) return 1;
    SmallArgs a;
    a.pv = make_view(c, nullptr);
    a.idx = d_idx; a.n_draws = n_draws; a.pos = nullptr; a.out = d_out; a.counts_out = nullptr; a.llr = nullptr; a.llr_stride = 0; a.use_gather4 = 0;
    const int rcap = (int)max_rows;
    const int bytes = SmallWinSmem::bytes(rcap);
    SmallWinArgs w;
    w.base = a; w.woff = d_woff; w.doff = d_doff; w.n_windows = n_windows; w.rcap = rcap;
    w.masks = masks ? masks : c->read_masks[0]; w.popc = popc ? popc : c->read_popc[0]; w.local = local ? 1 : 0;
    LDPGTA_CUDA(cudaFuncSetAttribute(k_small_win, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    k_small_win<<<(unsigned)std::min<int64_t>(n_windows, c->sm_count), SmallWinSmem::NWARPS * 32, bytes, c->stream>>>(w);
    c->launches++;
    LDPGTA_CUDA(cudaGetLastError());
    return 0;
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
    if (kind != LDPGTA_HOMOGENEOUS) {
        const int rc = kind == LDPGTA_RECENT_ADMIXTURE ? launch_small_both<LDPGTA_RECENT_ADMIXTURE>(c, a) : launch_small_both<LDPGTA_DISTANT_ADMIXTURE>(c, a);
        if (rc <= 0) return rc;
    }
    return launch_small_n<4>(c, a, kind);
}

}  // namespace ldpgta
