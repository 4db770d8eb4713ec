// bootstrap_api.cu -- the bootstrap loop and statistics() as ONE device-resident pipeline (include/ldpgta.h):
//   ldpgta_set_windows      windows, draw plan and (embryo, chromosome) segments made resident next to the read masks
//   ldpgta_bootstrap_stats  pick_reads on the device -> likelihoods -> LLRs -> window mean/variance -> segment sums -> finish
//   ldpgta_replay_stats     the same for draws given by the caller (the reference's random.sample stream, replayed)
//   ldpgta_window_stats     the reductions alone, for likelihoods that are on the host
// Likelihoods never leave HBM between the draw kernels and the reductions (ANEUPLOIDY_TEST.py:277-325 consumes them in place).
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "k_stats.cuh"

namespace ldpgta {

struct SizeClass {
    int n = 0;                     // reads per draw
    int64_t n_windows = 0, n_draws = 0;
    std::vector<int32_t> win;      // sample window of class window j
    int64_t *d_doff = nullptr;     // [n_windows + 1] class-local draw offsets
    int32_t *d_win = nullptr;      // device copy of win (nullptr: identity, the plan has one class)
    int32_t *d_idx = nullptr;      // [n_draws][n] draws of the class
    int64_t *d_pos = nullptr;      // [n_draws] place of each draw in the sample-wide order (nullptr: identity)
    int32_t *d_wod = nullptr;      // [n_draws] class window of each draw (the sampler's window lookup)
};

struct WindowPlan {
    int64_t n_windows = 0, n_draws = 0, n_wreads = 0, nnz = 0;
    int n_segments = 0, ncp = 0;
    int64_t max_draws = 0;         // most draws of a window
    int64_t max_rows = 0;          // most reads of a window (0 without window read ranges)
    bool consecutive = false;      // windows are runs of consecutive rows of the read-mask table
    bool has_windows = false;
    void *d_static = nullptr, *d_work = nullptr;
    int64_t *d_woff = nullptr, *d_doff = nullptr, *d_segoff = nullptr;
    int32_t *d_wreads = nullptr, *d_segcols = nullptr, *d_segfirst = nullptr;
    std::vector<SizeClass> classes;
    std::vector<int64_t> h_doff;   // draw offsets per window
    std::vector<int32_t> h_n;      // reads per draw per window
    double *d_lik = nullptr, *d_llr = nullptr, *d_mv = nullptr, *d_wsum = nullptr, *d_split = nullptr, *d_partials = nullptr, *d_stats = nullptr;
    uint64_t *d_seed = nullptr;
    unsigned int *d_arrived = nullptr;
    // chunks of up to CH_WIN consecutive windows of one segment (k_window_stats_chunks / _block); 0 when a window has more than
    // CHUNK_MAX_DRAWS draws
    int64_t n_chunks = 0;
    bool warp_chunks = false;
    std::vector<int64_t> h_chunk_win;   // host copy of the chunk table (window ranges of the download slices)
    int64_t *d_chunk_win = nullptr, *d_cseg_off = nullptr;
    double *d_colpart = nullptr;
    // windows given as read lists, many 4-read draws per window: the read masks in the order of the lists (every window a run
    // of consecutive rows of the copy) and the sampler's positions inside the windows, for k_small_win (ensure_wmasks)
    uint64_t *d_wmasks = nullptr;
    uint32_t *d_wpopc = nullptr;
    int32_t *d_pick = nullptr;     // [n_draws][4]
    uint64_t wm_epoch = 0;         // ldpgta_ctx::masks_epoch the copy was made at
    bool wm_refused = false;       // the copy did not fit next to the table: not tried again
    int64_t n_reads_planned = 0;   // rows of the read-mask table when the windows were checked against it
    bool picks_fresh = false;      // d_pick holds the positions of the current draws
    bool rows_fresh = true;        // classes[].d_idx holds the rows of the current draws (false: the sampler wrote positions only)
};

void free_window_plan(ldpgta_ctx *c)
{
    if (!c->plan) return;
    if (c->llr_ptr == c->plan->d_llr) { c->llr_ptr = nullptr; c->llr_off_ptr = nullptr; c->llr_draws = c->llr_windows = -1; c->llr_lazy_lik = nullptr; }
    if (c->plan->d_static) cudaFree(c->plan->d_static);
    if (c->plan->d_work) cudaFree(c->plan->d_work);
    if (c->plan->d_wmasks) cudaFree(c->plan->d_wmasks);
    if (c->plan->d_wpopc) cudaFree(c->plan->d_wpopc);
    if (c->plan->d_pick) cudaFree(c->plan->d_pick);
    delete c->plan;
    c->plan = nullptr;
}

template <int K>
static void launch_sample_k(ldpgta_ctx *c, const SampleArgs &a)
{
    const unsigned blocks = (unsigned)std::min<int64_t>((a.n_draws + 127) / 128, (int64_t)c->sm_count * 16);
    k_sample_draws<K><<<blocks, 128, 0, c->stream>>>(a);
    c->launches++;
}

int launch_sample(ldpgta_ctx *c, int n, const SampleArgs &a)
{
    if (a.n_draws == 0) return 0;
    switch (n) {
#define LDPGTA_K(K_) case K_: launch_sample_k<K_>(c, a); break;
        LDPGTA_K(2) LDPGTA_K(3) LDPGTA_K(4) LDPGTA_K(5) LDPGTA_K(6) LDPGTA_K(7) LDPGTA_K(8) LDPGTA_K(9) LDPGTA_K(10)
        LDPGTA_K(11) LDPGTA_K(12) LDPGTA_K(13) LDPGTA_K(14) LDPGTA_K(15) LDPGTA_K(16) LDPGTA_K(17) LDPGTA_K(18)
#undef LDPGTA_K
        default: return fail(LDPGTA_E_UNSUPPORTED, "draws of %d reads are outside the supported range 2..%d", n, MAXN);
    }
    LDPGTA_CUDA(cudaGetLastError());
    return 0;
}

static int time_mark(ldpgta_ctx *c, int what = 1)
{
    if (c->time_kernels != what) return 0;
    // marks come in (start, stop) pairs: bracket number t_marks / 2 is recorded when it is a multiple of the stride
    const uint64_t bracket = c->t_marks++ / 2;
    if (c->t_stride > 1 && bracket % (uint64_t)c->t_stride) return 0;
    if (c->t_used == c->t_ev.size()) {
        cudaEvent_t e;
        LDPGTA_CUDA(cudaEventCreate(&e));
        c->t_ev.push_back(e);
    }
    LDPGTA_CUDA(cudaEventRecord(c->t_ev[c->t_used++], c->stream));
    return 0;
}

// LLRs (when not already written by the draw kernels), window mean / variance, segment partial sums and (optionally) the
// finished segment statistics, all on c->stream
static int stats_on_device(ldpgta_ctx *c, const double *d_lik, const int64_t *d_doff, int64_t n_windows, int64_t n_draws,
                           const int64_t *d_segoff, const int32_t *d_segcols, const int32_t *d_segfirst, int n_segments, int ncp,
                           double *d_llr, double *d_mv, double *d_wsum, double *d_split, unsigned int *d_arrived, double *d_partials, double *d_stats,
                           double *mv_host = nullptr, int64_t max_draws_per_window = 0, const WindowPlan *chunks = nullptr, bool want_llr = true, bool slice_mv = true)
{
    if (chunks && chunks->n_chunks > 0 && n_windows > 0) {
        // no window has more than CHUNK_MAX_DRAWS draws: the column sums are formed next to the window statistics, chunk by chunk
        ChunkStatsArgs a;
        a.st.llr = d_llr; a.st.lik = d_lik; a.st.doff = d_doff; a.st.n_windows = n_windows; a.st.n_draws = n_draws; a.st.mean_var = d_mv; a.st.wsum = d_wsum;
        a.chunk_win = chunks->d_chunk_win; a.n_chunks = chunks->n_chunks; a.colpart = chunks->d_colpart; a.ncp = ncp; a.write_llr = want_llr ? 1 : 0;
        // the window statistics (80 bytes per window) go down while later windows are still being reduced: the chunks are cut
        // into slices, slice i's rows of mean_var[5][n_windows][2] travel under the kernel of slice i + 1
        static const int max_slices = [] { const char *e = getenv("LDPGTA_STATS_SLICES"); return e && atoi(e) > 0 ? atoi(e) : 2; }();      // measured on configs[1]: 1 / 2 / 4 / 8 slices -> 0.504 / 0.492 / 0.493 / 0.513 ms per call
        // (pageable destinations make every copy a synchronous staged one: no slices there)
        const int n_slices = mv_host && slice_mv && is_page_locked(mv_host) ? (int)std::max<int64_t>(1, std::min<int64_t>(max_slices, chunks->n_chunks / 512)) : 1;
        if (mv_host) if (int rc = ensure_pipeline(c, n_slices + 1)) return rc;
        const int64_t per_slice = (chunks->n_chunks + n_slices - 1) / n_slices;
        int si = 0;
        if (int rc = time_mark(c, 3)) return rc;
        for (int64_t c0 = 0; c0 < chunks->n_chunks; c0 += per_slice, ++si) {
            const int64_t c1 = std::min(c0 + per_slice, chunks->n_chunks);
            a.chunk_win = chunks->d_chunk_win + c0; a.n_chunks = c1 - c0; a.colpart = chunks->d_colpart + c0 * 5 * (ncp + 2);
            const unsigned blocks = (unsigned)std::min<int64_t>(a.n_chunks, (int64_t)c->sm_count * 64);
            if (max_draws_per_window <= 32) {
                if (chunks->warp_chunks) k_window_stats_warpchunks<<<(unsigned)std::min<int64_t>((a.n_chunks + 7) / 8, (int64_t)c->sm_count * 32), 256, 0, c->stream>>>(a);
                else k_window_stats_chunks<<<blocks, CH_WIN * 32, 0, c->stream>>>(a);       // persistent blocks with the next chunk's loads in flight: measured no faster
            }
            // threads x draws per thread by the longest window; the register budget (LLRs + column sums of DPT draws per pair)
            // decides how many blocks share an SM while one of them sits in a barrier
            else if (max_draws_per_window <= 256) k_window_stats_block<1, 256, 4><<<blocks, 256, 0, c->stream>>>(a);
            else if (max_draws_per_window <= 512) k_window_stats_block<2, 256, 4><<<blocks, 256, 0, c->stream>>>(a);
            else if (max_draws_per_window <= 1024) k_window_stats_block<4, 256, 2><<<blocks, 256, 0, c->stream>>>(a);      // measured: <2, 512, 2> is 20 % slower
            else k_window_stats_block<4, 512, 1><<<blocks, 512, 0, c->stream>>>(a);
            c->launches++;
            if (mv_host) {
                const int64_t w_lo = chunks->h_chunk_win[c0], w_hi = chunks->h_chunk_win[c1];
                LDPGTA_CUDA(cudaEventRecord(c->ev_in[si], c->stream));
                LDPGTA_CUDA(cudaStreamWaitEvent(c->s_out, c->ev_in[si], 0));
                LDPGTA_CUDA(cudaMemcpy2DAsync(mv_host + 2 * w_lo, (size_t)n_windows * 16, d_mv + 2 * w_lo, (size_t)n_windows * 16, (size_t)(w_hi - w_lo) * 16, 5,
                                              cudaMemcpyDeviceToHost, c->s_out));
            }
        }
        if (int rc = time_mark(c, 3)) return rc;
        SegChunkArgs sa;
        sa.colpart = chunks->d_colpart; sa.cseg_off = chunks->d_cseg_off; sa.seg_off = d_segoff; sa.seg_cols = d_segcols; sa.seg_first = d_segfirst;
        sa.ncp = ncp; sa.partials = d_partials; sa.stats = d_stats; sa.arrived = d_arrived;
        k_segment_from_chunks<<<dim3((unsigned)n_segments, 5, (unsigned)((ncp + 2 + SC_COLS - 1) / SC_COLS)), SC_ROWS * SC_COLS, 0, c->stream>>>(sa);
        c->launches++;
        LDPGTA_CUDA(cudaGetLastError());
        return 0;
    }
    // LLRs: windows of at most 32 draws form them inside k_window_stats (one draw per lane); longer windows would make
    // that one warp walk the window five times, so the LLRs come from the fully parallel k_draw_llr first
    if (d_lik && n_draws > 0 && max_draws_per_window > 32) {
        const unsigned blocks = (unsigned)std::min<int64_t>((n_draws + 255) / 256, (int64_t)c->sm_count * 8);
        k_draw_llr<<<blocks, 256, 0, c->stream>>>(d_lik, n_draws, d_llr);
        c->launches++;
        d_lik = nullptr;
    }
    if (n_windows > 0) {
        StatsArgs a;
        a.llr = d_llr; a.lik = d_lik; a.doff = d_doff; a.n_windows = n_windows; a.n_draws = n_draws; a.mean_var = d_mv; a.wsum = d_wsum;      // d_lik == nullptr: the draw kernels have written the LLRs already
        const unsigned blocks = (unsigned)std::min<int64_t>((n_windows + 7) / 8, (int64_t)c->sm_count * 32);
        k_window_stats<<<blocks, 256, 0, c->stream>>>(a);
        c->launches++;
        if (mv_host) {
            // the window statistics (80 bytes per window) start their way down while the segment sums are formed
            if (int rc = ensure_pipeline(c, 2)) return rc;
            LDPGTA_CUDA(cudaEventRecord(c->ev_in[0], c->stream));
            LDPGTA_CUDA(cudaStreamWaitEvent(c->s_out, c->ev_in[0], 0));
            LDPGTA_CUDA(cudaMemcpyAsync(mv_host, d_mv, (size_t)n_windows * 80, cudaMemcpyDeviceToHost, c->s_out));
        }
    }
    SegmentArgs sa;
    sa.llr = d_llr; sa.doff = d_doff; sa.wsum = d_wsum; sa.mean_var = d_mv; sa.seg_off = d_segoff; sa.seg_cols = d_segcols;
    sa.seg_first = d_segfirst; sa.n_windows = n_windows; sa.n_draws = n_draws; sa.ncp = ncp; sa.partials = d_partials; sa.stats = d_stats; sa.split = d_split; sa.arrived = d_arrived;
    k_segment_partials<<<dim3((unsigned)n_segments, 5, SEG_SPLIT), SEG_ROWS * SEG_COLS, 0, c->stream>>>(sa);
    c->launches++;
    LDPGTA_CUDA(cudaGetLastError());
    return 0;
}

constexpr int64_t CHUNK_MAX_DRAWS = 2048;           // k_window_stats_block<4, 512>

static size_t pad256(size_t b) { return (b + 255) & ~(size_t)255; }

// draws of every class -> likelihoods in the sample-wide order (plan->d_lik), optionally streaming them to the host
// The draw kernels for 5 or more reads write the five LLRs of a draw next to its likelihoods (their per-draw epilogue is
// negligible); for 2..4 reads (k_small: 0.4 ns per draw) the LLRs are a separate pass over the likelihoods (k_draw_llr),
// measured equal in total and leaving the dominant kernel to its AND / POPC work.
static bool llr_fused(const WindowPlan *p)
{
    for (const SizeClass &k : p->classes)
        if (k.n <= 4) return false;
    return true;
}

// idx_up: page-locked host draws of a one-class plan that still have to go up (ldpgta_replay_stats); lim = reads in the table
// sampled: the draws are the device sampler's (every draw inside its window), not the caller's
static int evaluate_classes_inner(ldpgta_ctx *c, WindowPlan *p, int kind, const double *proportions, double *lik_out_pinned,
                                  const int32_t *idx_up, uint32_t lim, bool sampled);

static int evaluate_classes(ldpgta_ctx *c, WindowPlan *p, int kind, const double *proportions, double *lik_out_pinned,
                            const int32_t *idx_up = nullptr, uint32_t lim = 0, bool sampled = false)
{
    if (int rc = time_mark(c)) return rc;
    if (int rc = evaluate_classes_inner(c, p, kind, proportions, lik_out_pinned, idx_up, lim, sampled)) return rc;
    return time_mark(c);
}

// mean draws per window from which k_small_win replaces k_small (0 = never)
static int64_t win_min_draws()
{
    static const int64_t v = [] { const char *e = getenv("LDPGTA_WIN_MIN_DRAWS"); return e ? (int64_t)atoll(e) : (int64_t)256; }();
    return v;
}

// one group, 4 reads per draw, enough draws per window for the window-resident kernel
static bool plan_wants_small_windows(const ldpgta_ctx *c, const WindowPlan *p, int kind)
{
    return kind == LDPGTA_HOMOGENEOUS && p->classes.size() == 1 && p->classes[0].n == 4 && win_min_draws() > 0 && p->n_windows > 0 &&
           p->n_draws >= win_min_draws() * p->n_windows && small_windows_ok(c, p->max_rows);
}

// Windows given as read lists (packed.py: reads that fail min_score leave gaps): the plan keeps the read masks in the order of
// the lists, so that every window is a run of consecutive rows of the copy, and the sampler also writes each draw's positions
// inside its window.  *ok = false when the plan is not eligible or the copy does not fit (the caller takes k_small).
static int ensure_wmasks(ldpgta_ctx *c, WindowPlan *p, int kind, bool *ok)
{
    *ok = false;
    if (!p->n_wreads || p->wm_refused || !plan_wants_small_windows(c, p, kind)) return 0;
    if (p->d_wmasks && p->wm_epoch == c->masks_epoch) { *ok = true; return 0; }
    if (!p->d_wmasks) {
        const size_t need = (size_t)p->n_wreads * 644 + (size_t)p->n_draws * 16;
        size_t free_b = 0, total_b = 0;
        LDPGTA_CUDA(cudaMemGetInfo(&free_b, &total_b));
        if (free_b < need + ((size_t)1 << 30)) { p->wm_refused = true; return 0; }
        if (cudaMalloc((void **)&p->d_wmasks, (size_t)p->n_wreads * 640) != cudaSuccess || cudaMalloc((void **)&p->d_wpopc, (size_t)p->n_wreads * 4) != cudaSuccess ||
            cudaMalloc((void **)&p->d_pick, (size_t)p->n_draws * 16) != cudaSuccess) {
            (void)cudaGetLastError();                                  // an optimisation that does not fit is not an error
            if (p->d_wmasks) cudaFree(p->d_wmasks);
            if (p->d_wpopc) cudaFree(p->d_wpopc);
            if (p->d_pick) cudaFree(p->d_pick);
            p->d_wmasks = nullptr; p->d_wpopc = nullptr; p->d_pick = nullptr;
            p->wm_refused = true;
            return 0;
        }
    }
    const int64_t total = p->n_wreads * 40;
    k_gather_listed_rows<<<(unsigned)std::min<int64_t>((total + 255) / 256, (int64_t)c->sm_count * 16), 256, 0, c->stream>>>(
        c->read_masks[0], c->read_popc[0], p->d_wreads, p->n_wreads, 80, p->d_wmasks, p->d_wpopc);
    c->launches++;
    LDPGTA_CUDA(cudaGetLastError());
    p->wm_epoch = c->masks_epoch;
    p->picks_fresh = false;
    *ok = true;
    return 0;
}

// the sampler of every class; with window-ordered masks it also leaves the positions inside the windows (k_small_win), and
// then writes the rows of the table only when somebody reads them (need_rows: the caller's `draws`, or k_small under a
// download of the likelihoods) -- 16 bytes per draw and four dependent loads less
static int sample_classes(ldpgta_ctx *c, WindowPlan *p, int kind, uint64_t seed, bool need_rows)
{
    bool picks = false;
    if (int rc = ensure_wmasks(c, p, kind, &picks)) return rc;
    p->rows_fresh = !picks || need_rows;
    for (const SizeClass &k : p->classes) {
        SampleArgs sa;
        memset(&sa, 0, sizeof sa);
        sa.draw_off = k.d_doff; sa.win_of_draw = k.d_wod; sa.win_off = p->d_woff; sa.win_reads = p->d_wreads; sa.cls_win = k.d_win;
        sa.n_windows = k.n_windows; sa.n_draws = k.n_draws; sa.seed = seed + (uint64_t)k.n; sa.idx_out = k.d_idx;
        sa.pick_out = picks ? p->d_pick : nullptr;
        if (!p->rows_fresh) sa.idx_out = nullptr;
        if (int rc = launch_sample(c, k.n, sa)) return rc;
    }
    p->picks_fresh = picks;
    return 0;
}

static int evaluate_classes_inner(ldpgta_ctx *c, WindowPlan *p, int kind, const double *proportions, double *lik_out_pinned,
                                  const int32_t *idx_up, uint32_t lim, bool sampled)
{
    if (idx_up && !lik_out_pinned) {
        // host-chosen draws of one size, statistics only: the draws go up in chunks on the upload stream, chunk i + 1 under the
        // kernel of chunk i (one copy followed by one kernel: 0.78 ms per configs[1] step, of which 0.27 ms is the 14.7 MB copy)
        const SizeClass &k = p->classes[0];
        static const int64_t max_chunks = [] { const char *e = getenv("LDPGTA_REPLAY_CHUNKS"); return e && atoi(e) > 0 ? (int64_t)atoi(e) : (int64_t)4; }();
        const int64_t n_chunks = std::max<int64_t>(1, std::min<int64_t>(max_chunks, p->n_draws / 65536));
        const int64_t per = ((p->n_draws + n_chunks - 1) / n_chunks + 63) & ~(int64_t)63;
        if (int rc = ensure_pipeline(c, (int)n_chunks + 1)) return rc;
        LDPGTA_CUDA(cudaEventRecord(c->ev_k[0], c->stream));                    // the upload stream starts after what is queued here (flag reset)
        LDPGTA_CUDA(cudaStreamWaitEvent(c->s_in, c->ev_k[0], 0));
        int ci = 0;
        for (int64_t d0 = 0; d0 < p->n_draws; d0 += per, ++ci) {
            const int64_t nd = std::min(per, p->n_draws - d0), cnt = nd * k.n;
            LDPGTA_CUDA(cudaMemcpyAsync(k.d_idx + d0 * k.n, idx_up + d0 * k.n, (size_t)cnt * 4, cudaMemcpyHostToDevice, c->s_in));
            LDPGTA_CUDA(cudaEventRecord(c->ev_in[ci], c->s_in));
            LDPGTA_CUDA(cudaStreamWaitEvent(c->stream, c->ev_in[ci], 0));
            k_sanitize_idx<<<(unsigned)std::min<int64_t>((cnt / 4 + 255) / 256 + 1, (int64_t)c->sm_count * 4), 256, 0, c->stream>>>(k.d_idx + d0 * k.n, cnt, lim, c->d_flag);
            c->launches++;
            if (int rc = run_uniform(c, kind, k.n, k.d_idx + d0 * k.n, nd, nullptr, proportions, p->d_lik + 4 * d0, nullptr, false, nullptr,
                                     llr_fused(p) ? p->d_llr + d0 : nullptr, p->n_draws))
                return rc;
        }
        return 0;
    }
    if (p->classes.size() == 1 && lik_out_pinned && p->n_draws >= 65536 && p->rows_fresh) {
        // one draw size (the usual plan): chunks, the download of chunk i under the kernel of chunk i + 1
        const SizeClass &k = p->classes[0];
        static const int64_t max_chunks = [] { const char *e = getenv("LDPGTA_FULL_CHUNKS"); return e && atoi(e) > 0 ? (int64_t)atoi(e) : (int64_t)8; }();      // measured: 4 / 8 / 16 / 32 chunks -> 0.72 / 0.70 / 0.74 / 0.85 ms
        const int64_t n_chunks = std::max<int64_t>(1, std::min<int64_t>(max_chunks, p->n_draws / 32768));
        const int64_t per = ((p->n_draws + n_chunks - 1) / n_chunks + 63) & ~(int64_t)63;
        if (int rc = ensure_pipeline(c, (int)n_chunks + 1)) return rc;
        int ci = 0;
        for (int64_t d0 = 0; d0 < p->n_draws; d0 += per, ++ci) {
            const int64_t nd = std::min(per, p->n_draws - d0);
            if (int rc = run_uniform(c, kind, k.n, k.d_idx + d0 * k.n, nd, nullptr, proportions, p->d_lik + 4 * d0, nullptr, false, nullptr,
                                     llr_fused(p) ? p->d_llr + d0 : nullptr, p->n_draws))
                return rc;
            LDPGTA_CUDA(cudaEventRecord(c->ev_k[ci], c->stream));
            LDPGTA_CUDA(cudaStreamWaitEvent(c->s_out, c->ev_k[ci], 0));
            LDPGTA_CUDA(cudaMemcpyAsync(lik_out_pinned + 4 * d0, p->d_lik + 4 * d0, (size_t)nd * 32, cudaMemcpyDeviceToHost, c->s_out));
        }
        return 0;
    }
    // many draws per window (e.g. 1000 subsamples), 4 reads each, drawn on the device: the window-resident kernel, on the
    // read-mask table itself (windows = runs of consecutive rows) or on the plan's window-ordered copy (windows = read lists).
    // Draws of the caller (ldpgta_replay_stats) may name any row of the table and take k_small.
    const bool by_lists = p->picks_fresh && p->d_wmasks && p->wm_epoch == c->masks_epoch;
    if (sampled && (p->consecutive || by_lists) && plan_wants_small_windows(c, p, kind)) {
        const int rc = p->consecutive
                           ? launch_small_windows(c, p->classes[0].d_idx, p->n_draws, p->d_lik, p->d_woff, p->d_doff, p->n_windows, p->max_rows)
                           : launch_small_windows(c, p->d_pick, p->n_draws, p->d_lik, p->d_woff, p->d_doff, p->n_windows, p->max_rows,
                                                  p->d_wmasks, p->d_wpopc, true);
        if (rc < 0) return rc;
        if (rc == 0) {
            if (lik_out_pinned) {
                if (int rc2 = ensure_pipeline(c, 2)) return rc2;
                LDPGTA_CUDA(cudaEventRecord(c->ev_k[0], c->stream));
                LDPGTA_CUDA(cudaStreamWaitEvent(c->s_out, c->ev_k[0], 0));
                LDPGTA_CUDA(cudaMemcpyAsync(lik_out_pinned, p->d_lik, (size_t)p->n_draws * 32, cudaMemcpyDeviceToHost, c->s_out));
            }
            return 0;
        }
    }
    if (!p->rows_fresh) return fail(LDPGTA_E_STATE, "the sampler left positions only, but the draw kernel needs the rows of the draws");
    for (const SizeClass &k : p->classes)
        if (int rc = run_uniform(c, kind, k.n, k.d_idx, k.n_draws, k.d_pos, proportions, p->d_lik, nullptr, false, nullptr,
                                 llr_fused(p) ? p->d_llr : nullptr, p->n_draws))
            return rc;
    if (lik_out_pinned) {
        if (int rc = ensure_pipeline(c, 2)) return rc;
        LDPGTA_CUDA(cudaEventRecord(c->ev_k[0], c->stream));
        LDPGTA_CUDA(cudaStreamWaitEvent(c->s_out, c->ev_k[0], 0));
        LDPGTA_CUDA(cudaMemcpyAsync(lik_out_pinned, p->d_lik, (size_t)p->n_draws * 32, cudaMemcpyDeviceToHost, c->s_out));
    }
    return 0;
}

static int check_plan(const ldpgta_ctx *c, int kind, const double *proportions)
{
    if (!c) return fail(LDPGTA_E_INVALID, "null context");
    if (!c->plan) return fail(LDPGTA_E_STATE, "no windows are resident: call ldpgta_set_windows first");
    if (int rc = check_kind(c, kind, proportions)) return rc;
    for (const SizeClass &k : c->plan->classes)
        if (int rc = check_uniform_size(kind, k.n)) return rc;
    // the windows were checked against the table that was resident then; a smaller table would be read out of bounds
    if (c->plan->has_windows && c->n_reads < c->plan->n_reads_planned)
        return fail(LDPGTA_E_STATE, "the read-mask table shrank from %lld to %lld rows under the resident windows: call ldpgta_set_windows again",
                    (long long)c->plan->n_reads_planned, (long long)c->n_reads);
    return 0;
}

// reductions of the resident likelihoods + downloads of what the caller asked for; waits for everything
static int reduce_and_download(ldpgta_ctx *c, WindowPlan *p, double *lik_out, double *lik_pinned, double *llr_out,
                               double *window_mean_var, double *segment_partials, double *segment_stats)
{
    double *d_llr = p->d_llr;
    c->llr_draws = c->llr_windows = -1;
    c->llr_lazy_lik = nullptr;
    // the LLR matrix is written when the caller wants it; otherwise ldpgta_bin_stats forms it from the resident likelihoods
    // the first time it is asked for (k_draw_llr)
    const bool lazy = !llr_fused(p) && p->n_chunks > 0 && !llr_out;
    if (int rc = stats_on_device(c, llr_fused(p) ? nullptr : p->d_lik, p->d_doff, p->n_windows, p->n_draws, p->d_segoff, p->d_segcols, p->d_segfirst,
                                 p->n_segments, p->ncp, d_llr, p->d_mv, p->d_wsum, p->d_split, p->d_arrived, p->d_partials, p->d_stats, window_mean_var, p->max_draws,
                                 p, !lazy, !lik_out))                 // the likelihoods' own download keeps the copy engine busy: no slices
        return rc;
    const size_t cp_b = (size_t)p->n_segments * 5 * (p->ncp + 3) * 8;
    if (segment_partials) LDPGTA_CUDA(cudaMemcpyAsync(segment_partials, p->d_partials, cp_b, cudaMemcpyDeviceToHost, c->stream));
    if (segment_stats) LDPGTA_CUDA(cudaMemcpyAsync(segment_stats, p->d_stats, (size_t)p->n_segments * 120, cudaMemcpyDeviceToHost, c->stream));
    if (llr_out && p->n_draws) LDPGTA_CUDA(cudaMemcpyAsync(llr_out, d_llr, (size_t)p->n_draws * 40, cudaMemcpyDeviceToHost, c->stream));
    LDPGTA_CUDA(cudaStreamSynchronize(c->stream));
    if (lik_out || window_mean_var) LDPGTA_CUDA(cudaStreamSynchronize(c->s_out));
    if (lik_out && lik_pinned != lik_out) memcpy(lik_out, lik_pinned, (size_t)p->n_draws * 32);
    c->llr_ptr = p->d_llr; c->llr_off_ptr = p->d_doff;
    c->llr_draws = p->n_draws; c->llr_windows = p->n_windows;
    c->llr_lazy_lik = lazy ? p->d_lik : nullptr;
    return 0;
}

}  // namespace ldpgta

using namespace ldpgta;

extern "C" {

int ldpgta_set_windows(ldpgta_ctx *c, const int64_t *window_offsets, const int32_t *window_reads, int64_t n_windows,
                       const int64_t *draw_offsets, const int32_t *reads_per_draw, const int64_t *segment_offsets,
                       int n_segments, const int32_t *segment_cols, const int32_t *segment_first)
{
    if (!c || !draw_offsets || !reads_per_draw || n_windows < 0) return fail(LDPGTA_E_INVALID, "bad arguments");
    if (window_reads && !window_offsets) return fail(LDPGTA_E_INVALID, "window_reads without window_offsets");
    LDPGTA_CUDA(cudaSetDevice(c->device));
    if (c->n_reads <= 0 && window_offsets) return fail(LDPGTA_E_STATE, "read masks are missing (upload or build them first)");
    const int64_t one_seg[2] = {0, n_windows};
    if (!segment_offsets) { segment_offsets = one_seg; n_segments = 1; }
    if (n_segments < 1) return fail(LDPGTA_E_INVALID, "at least one segment is needed");
    // consecutive-row windows address the read-mask table directly and may start anywhere in it (a shard of a sample)
    if (draw_offsets[0] != 0 || (window_reads && window_offsets[0] != 0) || (window_offsets && window_offsets[0] < 0) || segment_offsets[0] != 0)
        return fail(LDPGTA_E_INVALID, "offsets must start at 0");
    if (segment_offsets[n_segments] != n_windows) return fail(LDPGTA_E_INVALID, "segment_offsets must end at n_windows");
    for (int s = 0; s < n_segments; ++s)
        if (segment_offsets[s + 1] < segment_offsets[s]) return fail(LDPGTA_E_INVALID, "segment_offsets must not decrease (segment %d)", s);
    int64_t per_n[MAXN + 1] = {}, win_n[MAXN + 1] = {}, nnz = 0;
    int64_t max_draws = 0, max_rows = 0;
    for (int64_t w = 0; w < n_windows; ++w) {
        const int64_t nd = draw_offsets[w + 1] - draw_offsets[w];
        const int n = reads_per_draw[w];
        if (nd < 1) return fail(LDPGTA_E_INVALID, "window %lld has no draws", (long long)w);
        if (n < 2 || n > MAXN) return fail(LDPGTA_E_UNSUPPORTED, "window %lld: draws of %d reads, supported range is 2..%d", (long long)w, n, MAXN);
        if (window_offsets) {
            const int64_t R = window_offsets[w + 1] - window_offsets[w];
            if (R < n) return fail(LDPGTA_E_INVALID, "window %lld has %lld reads, fewer than the %d of a draw", (long long)w, (long long)R, n);
            if (R > 0x7fffffff) return fail(LDPGTA_E_INVALID, "window %lld has too many reads", (long long)w);
            max_rows = std::max(max_rows, R);
        }
        per_n[n] += nd; win_n[n]++; nnz += nd * n;
        max_draws = std::max(max_draws, nd);
    }
    const int64_t n_draws = draw_offsets[n_windows];
    const int64_t n_wreads = window_offsets ? window_offsets[n_windows] : 0;
    if (window_offsets && !window_reads && n_wreads > c->n_reads)
        return fail(LDPGTA_E_INVALID, "window_offsets address read rows up to %lld, the context holds %lld", (long long)n_wreads, (long long)c->n_reads);

    // segments: columns every window has (zip truncation, ANEUPLOIDY_TEST.py:73) and the first window's draws (:71)
    std::vector<int32_t> segcols(n_segments), segfirst(n_segments);
    int ncp = 0;
    for (int s = 0; s < n_segments; ++s) {
        const int64_t lo = segment_offsets[s], hi = segment_offsets[s + 1];
        int64_t mn = hi > lo ? INT64_MAX : 0;
        for (int64_t w = lo; w < hi; ++w) mn = std::min(mn, draw_offsets[w + 1] - draw_offsets[w]);
        segcols[s] = segment_cols ? segment_cols[s] : (int32_t)std::min<int64_t>(mn, 0x7fffffff);
        segfirst[s] = segment_first ? segment_first[s] : (hi > lo ? (int32_t)(draw_offsets[lo + 1] - draw_offsets[lo]) : 0);
        if (segcols[s] < 0 || (hi > lo && segcols[s] > mn))
            return fail(LDPGTA_E_INVALID, "segment %d: %d columns, but one of its windows has only %lld draws", s, segcols[s], (long long)mn);
        ncp = std::max(ncp, (int)segcols[s]);
    }
    LDPGTA_CUDA(cudaStreamSynchronize(c->stream));
    free_window_plan(c);
    WindowPlan *p = new WindowPlan();
    c->plan = p;
    p->n_windows = n_windows; p->n_draws = n_draws; p->n_wreads = window_reads ? n_wreads : 0; p->nnz = nnz;
    p->n_segments = n_segments; p->ncp = ncp; p->max_draws = max_draws;
    p->has_windows = window_offsets != nullptr;
    p->n_reads_planned = window_offsets ? c->n_reads : 0;
    p->consecutive = window_offsets != nullptr && window_reads == nullptr;
    p->max_rows = max_rows;
    p->h_doff.assign(draw_offsets, draw_offsets + n_windows + 1);
    p->h_n.assign(reads_per_draw, reads_per_draw + n_windows);
    int n_classes = 0;
    for (int n = 2; n <= MAXN; ++n) n_classes += per_n[n] > 0;
    const bool single = n_classes <= 1;

    // host images of the static arrays, one allocation and one copy
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off += pad256(bytes); return o; };
    const size_t o_woff = take(window_offsets ? (size_t)(n_windows + 1) * 8 : 0), o_doff = take((size_t)(n_windows + 1) * 8);
    const size_t o_seg = take((size_t)(n_segments + 1) * 8), o_cols = take((size_t)n_segments * 4), o_first = take((size_t)n_segments * 4);
    const size_t o_wreads = take((size_t)p->n_wreads * 4), o_seed = take(8);
    // chunk table of the fused window / column reductions
    static const bool no_chunks = [] { const char *e = getenv("LDPGTA_NO_CHUNK_STATS"); return e && atoi(e) != 0; }();
    std::vector<int64_t> chunk_win, cseg_off;
    if (!no_chunks && max_draws <= CHUNK_MAX_DRAWS && n_windows > 0) {
        // windows per chunk: CH_WIN; with at most 32 draws per window a chunk may also be walked by one warp (LDPGTA_STATS_WARP=k)
        static const int warp_chunk = [] { const char *e = getenv("LDPGTA_STATS_WARP"); return e ? atoi(e) : 0; }();
        p->warp_chunks = max_draws <= 32 && warp_chunk > 0;
        const int per_chunk = p->warp_chunks ? warp_chunk : CH_WIN;
        cseg_off.push_back(0);
        for (int s = 0; s < n_segments; ++s) {
            for (int64_t w = segment_offsets[s]; w < segment_offsets[s + 1]; w += per_chunk) chunk_win.push_back(w);
            cseg_off.push_back((int64_t)chunk_win.size());
        }
        p->n_chunks = (int64_t)chunk_win.size();
        chunk_win.push_back(n_windows);
    }
    const size_t o_chunk = take(chunk_win.size() * 8), o_cseg = take(cseg_off.size() * 8);
    struct ClsOff { size_t doff, win, pos, wod; };
    std::vector<ClsOff> co;
    for (int n = 2; n <= MAXN; ++n) {
        if (!per_n[n]) continue;
        SizeClass k;
        k.n = n; k.n_windows = win_n[n]; k.n_draws = per_n[n];
        p->classes.push_back(k);
        ClsOff o;
        o.doff = take((size_t)(win_n[n] + 1) * 8);
        o.win = single ? 0 : take((size_t)win_n[n] * 4);
        o.pos = single ? 0 : take((size_t)per_n[n] * 8);
        o.wod = window_offsets ? take((size_t)per_n[n] * 4) : 0;
        co.push_back(o);
    }
    const size_t static_bytes = std::max<size_t>(off, 256);
    std::vector<unsigned char> img(static_bytes, 0);
    if (window_offsets) memcpy(img.data() + o_woff, window_offsets, (size_t)(n_windows + 1) * 8);
    memcpy(img.data() + o_doff, draw_offsets, (size_t)(n_windows + 1) * 8);
    memcpy(img.data() + o_seg, segment_offsets, (size_t)(n_segments + 1) * 8);
    memcpy(img.data() + o_cols, segcols.data(), (size_t)n_segments * 4);
    memcpy(img.data() + o_first, segfirst.data(), (size_t)n_segments * 4);
    if (p->n_wreads) memcpy(img.data() + o_wreads, window_reads, (size_t)p->n_wreads * 4);
    if (p->n_chunks) {
        memcpy(img.data() + o_chunk, chunk_win.data(), chunk_win.size() * 8);
        p->h_chunk_win = chunk_win;
        memcpy(img.data() + o_cseg, cseg_off.data(), cseg_off.size() * 8);
    }
    for (size_t ci = 0; ci < p->classes.size(); ++ci) {
        SizeClass &k = p->classes[ci];
        int64_t *cd = (int64_t *)(img.data() + co[ci].doff);
        int32_t *cwn = single ? nullptr : (int32_t *)(img.data() + co[ci].win);
        int64_t *cp = single ? nullptr : (int64_t *)(img.data() + co[ci].pos);
        int32_t *cwod = window_offsets ? (int32_t *)(img.data() + co[ci].wod) : nullptr;
        int64_t j = 0, d = 0;
        cd[0] = 0;
        k.win.reserve((size_t)k.n_windows);
        for (int64_t w = 0; w < n_windows; ++w) {
            if (reads_per_draw[w] != k.n) continue;
            const int64_t nd = draw_offsets[w + 1] - draw_offsets[w];
            k.win.push_back((int32_t)w);
            if (cwod) std::fill(cwod + d, cwod + d + nd, (int32_t)j);
            if (!single) {
                cwn[j] = (int32_t)w;
                for (int64_t t = 0; t < nd; ++t) cp[d + t] = draw_offsets[w] + t;
            }
            d += nd;
            cd[++j] = d;
        }
    }
    LDPGTA_CUDA(cudaMalloc(&p->d_static, static_bytes));
    LDPGTA_CUDA(cudaMemcpyAsync(p->d_static, img.data(), static_bytes, cudaMemcpyHostToDevice, c->stream));
    char *base = (char *)p->d_static;
    p->d_woff = window_offsets ? (int64_t *)(base + o_woff) : nullptr;
    p->d_doff = (int64_t *)(base + o_doff);
    p->d_segoff = (int64_t *)(base + o_seg);
    p->d_segcols = (int32_t *)(base + o_cols);
    p->d_segfirst = (int32_t *)(base + o_first);
    p->d_wreads = p->n_wreads ? (int32_t *)(base + o_wreads) : nullptr;
    p->d_seed = (uint64_t *)(base + o_seed);
    p->d_chunk_win = p->n_chunks ? (int64_t *)(base + o_chunk) : nullptr;
    p->d_cseg_off = p->n_chunks ? (int64_t *)(base + o_cseg) : nullptr;
    for (size_t ci = 0; ci < p->classes.size(); ++ci) {
        p->classes[ci].d_doff = (int64_t *)(base + co[ci].doff);
        p->classes[ci].d_win = single ? nullptr : (int32_t *)(base + co[ci].win);
        p->classes[ci].d_pos = single ? nullptr : (int64_t *)(base + co[ci].pos);
        p->classes[ci].d_wod = window_offsets ? (int32_t *)(base + co[ci].wod) : nullptr;
    }
    // window read lists are checked on the device (the copy is clamped, the call fails)
    if (p->n_wreads) {
        if (int rc = ensure_pipeline(c, 1)) return rc;
        LDPGTA_CUDA(cudaMemsetAsync(c->d_flag, 0, 4, c->stream));
        const uint32_t lim = (uint32_t)std::min<int64_t>(c->n_reads, 0x7fffffff);
        k_sanitize_idx<<<(unsigned)std::min<int64_t>((p->n_wreads / 4 + 255) / 256 + 1, (int64_t)c->sm_count * 4), 256, 0, c->stream>>>(
            p->d_wreads, p->n_wreads, lim, c->d_flag);
        c->launches++;
        LDPGTA_CUDA(cudaMemcpyAsync(c->h_flag, c->d_flag, 4, cudaMemcpyDeviceToHost, c->stream));
    }
    // work buffers: draws per class, likelihoods, window statistics, runs, segment partials and statistics
    off = 0;
    std::vector<size_t> o_idx;
    for (const SizeClass &k : p->classes) o_idx.push_back(take((size_t)k.n_draws * k.n * 4));
    const size_t o_lik = take((size_t)n_draws * 32), o_llr = take((size_t)n_draws * 40), o_mv = take((size_t)n_windows * 80);
    const size_t o_wsum = take((size_t)n_windows * 40), o_split = take((size_t)n_segments * 5 * SEG_SPLIT * (ncp + 3) * 8), o_arr = take((size_t)n_segments * 20);
    const size_t o_part = take((size_t)n_segments * 5 * (ncp + 3) * 8), o_stats = take((size_t)n_segments * 120);
    const size_t o_colpart = take((size_t)p->n_chunks * 5 * (ncp + 2) * 8);
    LDPGTA_CUDA(cudaMalloc(&p->d_work, std::max<size_t>(off, 256)));
    base = (char *)p->d_work;
    for (size_t ci = 0; ci < p->classes.size(); ++ci) p->classes[ci].d_idx = (int32_t *)(base + o_idx[ci]);
    p->d_lik = (double *)(base + o_lik);
    p->d_llr = (double *)(base + o_llr);
    p->d_mv = (double *)(base + o_mv);
    p->d_wsum = (double *)(base + o_wsum);
    p->d_split = (double *)(base + o_split);
    p->d_arrived = (unsigned int *)(base + o_arr);
    LDPGTA_CUDA(cudaMemsetAsync(p->d_arrived, 0, (size_t)n_segments * 20, c->stream));
    p->d_partials = (double *)(base + o_part);
    p->d_stats = (double *)(base + o_stats);
    p->d_colpart = p->n_chunks ? (double *)(base + o_colpart) : nullptr;
    LDPGTA_CUDA(cudaStreamSynchronize(c->stream));
    if (p->n_wreads && *c->h_flag) {
        free_window_plan(c);
        return fail_bad_index(c, window_reads, n_wreads);
    }
    return 0;
}

int ldpgta_bootstrap_stats(ldpgta_ctx *c, int kind, uint64_t seed, const double *proportions, double *lik_out, int32_t *idx_out,
                           double *llr_out, double *window_mean_var, double *segment_partials, double *segment_stats)
{
    if (int rc = check_plan(c, kind, proportions)) return rc;
    LDPGTA_CUDA(cudaSetDevice(c->device));
    WindowPlan *p = c->plan;
    if (!p->has_windows) return fail(LDPGTA_E_STATE, "the resident plan has no window read lists (replay-only): use ldpgta_replay_stats");
    double *lik_pinned = lik_out;
    if (lik_out && !is_page_locked(lik_out)) {
        if (int rc = ensure_pin(c, (size_t)p->n_draws * 32)) return rc;
        lik_pinned = (double *)c->h_pin;
    }
    if (int rc = sample_classes(c, p, kind, seed, idx_out != nullptr || lik_out != nullptr)) return rc;
    if (int rc = evaluate_classes(c, p, kind, proportions, lik_pinned, nullptr, 0, true)) return rc;
    if (idx_out) {
        // draws in the sample-wide order: window by window (classes interleave only when draw sizes differ)
        if (p->classes.size() == 1) {
            LDPGTA_CUDA(cudaMemcpyAsync(idx_out, p->classes[0].d_idx, (size_t)p->nnz * 4, cudaMemcpyDeviceToHost, c->stream));
        } else {
            std::vector<int32_t> tmp;
            std::vector<int64_t> start(p->n_windows + 1, 0);
            for (int64_t w = 0; w < p->n_windows; ++w) start[w + 1] = start[w] + (p->h_doff[w + 1] - p->h_doff[w]) * p->h_n[w];
            for (const SizeClass &k : p->classes) {
                tmp.resize((size_t)k.n_draws * k.n);
                LDPGTA_CUDA(cudaMemcpyAsync(tmp.data(), k.d_idx, tmp.size() * 4, cudaMemcpyDeviceToHost, c->stream));
                LDPGTA_CUDA(cudaStreamSynchronize(c->stream));
                int64_t at = 0;
                for (int32_t w : k.win) {
                    const int64_t cnt = (p->h_doff[w + 1] - p->h_doff[w]) * k.n;
                    memcpy(idx_out + start[w], tmp.data() + at, (size_t)cnt * 4);
                    at += cnt;
                }
            }
        }
    }
    return reduce_and_download(c, p, lik_out, lik_pinned, llr_out, window_mean_var, segment_partials, segment_stats);
}

int ldpgta_replay_stats(ldpgta_ctx *c, int kind, const int32_t *read_idx, const double *proportions, double *lik_out,
                        double *llr_out, double *window_mean_var, double *segment_partials, double *segment_stats)
{
    if (int rc = check_plan(c, kind, proportions)) return rc;
    if (!read_idx) return fail(LDPGTA_E_INVALID, "null read_idx");
    LDPGTA_CUDA(cudaSetDevice(c->device));
    WindowPlan *p = c->plan;
    double *lik_pinned = lik_out;
    size_t pin_lik = 0;
    if (lik_out && !is_page_locked(lik_out)) pin_lik = (size_t)p->n_draws * 32;
    const bool single = p->classes.size() == 1;
    const bool stage_idx = !single || !is_page_locked(read_idx);
    const size_t idx_b = pad256((size_t)p->nnz * 4);
    if (pin_lik || stage_idx)
        if (int rc = ensure_pin(c, (stage_idx ? idx_b : 0) + pin_lik)) return rc;
    if (pin_lik) lik_pinned = (double *)((char *)c->h_pin + (stage_idx ? idx_b : 0));
    if (int rc = ensure_pipeline(c, 1)) return rc;
    LDPGTA_CUDA(cudaMemsetAsync(c->d_flag, 0, 4, c->stream));
    const uint32_t lim = (uint32_t)std::min<int64_t>(c->n_reads, 0x7fffffff);
    const int32_t *src = read_idx;
    if (stage_idx) {
        int32_t *h = (int32_t *)c->h_pin;
        if (single) memcpy(h, read_idx, (size_t)p->nnz * 4);
        else {
            // regroup by draw size: the draws of a window are one block in both orders
            std::vector<int64_t> start(p->n_windows + 1, 0);
            for (int64_t w = 0; w < p->n_windows; ++w) start[w + 1] = start[w] + (p->h_doff[w + 1] - p->h_doff[w]) * p->h_n[w];
            int64_t at = 0;
            for (const SizeClass &k : p->classes)
                for (int32_t w : k.win) {
                    const int64_t cnt = (p->h_doff[w + 1] - p->h_doff[w]) * k.n;
                    memcpy(h + at, read_idx + start[w], (size_t)cnt * 4);
                    at += cnt;
                }
        }
        src = h;
    }
    // one draw size, statistics only, enough draws: the upload is chunked under the kernels (evaluate_classes)
    const bool chunked_up = single && !lik_out && p->n_draws >= 131072;
    int64_t at = 0;
    if (!chunked_up)
        for (const SizeClass &k : p->classes) {
            const int64_t cnt = k.n_draws * k.n;
            LDPGTA_CUDA(cudaMemcpyAsync(k.d_idx, src + at, (size_t)cnt * 4, cudaMemcpyHostToDevice, c->stream));
            k_sanitize_idx<<<(unsigned)std::min<int64_t>((cnt / 4 + 255) / 256 + 1, (int64_t)c->sm_count * 4), 256, 0, c->stream>>>(k.d_idx, cnt, lim, c->d_flag);
            c->launches++;
            at += cnt;
        }
    p->picks_fresh = false;                                           // the draws are the caller's now
    p->rows_fresh = true;
    if (int rc = evaluate_classes(c, p, kind, proportions, lik_out ? lik_pinned : nullptr, chunked_up ? src : nullptr, lim)) return rc;
    LDPGTA_CUDA(cudaMemcpyAsync(c->h_flag, c->d_flag, 4, cudaMemcpyDeviceToHost, c->stream));
    if (int rc = reduce_and_download(c, p, lik_out, lik_pinned, llr_out, window_mean_var, segment_partials, segment_stats)) return rc;
    if (*c->h_flag) return fail_bad_index(c, read_idx, p->nnz);
    return 0;
}

int ldpgta_bootstrap_stats_dev(ldpgta_ctx *c, int kind, uint64_t seed, const double *proportions, double *segment_partials_dev)
{
    if (int rc = check_plan(c, kind, proportions)) return rc;
    LDPGTA_CUDA(cudaSetDevice(c->device));
    WindowPlan *p = c->plan;
    if (!p->has_windows) return fail(LDPGTA_E_STATE, "the resident plan has no window read lists");
    if (int rc = time_mark(c, 4)) return rc;
    if (int rc = sample_classes(c, p, kind, seed, false)) return rc;
    if (int rc = time_mark(c, 4)) return rc;
    if (int rc = evaluate_classes(c, p, kind, proportions, nullptr, nullptr, 0, true)) return rc;
    // the resident likelihoods change and the LLR matrix is not written: what ldpgta_bin_stats would read is gone
    if (c->llr_ptr == p->d_llr) { c->llr_draws = c->llr_windows = -1; c->llr_lazy_lik = nullptr; }
    if (int rc = time_mark(c, 2)) return rc;
    const int rc_stats = stats_on_device(c, llr_fused(p) ? nullptr : p->d_lik, p->d_doff, p->n_windows, p->n_draws, p->d_segoff, p->d_segcols, p->d_segfirst, p->n_segments,
                           p->ncp, p->d_llr, p->d_mv, p->d_wsum, p->d_split, p->d_arrived, segment_partials_dev ? segment_partials_dev : p->d_partials,
                           segment_partials_dev ? nullptr : p->d_stats, nullptr, p->max_draws, p, false);
    if (rc_stats) return rc_stats;
    return time_mark(c, 2);
}

int ldpgta_finish_segments_dev(ldpgta_ctx *c, const double *segment_partials_dev, double *segment_stats_dev)
{
    if (!c || !c->plan) return fail(LDPGTA_E_STATE, "no windows are resident: call ldpgta_set_windows first");
    LDPGTA_CUDA(cudaSetDevice(c->device));
    WindowPlan *p = c->plan;
    k_finish_segments<<<(p->n_segments * 5 * 32 + 127) / 128, 128, 0, c->stream>>>(segment_partials_dev ? segment_partials_dev : p->d_partials,
                                                                            p->d_segcols, p->d_segfirst, p->n_segments, p->ncp,
                                                                            segment_stats_dev ? segment_stats_dev : p->d_stats);
    c->launches++;
    LDPGTA_CUDA(cudaGetLastError());
    return 0;
}

int ldpgta_kernel_timing(ldpgta_ctx *c, int enable)
{
    if (!c) return fail(LDPGTA_E_INVALID, "null context");
    const int mode = enable & 0xff, stride = (enable >> 8) & 0xffff;
    c->time_kernels = mode >= 2 && mode <= 4 ? mode : (mode != 0);
    c->t_stride = stride > 1 ? stride : 1;
    c->t_marks = 0;
    c->t_used = 0;
    return 0;
}

int ldpgta_kernel_times(ldpgta_ctx *c, double *ms_out, int max_entries)
{
    if (!c || (!ms_out && max_entries > 0)) return fail(LDPGTA_E_INVALID, "bad arguments");
    LDPGTA_CUDA(cudaSetDevice(c->device));
    LDPGTA_CUDA(cudaStreamSynchronize(c->stream));
    const int n = (int)(c->t_used / 2);
    for (int i = 0; i < n && i < max_entries; ++i) {
        float ms = 0.f;
        LDPGTA_CUDA(cudaEventElapsedTime(&ms, c->t_ev[2 * i], c->t_ev[2 * i + 1]));
        ms_out[i] = (double)ms;
    }
    c->t_used = 0;
    return n;
}

int ldpgta_plan_info(const ldpgta_ctx *c, int64_t *out)
{
    if (!c || !out) return fail(LDPGTA_E_INVALID, "null argument");
    if (!c->plan) return fail(LDPGTA_E_STATE, "no windows are resident: call ldpgta_set_windows first");
    const WindowPlan *p = c->plan;
    out[0] = p->n_windows; out[1] = p->n_draws; out[2] = p->n_segments; out[3] = p->ncp;
    out[4] = (int64_t)p->classes.size(); out[5] = p->nnz;
    out[6] = (int64_t)(uintptr_t)p->d_lik; out[7] = (int64_t)(uintptr_t)p->d_mv;
    out[8] = (int64_t)(uintptr_t)p->d_partials; out[9] = (int64_t)(uintptr_t)p->d_stats;
    return 0;
}

int ldpgta_window_stats(ldpgta_ctx *c, const double *lik, const int64_t *window_offsets, int64_t n_windows, int n_cols,
                        double *llr_out, double *window_mean_var, double *chrom_partials)
{
    if (!c || !lik || !window_offsets || n_windows < 0 || n_cols < 0) return fail(LDPGTA_E_INVALID, "bad arguments");
    LDPGTA_CUDA(cudaSetDevice(c->device));
    if (window_offsets[0] != 0) return fail(LDPGTA_E_INVALID, "window_offsets must start at 0");
    const int64_t n_draws = window_offsets[n_windows];
    for (int64_t w = 0; w < n_windows; ++w) {
        const int64_t k = window_offsets[w + 1] - window_offsets[w];
        if (k < 1) return fail(LDPGTA_E_INVALID, "window %lld has no draws", (long long)w);
        if (k < n_cols) return fail(LDPGTA_E_INVALID, "window %lld has %lld draws < n_cols=%d", (long long)w, (long long)k, n_cols);
    }
    const int stride = n_cols + 3;
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off += pad256(bytes); return o; };
    const size_t o_lik = take((size_t)n_draws * 32), o_seg = take(16), o_meta = take(8), o_mv = take((size_t)n_windows * 80);
    const size_t o_wsum = take((size_t)n_windows * 40), o_cp = take((size_t)5 * stride * 8), o_split = take((size_t)5 * SEG_SPLIT * stride * 8), o_arr = take(32);
    if (int rc = ensure_dev(c, off)) return rc;
    // the LLRs and the window offsets stay on the device for ldpgta_bin_stats
    c->llr_draws = c->llr_windows = -1;
    if (int rc = ensure_own(&c->d_llr, &c->d_llr_bytes, (size_t)n_draws * 40 + 16)) return rc;
    if (int rc = ensure_own(&c->d_llr_off, &c->d_llr_off_bytes, (size_t)(n_windows + 1) * 8)) return rc;
    char *base = (char *)c->d_buf;
    double *d_lik = (double *)(base + o_lik), *d_mv = (double *)(base + o_mv), *d_wsum = (double *)(base + o_wsum), *d_cp = (double *)(base + o_cp);
    int64_t *d_seg = (int64_t *)(base + o_seg);
    int32_t *d_meta = (int32_t *)(base + o_meta);
    const int64_t h_seg[2] = {0, n_windows};
    const int32_t h_meta[2] = {n_cols, n_windows ? (int32_t)(window_offsets[1] - window_offsets[0]) : 0};
    if (n_draws) LDPGTA_CUDA(cudaMemcpyAsync(d_lik, lik, (size_t)n_draws * 32, cudaMemcpyHostToDevice, c->stream));
    LDPGTA_CUDA(cudaMemcpyAsync(c->d_llr_off, window_offsets, (size_t)(n_windows + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    LDPGTA_CUDA(cudaMemcpyAsync(d_seg, h_seg, 16, cudaMemcpyHostToDevice, c->stream));
    LDPGTA_CUDA(cudaMemcpyAsync(d_meta, h_meta, 8, cudaMemcpyHostToDevice, c->stream));
    LDPGTA_CUDA(cudaMemsetAsync(base + o_arr, 0, 32, c->stream));
    int64_t most = 0;
    for (int64_t w = 0; w < n_windows; ++w) most = std::max(most, window_offsets[w + 1] - window_offsets[w]);
    if (int rc = stats_on_device(c, d_lik, c->d_llr_off, n_windows, n_draws, d_seg, d_meta, d_meta + 1, 1, n_cols, c->d_llr, d_mv,
                                 d_wsum, (double *)(base + o_split), (unsigned int *)(base + o_arr), d_cp, nullptr, nullptr, most))
        return rc;
    if (llr_out && n_draws) LDPGTA_CUDA(cudaMemcpyAsync(llr_out, c->d_llr, (size_t)n_draws * 40, cudaMemcpyDeviceToHost, c->stream));
    if (window_mean_var && n_windows) LDPGTA_CUDA(cudaMemcpyAsync(window_mean_var, d_mv, (size_t)n_windows * 80, cudaMemcpyDeviceToHost, c->stream));
    if (chrom_partials) LDPGTA_CUDA(cudaMemcpyAsync(chrom_partials, d_cp, (size_t)5 * stride * 8, cudaMemcpyDeviceToHost, c->stream));
    LDPGTA_CUDA(cudaStreamSynchronize(c->stream));
    c->llr_ptr = c->d_llr; c->llr_off_ptr = c->d_llr_off;
    c->llr_draws = n_draws;
    c->llr_windows = n_windows;
    return 0;
}

}  // extern "C"
