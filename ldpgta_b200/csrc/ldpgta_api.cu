// ldpgta_api.cu -- the C-ABI of include/ldpgta.h on top of the sm_100a kernels.
// Build (see __graft_entry__.build):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -shared -Xcompiler -fPIC
#include <stdarg.h>
#include <math.h>
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "k_small.cuh"     // SmallArgs
#include "k_general.cuh"   // GeneralArgs
#include "k_stats.cuh"

namespace ldpgta {

thread_local std::string g_last_error;

int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

int ensure_dev(ldpgta_ctx *c, size_t bytes)
{
    if (c->d_buf_bytes >= bytes) return 0;
    if (c->d_buf) LDPGTA_CUDA(cudaFree(c->d_buf));
    c->d_buf = nullptr; c->d_buf_bytes = 0;
    bytes = bytes + bytes / 4 + 4096;
    LDPGTA_CUDA(cudaMalloc(&c->d_buf, bytes));
    c->d_buf_bytes = bytes;
    return 0;
}

int ensure_pin(ldpgta_ctx *c, size_t bytes)
{
    if (c->h_pin_bytes >= bytes) return 0;
    if (c->h_pin) LDPGTA_CUDA(cudaFreeHost(c->h_pin));
    c->h_pin = nullptr; c->h_pin_bytes = 0;
    bytes = bytes + bytes / 4 + 4096;
    LDPGTA_CUDA(cudaMallocHost(&c->h_pin, bytes));
    c->h_pin_bytes = bytes;
    return 0;
}

// copy streams and per-chunk events of the pipelined host path
int ensure_pipeline(ldpgta_ctx *c, int n_chunks)
{
    if (!c->s_in) LDPGTA_CUDA(cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking));
    if (!c->s_out) LDPGTA_CUDA(cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking));
    if (!c->d_flag) LDPGTA_CUDA(cudaMalloc((void **)&c->d_flag, 16));
    if (!c->h_flag) LDPGTA_CUDA(cudaMallocHost((void **)&c->h_flag, 16));
    if (!c->ev_fork) {
        LDPGTA_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
        LDPGTA_CUDA(cudaEventCreateWithFlags(&c->ev_join_in, cudaEventDisableTiming));
        LDPGTA_CUDA(cudaEventCreateWithFlags(&c->ev_join_out, cudaEventDisableTiming));
    }
    while ((int)c->ev_in.size() < n_chunks) {
        cudaEvent_t a, b;
        LDPGTA_CUDA(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
        LDPGTA_CUDA(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
        c->ev_in.push_back(a);
        c->ev_k.push_back(b);
    }
    return 0;
}

// page-locked (or managed) host memory: usable by asynchronous copies without staging and readable on the host
bool is_page_locked(const void *p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged;
}

// (re)computes the per-read popcount table of group g after its masks changed
static int refresh_popc(ldpgta_ctx *c, int g)
{
    if (c->read_popc[g]) { LDPGTA_CUDA(cudaFree(c->read_popc[g])); c->read_popc[g] = nullptr; }
    c->both_valid = false;                                    // the combined two-group rows are rebuilt on their next use
    c->masks_epoch++;
    const int64_t n = c->n_reads_g[g];
    LDPGTA_CUDA(cudaMalloc((void **)&c->read_popc[g], std::max<size_t>((size_t)n * 4, 16)));
    if (n) {
        const unsigned blocks = (unsigned)std::min<int64_t>((n * 32 + 255) / 256, (int64_t)c->sm_count * 16);
        k_row_popc<<<blocks, 256, 0, c->stream>>>(c->read_masks[g], n, c->wp[g], c->read_popc[g]);
        c->launches++;
        LDPGTA_CUDA(cudaGetLastError());
    }
    return make_read_mask_tmap(c, g);
}

// Index arrays are validated where they are used: the uploaded copy is checked (and clamped) by k_sanitize_idx on the
// stream, the caller reads c->h_flag after its synchronisation and only then looks for the offending entry on the host.
static int check_idx_on_device(ldpgta_ctx *c, int32_t *d_idx, int64_t count, int64_t lim, bool first)
{
    if (int rc = ensure_pipeline(c, 1)) return rc;
    if (first) LDPGTA_CUDA(cudaMemsetAsync(c->d_flag, 0, 4, c->stream));
    if (count > 0) {
        k_sanitize_idx<<<(unsigned)std::min<int64_t>((count / 4 + 255) / 256 + 1, (int64_t)c->sm_count * 4), 256, 0, c->stream>>>(
            d_idx, count, (uint32_t)std::min<int64_t>(lim, 0x7fffffff), c->d_flag);
        c->launches++;
    }
    LDPGTA_CUDA(cudaMemcpyAsync(c->h_flag, c->d_flag, 4, cudaMemcpyDeviceToHost, c->stream));
    return 0;
}

static int fail_bad_allele(const int32_t *allele_idx, int64_t nnz, int64_t n_alleles)
{
    for (int64_t k = 0; k < nnz; ++k)
        if (allele_idx[k] < 0 || allele_idx[k] >= n_alleles)
            return fail(LDPGTA_E_INVALID, "allele index %d at %lld is outside 0..%lld", allele_idx[k], (long long)k, (long long)n_alleles - 1);
    return fail(LDPGTA_E_INVALID, "an allele index is outside 0..%lld", (long long)n_alleles - 1);
}

int check_kind(const ldpgta_ctx *c, int kind, const double *proportions)
{
    if (kind == LDPGTA_HOMOGENEOUS && c->G != 1)
        return fail(LDPGTA_E_INVALID, "homogeneous model needs exactly one reference-panel group, context has %d", c->G);
    if (kind == LDPGTA_RECENT_ADMIXTURE && c->G != 2)
        return fail(LDPGTA_E_INVALID, "recent-admixture model needs exactly two reference-panel groups, context has %d", c->G);
    if (kind == LDPGTA_DISTANT_ADMIXTURE && !proportions)
        return fail(LDPGTA_E_INVALID, "distant-admixture model needs ancestral proportions");
    if (kind < 0 || kind > 2) return fail(LDPGTA_E_INVALID, "unknown model kind %d", kind);
    for (int g = 0; g < c->G; ++g)
        if (!c->read_masks[g] || c->n_reads_g[g] != c->n_reads)
            return fail(LDPGTA_E_STATE, "read masks of group %d are missing (upload or build them first)", g);
    return 0;
}

// ------------------------------------------------------------------ launchers

// draws of one size, device-resident indices and outputs
int run_uniform(ldpgta_ctx *c, int kind, int n, const int32_t *idx_dev, int64_t n_draws, const int64_t *pos_dev,
                const double *proportions, double *out_dev, uint32_t *counts_dev, bool force_general,
                const PanelView *override_view, double *llr_dev, int64_t llr_stride)
{
    if (n_draws == 0) return 0;
    if (n < 2 || n > MAXN) return fail(LDPGTA_E_UNSUPPORTED, "draws of %d reads are outside the supported range 2..%d", n, MAXN);
    if (n <= 4 && !force_general) {
        SmallArgs a;
        a.pv = override_view ? *override_view : make_view(c, proportions);
        a.idx = idx_dev; a.n_draws = n_draws; a.pos = pos_dev; a.out = out_dev; a.counts_out = counts_dev;
        a.llr = llr_dev; a.llr_stride = llr_stride;
        return launch_small(c, a, n, kind);
    }
    GeneralArgs a;
    a.pv = override_view ? *override_view : make_view(c, proportions);
    a.kind = kind; a.n = n; a.kl = a.kh = a.wc = 0; a.etab = 0; a.etab_scratch = nullptr; a.etab_stride = 0;
    a.idx = idx_dev; a.n_draws = n_draws; a.pos = pos_dev; a.out = out_dev; a.counts_out = counts_dev;
    a.llr = llr_dev; a.llr_stride = llr_stride;
    a.terms = nullptr; a.n_terms = 0; a.wide = 0; a.table_scratch = nullptr; a.table_stride = 0;
    // measured (tools/bench_lanes.sh): k_mid leads at 5 reads of a one-group model, k_lanes from 6 reads on
    const bool mid_ok = (n == 5 || n == 6) && kind == LDPGTA_HOMOGENEOUS && !force_general;
    if (n >= 6 && n <= 12) {
        const int rc = launch_lanes(c, a);
        if (rc <= 0) return rc;                       // 1 = outside the register kernel's envelope
    }
    if (mid_ok) {
        SmallArgs sa;
        sa.pv = a.pv;
        sa.idx = idx_dev; sa.n_draws = n_draws; sa.pos = pos_dev; sa.out = out_dev; sa.counts_out = counts_dev;
        sa.llr = llr_dev; sa.llr_stride = llr_stride;
        const int rc = launch_mid(c, sa, n);
        if (rc <= 0) return rc;
    }
    if (n == 5) {
        const int rc = launch_lanes(c, a);
        if (rc <= 0) return rc;
    }
    return launch_general(c, a);
}

// message for the first read index outside the table (slow path, only after a failed validation)
int fail_bad_index(const ldpgta_ctx *c, const int32_t *read_idx, int64_t nnz)
{
    for (int64_t k = 0; k < nnz; ++k)
        if (read_idx[k] < 0 || read_idx[k] >= c->n_reads)
            return fail(LDPGTA_E_INVALID, "read index %d at %lld is outside 0..%lld", read_idx[k], (long long)k, (long long)c->n_reads - 1);
    return fail(LDPGTA_E_INVALID, "a read index is outside 0..%lld", (long long)c->n_reads - 1);
}

// Uniform draws from page-locked host buffers: chunks flow through upload -> validation + kernel -> download on
// three streams.  Nothing is scanned on the host before the first copy: indices are validated on the device, and
// draw_offsets (optional) are checked chunk by chunk while earlier chunks are in flight.
// Returns 0, a negative error, or 1 when draw_offsets turn out not to be uniform (the caller regroups).
// enqueues every chunk of the pipeline on the three streams (eagerly, or into a stream capture)
static int issue_pipeline(ldpgta_ctx *c, int kind, int n0, const int32_t *read_idx, int64_t n_draws, const double *proportions,
                          double *out, int64_t per, int32_t *d_idx, double *d_out)
{
    const uint32_t lim = (uint32_t)std::min<int64_t>(c->n_reads, 0x7fffffff);
    LDPGTA_CUDA(cudaMemsetAsync(c->d_flag, 0, 4, c->stream));
    int ci = 0;
    for (int64_t d0 = 0; d0 < n_draws; d0 += per, ++ci) {
        const int64_t nd = std::min(per, n_draws - d0);
        LDPGTA_CUDA(cudaMemcpyAsync(d_idx + d0 * n0, read_idx + d0 * n0, (size_t)nd * n0 * 4, cudaMemcpyHostToDevice, c->s_in));
        LDPGTA_CUDA(cudaEventRecord(c->ev_in[ci], c->s_in));
        LDPGTA_CUDA(cudaStreamWaitEvent(c->stream, c->ev_in[ci], 0));
        const int64_t cnt = nd * n0;
        k_sanitize_idx<<<(unsigned)std::min<int64_t>((cnt / 4 + 255) / 256 + 1, (int64_t)c->sm_count * 4), 256, 0, c->stream>>>(
            d_idx + d0 * n0, cnt, lim, c->d_flag);
        c->launches++;
        if (int rc = run_uniform(c, kind, n0, d_idx + d0 * n0, nd, nullptr, proportions, d_out + 4 * d0, nullptr)) return rc;
        LDPGTA_CUDA(cudaEventRecord(c->ev_k[ci], c->stream));
        LDPGTA_CUDA(cudaStreamWaitEvent(c->s_out, c->ev_k[ci], 0));
        LDPGTA_CUDA(cudaMemcpyAsync(out + 4 * d0, d_out + 4 * d0, (size_t)nd * 32, cudaMemcpyDeviceToHost, c->s_out));
    }
    LDPGTA_CUDA(cudaMemcpyAsync(c->h_flag, c->d_flag, 4, cudaMemcpyDeviceToHost, c->s_out));
    return 0;
}

static uint64_t fnv1a(const void *p, size_t n, uint64_t h = 1469598103934665603ull)
{
    const unsigned char *b = (const unsigned char *)p;
    for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}

// Uniform draws from page-locked host buffers: chunks flow through upload -> validation + kernel -> download on
// three streams.  Nothing is scanned on the host before the first copy: indices are validated on the device, and
// draw_offsets (optional) are checked on the host while the GPU works.  A call that repeats the previous one's shape
// and buffers (the streaming case: the caller refills the same page-locked buffers) replays the whole pipeline as ONE
// CUDA graph launch instead of ~60 stream operations.
// Returns 0, a negative error, or 1 when draw_offsets turn out not to be uniform (the caller regroups).
static int run_pipelined(ldpgta_ctx *c, int kind, int n0, const int32_t *read_idx, const int64_t *draw_offsets, int64_t n_draws,
                         const double *proportions, double *out)
{
    static const int max_chunks = [] { const char *e = getenv("LDPGTA_CHUNKS"); return e ? std::max(1, atoi(e)) : 8; }();
    static const bool use_graph = getenv("LDPGTA_NO_GRAPH") == nullptr;
    const size_t idx_bytes = ((size_t)n_draws * n0 * 4 + 15) & ~(size_t)15;
    if (int rc = ensure_dev(c, idx_bytes + (size_t)n_draws * 32)) return rc;
    int32_t *d_idx = (int32_t *)c->d_buf;
    double *d_out = (double *)((char *)c->d_buf + idx_bytes);
    const int64_t n_chunks = std::max<int64_t>(1, std::min<int64_t>(max_chunks, n_draws / 32768));
    const int64_t per = ((n_draws + n_chunks - 1) / n_chunks + 63) & ~(int64_t)63;
    if (int rc = ensure_pipeline(c, (int)n_chunks + 1)) return rc;

    // everything the enqueued operations depend on
    struct { int kind, n0, G; int64_t n_draws, per, n_reads; const void *idx, *out, *dbuf, *scratch, *escratch; PanelView pv; } sig;
    memset(&sig, 0, sizeof sig);
    sig.kind = kind; sig.n0 = n0; sig.G = c->G; sig.n_draws = n_draws; sig.per = per; sig.n_reads = c->n_reads;
    sig.idx = read_idx; sig.out = out; sig.dbuf = c->d_buf; sig.scratch = c->table_scratch; sig.escratch = c->etab_scratch;
    sig.pv = make_view(c, proportions);
    const uint64_t key = fnv1a(&sig, sizeof sig) | 1ull;

    int rc = 0;
    bool launched = false;
    if (use_graph && c->pipe_exec && c->pipe_key == key) {
        LDPGTA_CUDA(cudaGraphLaunch(c->pipe_exec, c->stream));
        c->launches += c->pipe_launches;
        c->pipe_replays++;
        launched = true;
    } else if (use_graph && c->pipe_seen == key && c->pipe_bad != key) {
        // second call in a row with this signature: every lazy allocation of the launchers has happened, capture it
        if (c->pipe_exec) { cudaGraphExecDestroy(c->pipe_exec); c->pipe_exec = nullptr; c->pipe_key = 0; }
        const int64_t l0 = c->launches;
        cudaGraph_t graph = nullptr;
        bool ok = cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeRelaxed) == cudaSuccess;
        if (ok) {
            ok = cudaEventRecord(c->ev_fork, c->stream) == cudaSuccess && cudaStreamWaitEvent(c->s_in, c->ev_fork, 0) == cudaSuccess &&
                 cudaStreamWaitEvent(c->s_out, c->ev_fork, 0) == cudaSuccess;
            if (ok) ok = issue_pipeline(c, kind, n0, read_idx, n_draws, proportions, out, per, d_idx, d_out) == 0;
            if (ok)
                ok = cudaEventRecord(c->ev_join_in, c->s_in) == cudaSuccess && cudaStreamWaitEvent(c->stream, c->ev_join_in, 0) == cudaSuccess &&
                     cudaEventRecord(c->ev_join_out, c->s_out) == cudaSuccess && cudaStreamWaitEvent(c->stream, c->ev_join_out, 0) == cudaSuccess;
            const cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
            ok = ok && e == cudaSuccess && graph != nullptr;
        }
        const int64_t per_graph = c->launches - l0;
        c->launches = l0;
        if (ok) ok = cudaGraphInstantiate(&c->pipe_exec, graph, 0) == cudaSuccess;
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        if (ok) {
            c->pipe_key = key;
            c->pipe_launches = per_graph;
            LDPGTA_CUDA(cudaGraphLaunch(c->pipe_exec, c->stream));
            c->launches += per_graph;
            c->pipe_replays++;
            launched = true;
        } else {
            c->pipe_exec = nullptr;
            c->pipe_bad = key;                         // this signature does not capture: keep issuing it eagerly
        }
    }
    if (!launched) {
        rc = issue_pipeline(c, kind, n0, read_idx, n_draws, proportions, out, per, d_idx, d_out);
        c->pipe_seen = rc == 0 ? key : 0;
    }
    // meanwhile on the host: are the draw sizes really uniform?
    bool ragged = false;
    if (draw_offsets && rc == 0) {
        int64_t differs = 0;
        for (int64_t d = 0; d <= n_draws; ++d) differs |= draw_offsets[d] ^ (d * n0);
        ragged = differs != 0;
    }
    LDPGTA_CUDA(cudaStreamSynchronize(c->s_in));
    LDPGTA_CUDA(cudaStreamSynchronize(c->stream));
    LDPGTA_CUDA(cudaStreamSynchronize(c->s_out));
    if (rc) return rc;
    if (ragged) return 1;
    if (*c->h_flag) return fail_bad_index(c, read_idx, n_draws * n0);
    return 0;
}

int check_uniform_size(int kind, int64_t n)
{
    if (n < 2 || n > MAXN)
        return fail(LDPGTA_E_UNSUPPORTED, "draw 0 has %lld reads, supported range is 2..%d", (long long)n, MAXN);
    if (kind == LDPGTA_DISTANT_ADMIXTURE && n == 5)
        return fail(LDPGTA_E_UNSUPPORTED, "distant_admixture has no likelihoods5 (reference: DISTANT_ADMIXTURE_MODELS.py:312-313 raises AttributeError)");
    return 0;
}

}  // namespace ldpgta

using namespace ldpgta;

// ============================================================================ C-ABI

extern "C" {

const char *ldpgta_last_error(void) { return g_last_error.c_str(); }
int ldpgta_version(void) { return 100; }

int ldpgta_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int ldpgta_create(int device, int n_groups, const uint32_t *n_hap, ldpgta_ctx **out)
{
    if (!out || !n_hap) return fail(LDPGTA_E_INVALID, "null argument");
    if (n_groups < 1 || n_groups > MAXG) return fail(LDPGTA_E_INVALID, "n_groups must be 1..%d, got %d", MAXG, n_groups);
    for (int g = 0; g < n_groups; ++g)
        if (n_hap[g] == 0) return fail(LDPGTA_E_INVALID, "group %d has no haplotypes", g);
    int ndev = ldpgta_device_count();
    if (ndev < 1) return fail(LDPGTA_E_CUDA, "no CUDA device is visible: the likelihood engine has no CPU path");
    if (device < 0 || device >= ndev) return fail(LDPGTA_E_INVALID, "device %d out of range (0..%d)", device, ndev - 1);
    LDPGTA_CUDA(cudaSetDevice(device));
    ldpgta_ctx *c = new ldpgta_ctx();
    c->device = device; c->G = n_groups;
    for (int g = 0; g < n_groups; ++g) {
        c->nhap[g] = n_hap[g];
        c->words[g] = (int)((n_hap[g] + 63u) / 64u);
        c->wp[g] = (c->words[g] + 1) & ~1;
    }
    cudaDeviceProp prop;
    LDPGTA_CUDA(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    c->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    LDPGTA_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    for (int g = 0; g < n_groups; ++g) {
        LDPGTA_CUDA(cudaMalloc((void **)&c->norm[g], ((size_t)n_hap[g] + 1) * 8));
        k_norm_table<<<(n_hap[g] + 256) / 256, 256, 0, c->stream>>>(c->norm[g], n_hap[g]);
        c->launches++;
    }
    LDPGTA_CUDA(cudaGetLastError());
    LDPGTA_CUDA(cudaStreamSynchronize(c->stream));
    *out = c;
    return 0;
}

int ldpgta_destroy(ldpgta_ctx *c)
{
    if (!c) return 0;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    free_window_plan(c);
    free_comm(c);
    for (int g = 0; g < c->G; ++g) {
        if (c->allele_rows[g]) cudaFree(c->allele_rows[g]);
        if (c->read_masks[g] && c->masks_owned[g]) cudaFree(c->read_masks[g]);
        if (c->read_popc[g]) cudaFree(c->read_popc[g]);
        if (c->norm[g]) cudaFree(c->norm[g]);
    }
    for (int h = 0; h < 14; ++h)
        if (c->hi_terms[h]) cudaFree(c->hi_terms[h]);
    if (c->masks_both) cudaFree(c->masks_both);
    for (int h = 0; h < 14; ++h)
        for (int b = 0; b < 4; ++b)
            if (c->run_terms[h][b]) cudaFree(c->run_terms[h][b]);
    if (c->table_scratch) cudaFree(c->table_scratch);
    if (c->allele_complement) cudaFree(c->allele_complement);
    if (c->d_buf) cudaFree(c->d_buf);
    if (c->d_llr) cudaFree(c->d_llr);
    if (c->etab_scratch) cudaFree(c->etab_scratch);
    if (c->pipe_exec) cudaGraphExecDestroy(c->pipe_exec);
    if (c->ev_fork) { cudaEventDestroy(c->ev_fork); cudaEventDestroy(c->ev_join_in); cudaEventDestroy(c->ev_join_out); }
    if (c->d_flag) cudaFree(c->d_flag);
    if (c->h_flag) cudaFreeHost(c->h_flag);
    if (c->d_llr_off) cudaFree(c->d_llr_off);
    if (c->h_pin) cudaFreeHost(c->h_pin);
    if (c->h_io) cudaFreeHost(c->h_io);
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->s_in) cudaStreamDestroy(c->s_in);
    if (c->s_out) cudaStreamDestroy(c->s_out);
    for (cudaEvent_t e : c->ev_in) cudaEventDestroy(e);
    for (cudaEvent_t e : c->ev_k) cudaEventDestroy(e);
    for (cudaEvent_t e : c->t_ev) cudaEventDestroy(e);
    delete c;
    return 0;
}

int ldpgta_row_words(const ldpgta_ctx *c, int group)
{
    if (!c || group < 0 || group >= c->G) return fail(LDPGTA_E_INVALID, "bad context or group");
    return c->words[group];
}

int ldpgta_padded_row_words(const ldpgta_ctx *c, int group)
{
    if (!c || group < 0 || group >= c->G) return fail(LDPGTA_E_INVALID, "bad context or group");
    return c->wp[group];
}

int64_t ldpgta_num_reads(const ldpgta_ctx *c) { return c ? c->n_reads : 0; }
uint64_t ldpgta_stream(const ldpgta_ctx *c) { return c ? (uint64_t)(uintptr_t)c->stream : 0; }
int64_t ldpgta_graph_replays(const ldpgta_ctx *c) { return c ? c->pipe_replays : 0; }

int64_t ldpgta_launch_count(const ldpgta_ctx *c) { return c ? c->launches : 0; }

int ldpgta_synchronize(ldpgta_ctx *c)
{
    if (!c) return fail(LDPGTA_E_INVALID, "null context");
    LDPGTA_CUDA(cudaSetDevice(c->device));
    LDPGTA_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

static int pack_to_device(ldpgta_ctx *c, int group, const uint64_t *rows, int64_t n_rows, const uint8_t *complement,
                          uint64_t **dst)
{
    const int words = c->words[group], wp = c->wp[group];
    const size_t src_bytes = (size_t)n_rows * words * 8, flag_bytes = complement ? (size_t)n_rows : 0;
    const size_t flag_off = (src_bytes + 15) & ~(size_t)15;
    if (int rc = ensure_dev(c, flag_off + flag_bytes + 16)) return rc;
    if (*dst) { LDPGTA_CUDA(cudaFree(*dst)); *dst = nullptr; }
    LDPGTA_CUDA(cudaMalloc((void **)dst, std::max<size_t>((size_t)n_rows * wp * 8, 16)));
    if (n_rows == 0) return 0;
    LDPGTA_CUDA(cudaMemcpyAsync(c->d_buf, rows, src_bytes, cudaMemcpyHostToDevice, c->stream));
    if (complement)
        LDPGTA_CUDA(cudaMemcpyAsync((char *)c->d_buf + flag_off, complement, flag_bytes, cudaMemcpyHostToDevice, c->stream));
    const int64_t total = n_rows * wp;
    const unsigned blocks = (unsigned)std::min<int64_t>((total + 255) / 256, (int64_t)c->sm_count * 16);
    k_pack_rows<<<blocks, 256, 0, c->stream>>>((const uint64_t *)c->d_buf,
                                               complement ? (const uint8_t *)c->d_buf + flag_off : nullptr, *dst, n_rows,
                                               words, wp, c->nhap[group]);
    c->launches++;
    LDPGTA_CUDA(cudaGetLastError());
    LDPGTA_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

int ldpgta_upload_allele_rows(ldpgta_ctx *c, int group, const uint64_t *rows, int64_t n_rows, const uint8_t *complement)
{
    if (!c || group < 0 || group >= c->G) return fail(LDPGTA_E_INVALID, "bad context or group");
    if (n_rows < 0 || (n_rows > 0 && !rows)) return fail(LDPGTA_E_INVALID, "bad rows");
    LDPGTA_CUDA(cudaSetDevice(c->device));
    if (int rc = pack_to_device(c, group, rows, n_rows, complement, &c->allele_rows[group])) return rc;
    c->n_alleles[group] = n_rows;
    if (group == 0) {                                  // the flags are a property of the allele, not of the group
        if (c->allele_complement) { LDPGTA_CUDA(cudaFree(c->allele_complement)); c->allele_complement = nullptr; }
        LDPGTA_CUDA(cudaMalloc((void **)&c->allele_complement, std::max<size_t>((size_t)n_rows, 16)));
        if (complement) LDPGTA_CUDA(cudaMemcpy(c->allele_complement, complement, (size_t)n_rows, cudaMemcpyHostToDevice));
        else LDPGTA_CUDA(cudaMemset(c->allele_complement, 0, std::max<size_t>((size_t)n_rows, 16)));
    }
    return 0;
}

struct ldpgta_panel {
    int device = 0, G = 0;
    uint32_t nhap[MAXG] = {};
    int words[MAXG] = {}, wp[MAXG] = {};
    int64_t n_snps = 0;
    uint64_t *rows[MAXG] = {};
    cudaStream_t stream = nullptr;
    uint64_t *stage = nullptr;
    size_t stage_bytes = 0;
    int sm_count = 148;
};

int ldpgta_panel_create(int device, int n_groups, const uint32_t *n_hap, int64_t n_snps, ldpgta_panel **out)
{
    if (!out || !n_hap || n_snps < 0) return fail(LDPGTA_E_INVALID, "bad arguments");
    if (n_groups < 1 || n_groups > MAXG) return fail(LDPGTA_E_INVALID, "n_groups must be 1..%d, got %d", MAXG, n_groups);
    const int ndev = ldpgta_device_count();
    if (ndev < 1) return fail(LDPGTA_E_CUDA, "no CUDA device is visible: the likelihood engine has no CPU path");
    if (device < 0 || device >= ndev) return fail(LDPGTA_E_INVALID, "device %d out of range (0..%d)", device, ndev - 1);
    LDPGTA_CUDA(cudaSetDevice(device));
    ldpgta_panel *p = new ldpgta_panel();
    p->device = device; p->G = n_groups; p->n_snps = n_snps;
    LDPGTA_CUDA(cudaDeviceGetAttribute(&p->sm_count, cudaDevAttrMultiProcessorCount, device));
    for (int g = 0; g < n_groups; ++g) {
        if (n_hap[g] == 0) { delete p; return fail(LDPGTA_E_INVALID, "group %d has no haplotypes", g); }
        p->nhap[g] = n_hap[g];
        p->words[g] = (int)((n_hap[g] + 63u) / 64u);
        p->wp[g] = (p->words[g] + 1) & ~1;
        LDPGTA_CUDA(cudaMalloc((void **)&p->rows[g], std::max<size_t>((size_t)n_snps * p->wp[g] * 8, 16)));
    }
    *out = p;
    return 0;
}

int ldpgta_panel_upload(ldpgta_panel *p, int group, int64_t first_snp, const uint64_t *rows, int64_t n_rows)
{
    if (!p || group < 0 || group >= p->G || n_rows < 0 || (n_rows > 0 && !rows)) return fail(LDPGTA_E_INVALID, "bad arguments");
    if (first_snp < 0 || first_snp + n_rows > p->n_snps) return fail(LDPGTA_E_INVALID, "SNP range %lld..%lld is outside the panel", (long long)first_snp, (long long)(first_snp + n_rows));
    if (n_rows == 0) return 0;
    LDPGTA_CUDA(cudaSetDevice(p->device));
    // staging buffer and stream belong to the panel: chunks of one upload reuse them, nothing touches the default stream
    const size_t src_bytes = (size_t)n_rows * p->words[group] * 8;
    if (!p->stream) LDPGTA_CUDA(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
    if (p->stage_bytes < src_bytes) {
        if (p->stage) LDPGTA_CUDA(cudaFree(p->stage));
        p->stage = nullptr; p->stage_bytes = 0;
        LDPGTA_CUDA(cudaMalloc((void **)&p->stage, src_bytes));
        p->stage_bytes = src_bytes;
    }
    cudaError_t e = cudaMemcpyAsync(p->stage, rows, src_bytes, cudaMemcpyHostToDevice, p->stream);
    if (e == cudaSuccess) {
        const int64_t total = n_rows * p->wp[group];
        const unsigned blocks = (unsigned)std::min<int64_t>((total + 255) / 256, (int64_t)p->sm_count * 16);
        k_pack_rows<<<blocks, 256, 0, p->stream>>>(p->stage, nullptr, p->rows[group] + first_snp * p->wp[group], n_rows, p->words[group],
                                                   p->wp[group], p->nhap[group]);
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaStreamSynchronize(p->stream);        // the caller's rows may be reused on return
    }
    if (e != cudaSuccess) return fail(LDPGTA_E_CUDA, "panel upload failed: %s", cudaGetErrorString(e));
    return 0;
}

int ldpgta_panel_destroy(ldpgta_panel *p)
{
    if (!p) return 0;
    cudaSetDevice(p->device);
    for (int g = 0; g < p->G; ++g)
        if (p->rows[g]) cudaFree(p->rows[g]);
    if (p->stage) cudaFree(p->stage);
    if (p->stream) cudaStreamDestroy(p->stream);
    delete p;
    return 0;
}

int ldpgta_gather_allele_rows(ldpgta_ctx *c, const ldpgta_panel *p, const int64_t *legend_idx, const uint8_t *complement,
                              int64_t n_alleles)
{
    if (!c || !p || n_alleles < 0 || (n_alleles > 0 && !legend_idx)) return fail(LDPGTA_E_INVALID, "bad arguments");
    if (p->device != c->device || p->G != c->G) return fail(LDPGTA_E_INVALID, "panel and context differ in device or number of groups");
    for (int g = 0; g < c->G; ++g)
        if (p->nhap[g] != c->nhap[g]) return fail(LDPGTA_E_INVALID, "panel group %d has %u haplotypes, context expects %u", g, p->nhap[g], c->nhap[g]);
    for (int64_t a = 0; a < n_alleles; ++a)
        if (legend_idx[a] < 0 || legend_idx[a] >= p->n_snps)
            return fail(LDPGTA_E_INVALID, "legend index %lld at %lld is outside 0..%lld", (long long)legend_idx[a], (long long)a, (long long)p->n_snps - 1);
    LDPGTA_CUDA(cudaSetDevice(c->device));
    const size_t idx_b = ((size_t)n_alleles * 8 + 15) & ~(size_t)15;
    if (int rc = ensure_dev(c, idx_b + (size_t)n_alleles + 16)) return rc;
    int64_t *d_idx = (int64_t *)c->d_buf;
    uint8_t *d_comp = (uint8_t *)c->d_buf + idx_b;
    if (n_alleles) {
        LDPGTA_CUDA(cudaMemcpyAsync(d_idx, legend_idx, (size_t)n_alleles * 8, cudaMemcpyHostToDevice, c->stream));
        if (complement) LDPGTA_CUDA(cudaMemcpyAsync(d_comp, complement, (size_t)n_alleles, cudaMemcpyHostToDevice, c->stream));
    }
    for (int g = 0; g < c->G; ++g) {
        if (c->allele_rows[g]) { LDPGTA_CUDA(cudaFree(c->allele_rows[g])); c->allele_rows[g] = nullptr; }
        LDPGTA_CUDA(cudaMalloc((void **)&c->allele_rows[g], std::max<size_t>((size_t)n_alleles * c->wp[g] * 8, 16)));
        if (n_alleles) {
            const int64_t total = n_alleles * c->wp[g];
            const unsigned blocks = (unsigned)std::min<int64_t>((total + 255) / 256, (int64_t)c->sm_count * 16);
            k_gather_rows<<<blocks, 256, 0, c->stream>>>(p->rows[g], d_idx, complement ? d_comp : nullptr, c->allele_rows[g], n_alleles,
                                                         c->wp[g], c->nhap[g]);
            c->launches++;
            LDPGTA_CUDA(cudaGetLastError());
        }
        c->n_alleles[g] = n_alleles;
    }
    if (c->allele_complement) { LDPGTA_CUDA(cudaFree(c->allele_complement)); c->allele_complement = nullptr; }
    LDPGTA_CUDA(cudaMalloc((void **)&c->allele_complement, std::max<size_t>((size_t)n_alleles, 16)));
    if (n_alleles) {
        if (complement) LDPGTA_CUDA(cudaMemcpyAsync(c->allele_complement, d_comp, (size_t)n_alleles, cudaMemcpyDeviceToDevice, c->stream));
        else LDPGTA_CUDA(cudaMemsetAsync(c->allele_complement, 0, (size_t)n_alleles, c->stream));
    }
    LDPGTA_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

int ldpgta_upload_read_masks(ldpgta_ctx *c, int group, const uint64_t *rows, int64_t n_rows)
{
    if (!c || group < 0 || group >= c->G) return fail(LDPGTA_E_INVALID, "bad context or group");
    if (n_rows < 0 || (n_rows > 0 && !rows)) return fail(LDPGTA_E_INVALID, "bad rows");
    LDPGTA_CUDA(cudaSetDevice(c->device));
    if (c->read_masks[group] && !c->masks_owned[group]) c->read_masks[group] = nullptr;
    if (int rc = pack_to_device(c, group, rows, n_rows, nullptr, &c->read_masks[group])) return rc;
    c->masks_owned[group] = true;
    c->n_reads_g[group] = n_rows;
    c->n_reads = n_rows;
    if (int rc = refresh_popc(c, group)) return rc;
    LDPGTA_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

int ldpgta_attach_read_masks_dev(ldpgta_ctx *c, int group, const uint64_t *rows_dev, int64_t n_rows)
{
    if (!c || group < 0 || group >= c->G) return fail(LDPGTA_E_INVALID, "bad context or group");
    if (!rows_dev || n_rows < 0) return fail(LDPGTA_E_INVALID, "bad rows");
    if (((uintptr_t)rows_dev & 15) != 0) return fail(LDPGTA_E_INVALID, "device rows must be 16-byte aligned");
    LDPGTA_CUDA(cudaSetDevice(c->device));
    if (c->read_masks[group] && c->masks_owned[group]) LDPGTA_CUDA(cudaFree(c->read_masks[group]));
    c->read_masks[group] = const_cast<uint64_t *>(rows_dev);
    c->masks_owned[group] = false;
    c->n_reads_g[group] = n_rows;
    c->n_reads = n_rows;
    if (int rc = refresh_popc(c, group)) return rc;
    LDPGTA_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

int ldpgta_build_read_masks(ldpgta_ctx *c, const int32_t *allele_idx, const int64_t *read_offsets, int64_t n_reads)
{
    if (!c || !read_offsets || n_reads < 0) return fail(LDPGTA_E_INVALID, "bad arguments");
    LDPGTA_CUDA(cudaSetDevice(c->device));
    const int64_t n_alleles = c->n_alleles[0];
    for (int g = 0; g < c->G; ++g) {
        if (!c->allele_rows[g]) return fail(LDPGTA_E_STATE, "allele rows of group %d were not uploaded", g);
        if (c->n_alleles[g] != n_alleles) return fail(LDPGTA_E_STATE, "groups hold different numbers of allele rows");
    }
    const int64_t nnz = read_offsets[n_reads];
    if (read_offsets[0] != 0) return fail(LDPGTA_E_INVALID, "read_offsets must start at 0");
    for (int64_t r = 0; r < n_reads; ++r)
        if (read_offsets[r + 1] < read_offsets[r]) return fail(LDPGTA_E_INVALID, "read_offsets must be non-decreasing");
    if (nnz > 0 && !allele_idx) return fail(LDPGTA_E_INVALID, "null allele_idx");
    const size_t idx_bytes = ((size_t)nnz * 4 + 15) & ~(size_t)15, off_bytes = (size_t)(n_reads + 1) * 8;
    if (int rc = ensure_dev(c, idx_bytes + off_bytes)) return rc;
    int32_t *d_idx = (int32_t *)c->d_buf;
    int64_t *d_off = (int64_t *)((char *)c->d_buf + idx_bytes);
    if (nnz) LDPGTA_CUDA(cudaMemcpyAsync(d_idx, allele_idx, (size_t)nnz * 4, cudaMemcpyHostToDevice, c->stream));
    if (int rc = check_idx_on_device(c, d_idx, nnz, n_alleles, true)) return rc;
    LDPGTA_CUDA(cudaMemcpyAsync(d_off, read_offsets, off_bytes, cudaMemcpyHostToDevice, c->stream));
    for (int g = 0; g < c->G; ++g) {
        if (c->read_masks[g] && c->masks_owned[g]) LDPGTA_CUDA(cudaFree(c->read_masks[g]));
        c->read_masks[g] = nullptr;
        LDPGTA_CUDA(cudaMalloc((void **)&c->read_masks[g], std::max<size_t>((size_t)n_reads * c->wp[g] * 8, 16)));
        c->masks_owned[g] = true;
        c->n_reads_g[g] = n_reads;
        if (n_reads) {
            const int64_t total = n_reads * (c->wp[g] >> 1);
            const unsigned blocks = (unsigned)std::min<int64_t>((total + 255) / 256, (int64_t)c->sm_count * 16);
            k_build_read_masks<<<blocks, 256, 0, c->stream>>>(c->allele_rows[g], d_idx, d_off, n_reads, c->wp[g], c->read_masks[g]);
            c->launches++;
            LDPGTA_CUDA(cudaGetLastError());
        }
        if (int rc = refresh_popc(c, g)) return rc;
    }
    c->n_reads = n_reads;
    LDPGTA_CUDA(cudaStreamSynchronize(c->stream));
    if (*c->h_flag) return fail_bad_allele(allele_idx, nnz, n_alleles);
    return 0;
}

int ldpgta_read_scores(ldpgta_ctx *c, const int32_t *allele_idx, const int64_t *read_offsets, int64_t n_reads,
                       const double *alpha, double score_weight, double min_HF, int32_t *out_scores)
{
    if (!c || !read_offsets || !alpha || !out_scores || n_reads < 0) return fail(LDPGTA_E_INVALID, "bad arguments");
    LDPGTA_CUDA(cudaSetDevice(c->device));
    const int64_t n_alleles = c->n_alleles[0];
    for (int g = 0; g < c->G; ++g) {
        if (!c->allele_rows[g]) return fail(LDPGTA_E_STATE, "allele rows of group %d were not uploaded", g);
        if (c->n_alleles[g] != n_alleles) return fail(LDPGTA_E_STATE, "groups hold different numbers of allele rows");
    }
    if (n_reads == 0) return 0;
    if (read_offsets[0] != 0) return fail(LDPGTA_E_INVALID, "read_offsets must start at 0");
    const int64_t nnz = read_offsets[n_reads];
    for (int64_t r = 0; r < n_reads; ++r)
        if (read_offsets[r + 1] < read_offsets[r]) return fail(LDPGTA_E_INVALID, "read_offsets must be non-decreasing");
    if (nnz > 0 && !allele_idx) return fail(LDPGTA_E_INVALID, "null allele_idx");
    const size_t idx_b = ((size_t)nnz * 4 + 15) & ~(size_t)15, off_b = ((size_t)(n_reads + 1) * 8 + 15) & ~(size_t)15;
    if (int rc = ensure_dev(c, idx_b + off_b + (size_t)n_reads * 4)) return rc;
    char *base = (char *)c->d_buf;
    if (nnz) LDPGTA_CUDA(cudaMemcpyAsync(base, allele_idx, (size_t)nnz * 4, cudaMemcpyHostToDevice, c->stream));
    if (int rc = check_idx_on_device(c, (int32_t *)base, nnz, n_alleles, true)) return rc;
    LDPGTA_CUDA(cudaMemcpyAsync(base + idx_b, read_offsets, (size_t)(n_reads + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    ScoreArgs a;
    memset(&a, 0, sizeof a);
    a.G = c->G;
    for (int g = 0; g < c->G; ++g) { a.rows[g] = c->allele_rows[g]; a.wp[g] = c->wp[g]; a.nhap[g] = c->nhap[g]; a.alpha[g] = alpha[g]; }
    a.complement = c->allele_complement;
    a.allele_idx = (const int32_t *)base;
    a.off = (const int64_t *)(base + idx_b);
    a.n_reads = n_reads;
    a.w = score_weight; a.min_HF = min_HF; a.max_HF = 1 - min_HF;
    a.out = (int32_t *)(base + idx_b + off_b);
    const unsigned blocks = (unsigned)std::min<int64_t>((n_reads * 32 + 255) / 256, (int64_t)c->sm_count * 16);
    k_read_scores<<<blocks, 256, 0, c->stream>>>(a);
    c->launches++;
    LDPGTA_CUDA(cudaGetLastError());
    LDPGTA_CUDA(cudaMemcpyAsync(out_scores, a.out, (size_t)n_reads * 4, cudaMemcpyDeviceToHost, c->stream));
    LDPGTA_CUDA(cudaStreamSynchronize(c->stream));
    if (*c->h_flag) return fail_bad_allele(allele_idx, nnz, n_alleles);
    // reads that overlap more than 12 common SNPs (long or paired reads): their 2^k combinations are spread over many warps
    std::vector<int64_t> items;
    for (int64_t r = 0; r < n_reads; ++r)
        if (out_scores[r] < 0) {
            const int k = -out_scores[r];
            if (k > SCORE_MAXQ)
                return fail(LDPGTA_E_UNSUPPORTED, "read %lld overlaps %d common SNPs: more than 2^%d haplotype combinations per read are not enumerated",
                            (long long)r, k, SCORE_MAXQ);
            for (int64_t chunk = 0; chunk < ((int64_t)1 << (k - 12)); ++chunk) { items.push_back(r); items.push_back(chunk); }
            out_scores[r] = 0;
        }
    if (!items.empty()) {
        int64_t *d_items = nullptr;
        LDPGTA_CUDA(cudaMalloc((void **)&d_items, items.size() * 8));
        cudaError_t e = cudaMemcpyAsync(d_items, items.data(), items.size() * 8, cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(a.out, out_scores, (size_t)n_reads * 4, cudaMemcpyHostToDevice, c->stream);   // zeros where the sums go
        if (e == cudaSuccess) {
            a.long_items = d_items;
            a.n_long_items = (int64_t)items.size() / 2;
            const unsigned nb = (unsigned)std::min<int64_t>((a.n_long_items * 32 + 255) / 256, (int64_t)c->sm_count * 16);
            k_read_scores_long<<<nb, 256, 0, c->stream>>>(a);
            c->launches++;
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaMemcpyAsync(out_scores, a.out, (size_t)n_reads * 4, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        cudaFree(d_items);
        if (e != cudaSuccess) return fail(LDPGTA_E_CUDA, "scoring long reads failed: %s", cudaGetErrorString(e));
    }
    return 0;
}

int ldpgta_joint_counts(ldpgta_ctx *c, const int32_t *read_idx, int n, uint32_t *out_counts)
{
    if (!c || !read_idx || !out_counts) return fail(LDPGTA_E_INVALID, "null argument");
    if (n < 2 || n > MAXN) return fail(LDPGTA_E_UNSUPPORTED, "draws of %d reads are outside the supported range 2..%d", n, MAXN);
    LDPGTA_CUDA(cudaSetDevice(c->device));
    // counts do not depend on the model; pick the kind that matches the number of groups
    double prop[MAXG];
    for (int g = 0; g < MAXG; ++g) prop[g] = 1.0 / c->G;
    const int kind = c->G == 1 ? LDPGTA_HOMOGENEOUS : LDPGTA_DISTANT_ADMIXTURE;
    if (int rc = check_kind(c, kind, prop)) return rc;
    for (int i = 0; i < n; ++i)
        if (read_idx[i] < 0 || read_idx[i] >= c->n_reads)
            return fail(LDPGTA_E_INVALID, "read index %d is outside 0..%lld", read_idx[i], (long long)c->n_reads - 1);
    const size_t cnt_bytes = ((size_t)c->G << n) * 4;
    if (int rc = ensure_dev(c, 64 + 32 + cnt_bytes)) return rc;
    int32_t *d_idx = (int32_t *)c->d_buf;
    double *d_out = (double *)((char *)c->d_buf + 64);
    uint32_t *d_cnt = (uint32_t *)((char *)c->d_buf + 96);
    LDPGTA_CUDA(cudaMemcpyAsync(d_idx, read_idx, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
    if (int rc = run_uniform(c, kind, n, d_idx, 1, nullptr, prop, d_out, d_cnt)) return rc;
    LDPGTA_CUDA(cudaMemcpyAsync(out_counts, d_cnt, cnt_bytes, cudaMemcpyDeviceToHost, c->stream));
    LDPGTA_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

int ldpgta_get_likelihoods(ldpgta_ctx *c, int kind, const int32_t *allele_idx, const int64_t *read_offsets, int n,
                           const double *proportions, int general_formula, double *out, uint32_t *counts_out)
{
    if (!c || !allele_idx || !read_offsets || !out) return fail(LDPGTA_E_INVALID, "null argument");
    if (n < 2 || n > MAXN) return fail(LDPGTA_E_UNSUPPORTED, "%d haplotypes per call are outside the supported range 2..%d", n, MAXN);
    LDPGTA_CUDA(cudaSetDevice(c->device));
    if (kind == LDPGTA_HOMOGENEOUS && c->G != 1) return fail(LDPGTA_E_INVALID, "homogeneous model needs exactly one group");
    if (kind == LDPGTA_RECENT_ADMIXTURE && c->G != 2) return fail(LDPGTA_E_INVALID, "recent-admixture model needs exactly two groups");
    if (kind == LDPGTA_DISTANT_ADMIXTURE && !proportions) return fail(LDPGTA_E_INVALID, "distant-admixture model needs ancestral proportions");
    if (kind < 0 || kind > 2) return fail(LDPGTA_E_INVALID, "unknown model kind %d", kind);
    const int64_t n_alleles = c->n_alleles[0];
    for (int g = 0; g < c->G; ++g) {
        if (!c->allele_rows[g]) return fail(LDPGTA_E_STATE, "allele rows of group %d were not uploaded", g);
        if (c->n_alleles[g] != n_alleles) return fail(LDPGTA_E_STATE, "groups hold different numbers of allele rows");
    }
    if (read_offsets[0] != 0) return fail(LDPGTA_E_INVALID, "read_offsets must start at 0");
    const int64_t nnz = read_offsets[n];
    for (int i = 0; i < n; ++i)
        if (read_offsets[i + 1] <= read_offsets[i]) return fail(LDPGTA_E_INVALID, "haplotype %d has no alleles", i);
    for (int64_t k = 0; k < nnz; ++k)
        if (allele_idx[k] < 0 || allele_idx[k] >= n_alleles)
            return fail(LDPGTA_E_INVALID, "allele index %d is outside 0..%lld", allele_idx[k], (long long)n_alleles - 1);
    // The usual call (a handful of alleles, no count table wanted): inputs and result live in one page-locked block the
    // device reads and writes directly -- one preparation launch, the draw kernel, one synchronisation, no copies.
    constexpr int IO_IDX = 1024, IO_OUT = IO_IDX * 4 + (((MAXN + 1) * 8 + 15) & ~15), IO_BYTES = IO_OUT + 32;      // the result is stored as two double2
    if (!counts_out && nnz <= IO_IDX) {
        if (!c->h_io) {
            LDPGTA_CUDA(cudaHostAlloc((void **)&c->h_io, IO_BYTES, cudaHostAllocMapped));
            LDPGTA_CUDA(cudaHostGetDevicePointer((void **)&c->d_io, c->h_io, 0));
        }
        size_t off = 0;
        auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 15) & ~(size_t)15; return o; };
        const size_t o_draw = take((size_t)n * 4);
        size_t o_rows[MAXG], o_pc[MAXG];
        for (int g = 0; g < c->G; ++g) { o_rows[g] = take((size_t)n * c->wp[g] * 8); o_pc[g] = take((size_t)n * 4); }
        if (int rc = ensure_dev(c, off)) return rc;
        char *base = (char *)c->d_buf;
        memcpy(c->h_io, allele_idx, (size_t)nnz * 4);
        memcpy(c->h_io + IO_IDX * 4, read_offsets, (size_t)(n + 1) * 8);
        SingleArgs sa;
        memset(&sa, 0, sizeof sa);
        PanelView pv = make_view(c, proportions);
        for (int g = 0; g < c->G; ++g) {
            sa.rows[g] = c->allele_rows[g]; sa.masks[g] = (uint64_t *)(base + o_rows[g]); sa.popc[g] = (uint32_t *)(base + o_pc[g]);
            sa.wp[g] = c->wp[g];
            pv.masks[g] = sa.masks[g]; pv.popc[g] = sa.popc[g];
        }
        sa.allele_idx = (const int32_t *)c->d_io; sa.off = (const int64_t *)(c->d_io + IO_IDX * 4);
        sa.draw = (int32_t *)(base + o_draw); sa.n = n;
        k_prepare_single<<<c->G, 256, 0, c->stream>>>(sa);
        c->launches++;
        LDPGTA_CUDA(cudaGetLastError());
        double *d_res = (double *)(c->d_io + IO_OUT);
        if (int rc = run_uniform(c, kind, n, sa.draw, 1, nullptr, proportions, d_res, nullptr, general_formula != 0, &pv)) return rc;
        LDPGTA_CUDA(cudaStreamSynchronize(c->stream));
        memcpy(out, c->h_io + IO_OUT, 32);
        return 0;
    }
    // scratch layout: [allele_idx][offsets][draw idx 0..n-1][out 4][counts][scratch rows per group]
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 15) & ~(size_t)15; return o; };
    const size_t o_idx = take((size_t)nnz * 4), o_off = take((size_t)(n + 1) * 8), o_draw = take((size_t)n * 4);
    const size_t o_out = take(32), o_cnt = take(((size_t)c->G << n) * 4);
    size_t o_rows[MAXG], o_pc[MAXG];
    for (int g = 0; g < c->G; ++g) { o_rows[g] = take((size_t)n * c->wp[g] * 8); o_pc[g] = take((size_t)n * 4); }
    if (int rc = ensure_dev(c, off)) return rc;
    char *base = (char *)c->d_buf;
    int32_t h_draw[MAXN];
    for (int i = 0; i < n; ++i) h_draw[i] = i;
    LDPGTA_CUDA(cudaMemcpyAsync(base + o_idx, allele_idx, (size_t)nnz * 4, cudaMemcpyHostToDevice, c->stream));
    LDPGTA_CUDA(cudaMemcpyAsync(base + o_off, read_offsets, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    LDPGTA_CUDA(cudaMemcpyAsync(base + o_draw, h_draw, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
    PanelView pv = make_view(c, proportions);
    for (int g = 0; g < c->G; ++g) {
        uint64_t *rows = (uint64_t *)(base + o_rows[g]);
        k_build_read_masks<<<1, 256, 0, c->stream>>>(c->allele_rows[g], (const int32_t *)(base + o_idx),
                                                     (const int64_t *)(base + o_off), n, c->wp[g], rows);
        k_row_popc<<<1, 256, 0, c->stream>>>(rows, n, c->wp[g], (uint32_t *)(base + o_pc[g]));
        c->launches += 2;
        pv.masks[g] = rows;
        pv.popc[g] = (const uint32_t *)(base + o_pc[g]);
    }
    LDPGTA_CUDA(cudaGetLastError());
    if (int rc = run_uniform(c, kind, n, (const int32_t *)(base + o_draw), 1, nullptr, proportions, (double *)(base + o_out),
                             counts_out ? (uint32_t *)(base + o_cnt) : nullptr, general_formula != 0, &pv))
        return rc;
    LDPGTA_CUDA(cudaMemcpyAsync(out, base + o_out, 32, cudaMemcpyDeviceToHost, c->stream));
    if (counts_out)
        LDPGTA_CUDA(cudaMemcpyAsync(counts_out, base + o_cnt, ((size_t)c->G << n) * 4, cudaMemcpyDeviceToHost, c->stream));
    LDPGTA_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

int ldpgta_likelihoods_uniform_dev(ldpgta_ctx *c, int kind, int n, const int32_t *read_idx_dev, int64_t n_draws,
                                   const double *proportions, double *out_dev)
{
    if (!c || !read_idx_dev || !out_dev || n_draws < 0) return fail(LDPGTA_E_INVALID, "bad arguments");
    LDPGTA_CUDA(cudaSetDevice(c->device));
    if (int rc = check_kind(c, kind, proportions)) return rc;
    return run_uniform(c, kind, n, read_idx_dev, n_draws, nullptr, proportions, out_dev, nullptr);
}

int ldpgta_likelihoods_uniform(ldpgta_ctx *c, int kind, int n, const int32_t *read_idx, int64_t n_draws,
                               const double *proportions, double *out)
{
    if (!c || !out || n_draws < 0) return fail(LDPGTA_E_INVALID, "bad arguments");
    LDPGTA_CUDA(cudaSetDevice(c->device));
    if (int rc = check_kind(c, kind, proportions)) return rc;
    if (int rc = check_uniform_size(kind, n)) return rc;
    if (n_draws == 0) return 0;
    if (!read_idx) return fail(LDPGTA_E_INVALID, "null read_idx");
    const bool pin_in = is_page_locked(read_idx), pin_out = is_page_locked(out);
    if (pin_in && pin_out) return run_pipelined(c, kind, n, read_idx, nullptr, n_draws, proportions, out);
    // pageable buffers go through the context's page-locked staging area
    const size_t idx_b = ((size_t)n_draws * n * 4 + 63) & ~(size_t)63, out_b = (size_t)n_draws * 32;
    if (int rc = ensure_pin(c, idx_b + out_b)) return rc;
    int32_t *h_idx = (int32_t *)c->h_pin;
    double *h_out = (double *)((char *)c->h_pin + idx_b);
    const int32_t *src = read_idx;
    if (!pin_in) { memcpy(h_idx, read_idx, (size_t)n_draws * n * 4); src = h_idx; }
    const int rc = run_pipelined(c, kind, n, src, nullptr, n_draws, proportions, pin_out ? out : h_out);
    if (rc) return rc;
    if (!pin_out) memcpy(out, h_out, out_b);
    return 0;
}

int ldpgta_bootstrap_uniform(ldpgta_ctx *c, int kind, int n, const int64_t *window_offsets, const int32_t *window_reads,
                             int64_t n_windows, const int64_t *draw_offsets, uint64_t seed, const double *proportions,
                             double *out, int32_t *idx_out)
{
    if (!c || !window_offsets || !draw_offsets || !out || n_windows < 0) return fail(LDPGTA_E_INVALID, "bad arguments");
    LDPGTA_CUDA(cudaSetDevice(c->device));
    if (int rc = check_kind(c, kind, proportions)) return rc;
    if (int rc = check_uniform_size(kind, n)) return rc;
    if (window_offsets[0] != 0 || draw_offsets[0] != 0) return fail(LDPGTA_E_INVALID, "offsets must start at 0");
    const int64_t n_draws = draw_offsets[n_windows], n_wreads = window_offsets[n_windows];
    if (n_draws == 0) return 0;
    if (!window_reads) return fail(LDPGTA_E_INVALID, "null window_reads");
    for (int64_t w = 0; w < n_windows; ++w) {
        const int64_t R = window_offsets[w + 1] - window_offsets[w], nd = draw_offsets[w + 1] - draw_offsets[w];
        if (nd < 0 || R < 0) return fail(LDPGTA_E_INVALID, "offsets must not decrease (window %lld)", (long long)w);
        if (nd > 0 && R < n) return fail(LDPGTA_E_INVALID, "window %lld has %lld reads, fewer than the %d of a draw", (long long)w, (long long)R, n);
        if (R > 0x7fffffff) return fail(LDPGTA_E_INVALID, "window %lld has too many reads", (long long)w);
    }

    auto pad = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t woff_b = pad((size_t)(n_windows + 1) * 8), wr_b = pad((size_t)n_wreads * 4);
    const size_t idx_b = pad((size_t)n_draws * n * 4), out_b = (size_t)n_draws * 32;
    if (int rc = ensure_dev(c, 2 * woff_b + wr_b + idx_b + out_b)) return rc;
    char *p = (char *)c->d_buf;
    int64_t *d_woff = (int64_t *)p; p += woff_b;
    int64_t *d_doff = (int64_t *)p; p += woff_b;
    int32_t *d_wreads = (int32_t *)p; p += wr_b;
    int32_t *d_idx = (int32_t *)p; p += idx_b;
    double *d_out = (double *)p;
    if (int rc = ensure_pipeline(c, 9)) return rc;
    LDPGTA_CUDA(cudaMemcpyAsync(d_woff, window_offsets, (size_t)(n_windows + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    LDPGTA_CUDA(cudaMemcpyAsync(d_doff, draw_offsets, (size_t)(n_windows + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    LDPGTA_CUDA(cudaMemcpyAsync(d_wreads, window_reads, (size_t)n_wreads * 4, cudaMemcpyHostToDevice, c->stream));
    if (int rc = check_idx_on_device(c, d_wreads, n_wreads, c->n_reads, true)) return rc;
    {
        SampleArgs sa;
        memset(&sa, 0, sizeof sa);
        sa.draw_off = d_doff; sa.win_off = d_woff; sa.win_reads = d_wreads; sa.n_windows = n_windows; sa.n_draws = n_draws;
        sa.seed = seed; sa.idx_out = d_idx;
        launch_sample(c, n, sa);
    }
    LDPGTA_CUDA(cudaGetLastError());
    // kernel and download in chunks: the copy of chunk i runs under the kernel of chunk i + 1
    const bool pin_out = is_page_locked(out);
    double *h_out = out;
    if (!pin_out) {
        if (int rc = ensure_pin(c, out_b)) return rc;
        h_out = (double *)c->h_pin;
    }
    const int64_t n_chunks = std::max<int64_t>(1, std::min<int64_t>(8, n_draws / 32768));
    const int64_t per = ((n_draws + n_chunks - 1) / n_chunks + 63) & ~(int64_t)63;
    int ci = 0;
    for (int64_t d0 = 0; d0 < n_draws; d0 += per, ++ci) {
        const int64_t nd = std::min(per, n_draws - d0);
        if (int rc = run_uniform(c, kind, n, d_idx + d0 * n, nd, nullptr, proportions, d_out + 4 * d0, nullptr)) return rc;
        LDPGTA_CUDA(cudaEventRecord(c->ev_k[ci], c->stream));
        LDPGTA_CUDA(cudaStreamWaitEvent(c->s_out, c->ev_k[ci], 0));
        LDPGTA_CUDA(cudaMemcpyAsync(h_out + 4 * d0, d_out + 4 * d0, (size_t)nd * 32, cudaMemcpyDeviceToHost, c->s_out));
    }
    if (idx_out) LDPGTA_CUDA(cudaMemcpyAsync(idx_out, d_idx, (size_t)n_draws * n * 4, cudaMemcpyDeviceToHost, c->stream));
    LDPGTA_CUDA(cudaStreamSynchronize(c->stream));
    LDPGTA_CUDA(cudaStreamSynchronize(c->s_out));
    if (*c->h_flag) return fail_bad_index(c, window_reads, n_wreads);
    if (!pin_out) memcpy(out, h_out, out_b);
    return 0;
}

int ldpgta_likelihoods_batch(ldpgta_ctx *c, int kind, const int32_t *read_idx, const int64_t *draw_offsets, int64_t n_draws,
                             const double *proportions, double *out)
{
    if (!c || !draw_offsets || !out || n_draws < 0) return fail(LDPGTA_E_INVALID, "bad arguments");
    LDPGTA_CUDA(cudaSetDevice(c->device));
    if (int rc = check_kind(c, kind, proportions)) return rc;
    if (n_draws == 0) return 0;
    if (!read_idx) return fail(LDPGTA_E_INVALID, "null read_idx");
    if (draw_offsets[0] != 0) return fail(LDPGTA_E_INVALID, "draw_offsets must start at 0");

    // page-locked caller buffers holding draws of one size: the pipelined path, nothing scanned up front
    {
        const int64_t n0w = draw_offsets[1] - draw_offsets[0];
        if (n_draws >= 65536 && draw_offsets[n_draws] == n_draws * n0w && n0w >= 2 && n0w <= MAXN &&
            is_page_locked(read_idx) && is_page_locked(out)) {
            if (int rc = check_uniform_size(kind, n0w)) return rc;
            const int rc = run_pipelined(c, kind, (int)n0w, read_idx, draw_offsets, n_draws, proportions, out);
            if (rc <= 0) return rc;                  // 1: sizes differ after all, regroup below
        }
    }

    // size classes (branch-free first pass: the common batch has one draw size)
    int64_t per_n[MAXN + 1] = {};
    const int64_t n0w = draw_offsets[1] - draw_offsets[0];
    int64_t differs = 0;
    for (int64_t d = 0; d < n_draws; ++d) differs |= (draw_offsets[d + 1] - draw_offsets[d]) ^ n0w;
    const bool uniform = differs == 0;
    const int n0 = (int)n0w;
    if (uniform) {
        if (int rc = check_uniform_size(kind, n0w)) return rc;
        per_n[n0] = n_draws;
    } else {
        for (int64_t d = 0; d < n_draws; ++d) {
            const int64_t n = draw_offsets[d + 1] - draw_offsets[d];
            if (n < 2 || n > MAXN)
                return fail(LDPGTA_E_UNSUPPORTED, "draw %lld has %lld reads, supported range is 2..%d", (long long)d, (long long)n, MAXN);
            if (kind == LDPGTA_DISTANT_ADMIXTURE && n == 5)
                return fail(LDPGTA_E_UNSUPPORTED, "distant_admixture has no likelihoods5 (reference: DISTANT_ADMIXTURE_MODELS.py:312-313 raises AttributeError)");
            per_n[n]++;
        }
    }
    const int64_t nnz = draw_offsets[n_draws];
    {
        const uint32_t lim = (uint32_t)std::min<int64_t>(c->n_reads, 0x7fffffff);
        uint32_t bad = 0;
        for (int64_t k = 0; k < nnz; ++k) bad |= (uint32_t)((uint32_t)read_idx[k] >= lim);     // negative indices wrap above lim
        if (bad) return fail_bad_index(c, read_idx, nnz);
    }

    const size_t idx_bytes = ((size_t)nnz * 4 + 15) & ~(size_t)15;
    const size_t pos_bytes = uniform ? 0 : (size_t)n_draws * 8;
    const size_t out_bytes = (size_t)n_draws * 32;
    if (int rc = ensure_dev(c, idx_bytes + pos_bytes + out_bytes)) return rc;
    int32_t *d_idx = (int32_t *)c->d_buf;
    int64_t *d_pos = (int64_t *)((char *)c->d_buf + idx_bytes);
    double *d_out = (double *)((char *)c->d_buf + idx_bytes + pos_bytes);

    if (uniform) {
        const void *src = read_idx;
        if (!is_page_locked(read_idx)) {
            if (int rc = ensure_pin(c, std::max(idx_bytes, out_bytes))) return rc;
            memcpy(c->h_pin, read_idx, (size_t)nnz * 4);
            src = c->h_pin;
        }
        LDPGTA_CUDA(cudaMemcpyAsync(d_idx, src, (size_t)nnz * 4, cudaMemcpyHostToDevice, c->stream));
        if (int rc = run_uniform(c, kind, n0, d_idx, n_draws, nullptr, proportions, d_out, nullptr)) return rc;
    } else {
        // regroup the draws by size; outputs are scattered back to their original positions
        if (int rc = ensure_pin(c, idx_bytes + pos_bytes)) return rc;
        int32_t *h_idx = (int32_t *)c->h_pin;
        int64_t *h_pos = (int64_t *)((char *)c->h_pin + idx_bytes);
        int64_t idx_start[MAXN + 2] = {}, pos_start[MAXN + 2] = {};
        for (int n = 2; n <= MAXN; ++n) {
            idx_start[n + 1] = idx_start[n] + per_n[n] * n;
            pos_start[n + 1] = pos_start[n] + per_n[n];
        }
        int64_t fill[MAXN + 1] = {};
        for (int64_t d = 0; d < n_draws; ++d) {
            const int n = (int)(draw_offsets[d + 1] - draw_offsets[d]);
            const int64_t k = fill[n]++;
            memcpy(h_idx + idx_start[n] + k * n, read_idx + draw_offsets[d], (size_t)n * 4);
            h_pos[pos_start[n] + k] = d;
        }
        LDPGTA_CUDA(cudaMemcpyAsync(d_idx, h_idx, (size_t)nnz * 4, cudaMemcpyHostToDevice, c->stream));
        LDPGTA_CUDA(cudaMemcpyAsync(d_pos, h_pos, pos_bytes, cudaMemcpyHostToDevice, c->stream));
        for (int n = 2; n <= MAXN; ++n)
            if (per_n[n])
                if (int rc = run_uniform(c, kind, n, d_idx + idx_start[n], per_n[n], d_pos + pos_start[n], proportions, d_out, nullptr))
                    return rc;
    }
    if (is_page_locked(out)) {
        LDPGTA_CUDA(cudaMemcpyAsync(out, d_out, out_bytes, cudaMemcpyDeviceToHost, c->stream));
        LDPGTA_CUDA(cudaStreamSynchronize(c->stream));
    } else {
        LDPGTA_CUDA(cudaStreamSynchronize(c->stream));     // pinned staging may still feed the H2D copy
        if (int rc = ensure_pin(c, out_bytes)) return rc;
        LDPGTA_CUDA(cudaMemcpyAsync(c->h_pin, d_out, out_bytes, cudaMemcpyDeviceToHost, c->stream));
        LDPGTA_CUDA(cudaStreamSynchronize(c->stream));
        memcpy(out, c->h_pin, out_bytes);
    }
    return 0;
}

int ldpgta_bin_stats(ldpgta_ctx *c, const double *x, int n_series, const int64_t *window_offsets, int64_t n_windows,
                     const int64_t *bin_ranges, int64_t n_bins, double *out)
{
    if (!c || !window_offsets || !bin_ranges || !out || n_series < 1 || n_windows < 1 || n_bins < 0)
        return fail(LDPGTA_E_INVALID, "bad arguments");
    LDPGTA_CUDA(cudaSetDevice(c->device));
    if (window_offsets[0] != 0) return fail(LDPGTA_E_INVALID, "window_offsets must start at 0");
    const int64_t n_draws = window_offsets[n_windows];
    for (int64_t w = 0; w < n_windows; ++w)
        if (window_offsets[w + 1] - window_offsets[w] < 1) return fail(LDPGTA_E_INVALID, "window %lld has no draws", (long long)w);
    for (int64_t b = 0; b < n_bins; ++b)
        if (bin_ranges[2 * b] < 0 || bin_ranges[2 * b + 1] > n_windows || bin_ranges[2 * b] >= bin_ranges[2 * b + 1])
            return fail(LDPGTA_E_INVALID, "bin %lld: empty or out-of-range window range [%lld, %lld)", (long long)b,
                        (long long)bin_ranges[2 * b], (long long)bin_ranges[2 * b + 1]);
    if (n_bins == 0) return 0;
    if (n_series > 65535) return fail(LDPGTA_E_INVALID, "too many series");
    const size_t x_b = x ? (size_t)n_series * n_draws * 8 : 0, off_b = ((size_t)(n_windows + 1) * 8 + 15) & ~(size_t)15;
    const size_t rng_b = (size_t)n_bins * 16, out_b = (size_t)n_series * n_bins * 16;
    const double *d_x;
    const int64_t *d_off;
    if (int rc = ensure_dev(c, x_b + off_b + rng_b + out_b + 64)) return rc;
    char *p = (char *)c->d_buf;
    if (x) {
        LDPGTA_CUDA(cudaMemcpyAsync(p, x, x_b, cudaMemcpyHostToDevice, c->stream));
        d_x = (const double *)p; p += x_b;
        LDPGTA_CUDA(cudaMemcpyAsync(p, window_offsets, (size_t)(n_windows + 1) * 8, cudaMemcpyHostToDevice, c->stream));
        d_off = (const int64_t *)p; p += off_b;
    } else {
        if (n_series != 5 || c->llr_draws != n_draws || c->llr_windows != n_windows)
            return fail(LDPGTA_E_INVALID, "no resident LLRs for these windows: call ldpgta_window_stats first (5 series)");
        if (c->llr_lazy_lik && n_draws > 0) {
            // the statistics pass did not need the LLR matrix; form it now from the likelihoods it left on the device
            const unsigned blocks = (unsigned)std::min<int64_t>((n_draws + 255) / 256, (int64_t)c->sm_count * 8);
            k_draw_llr<<<blocks, 256, 0, c->stream>>>(c->llr_lazy_lik, n_draws, const_cast<double *>(c->llr_ptr));
            c->launches++;
            c->llr_lazy_lik = nullptr;
        }
        d_x = c->llr_ptr;
        d_off = c->llr_off_ptr;
    }
    int64_t *d_rng = (int64_t *)p; p += rng_b;
    double *d_out = (double *)p;
    LDPGTA_CUDA(cudaMemcpyAsync(d_rng, bin_ranges, rng_b, cudaMemcpyHostToDevice, c->stream));
    k_bin_stats<<<dim3((unsigned)n_bins, (unsigned)n_series), 128, 0, c->stream>>>(d_x, n_draws, d_off, d_rng, n_bins, d_out);
    c->launches++;
    LDPGTA_CUDA(cudaGetLastError());
    LDPGTA_CUDA(cudaMemcpyAsync(out, d_out, out_b, cudaMemcpyDeviceToHost, c->stream));
    LDPGTA_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

int ldpgta_finish_chromosome(const double *cp, int n_cols, int first_window_draws, double *out)
{
    if (!cp || !out || n_cols < 0 || first_window_draws < 1) return fail(LDPGTA_E_INVALID, "bad arguments");
    const int ncp = n_cols + 3;
    for (int p = 0; p < 5; ++p) {
        const double *q = cp + (size_t)p * ncp;
        const double M = q[1], N = (double)first_window_draws;
        const double mu = q[0] / N;                                   // sum_w (sum_i x_wi) / N   (:72)
        double ss = 0.0;
        for (int i = 0; i < n_cols; ++i) { const double d = q[3 + i] - mu; ss += d * d; }   // (:73)
        out[3 * p + 0] = mu / M;                                      // (:75)
        out[3 * p + 1] = sqrt(ss / (N - 1.0)) / M;                    // (:74); N = 1 divides by zero as the reference does
        out[3 * p + 2] = q[2] / M;                                    // fraction_of_negative_LLRs (:313)
    }
    return 0;
}

int ldpgta_host_alloc(void **ptr, int64_t bytes)
{
    if (!ptr || bytes < 0) return fail(LDPGTA_E_INVALID, "bad arguments");
    LDPGTA_CUDA(cudaMallocHost(ptr, (size_t)std::max<int64_t>(bytes, 16)));
    return 0;
}

int ldpgta_host_free(void *ptr)
{
    if (ptr) LDPGTA_CUDA(cudaFreeHost(ptr));
    return 0;
}

}  // extern "C"
