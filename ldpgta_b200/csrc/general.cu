// general.cu -- kernels, shared-memory planning and launcher for draws of 5..16 reads (k_general.cuh).
#include <algorithm>
#include <stdlib.h>
#include <vector>

#include "common.cuh"
#include "k_general.cuh"

namespace ldpgta {

struct GeneralPlan { int th, tl, nw, kl, kh, wc; bool u16, global_table, etab; size_t smem, table_bytes; };

static int plan_general(const ldpgta_ctx *c, int n, int kind, GeneralPlan *p)
{
    static const int TH[19] = {0, 0, 0, 0, 0, 1, 2, 4, 8, 2, 4, 4, 8, 8, 8, 8, 8, 8, 8};
    static const int TL[19] = {0, 0, 0, 0, 0, 1, 1, 1, 1, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4};
    static const int NW[19] = {0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 4, 4, 8, 16, 16, 16, 16, 16};
    static const int BUDGET_KB[19] = {0, 0, 16, 16, 16, 16, 16, 14, 14, 24, 24, 48, 56, 110, 220, 220, 220, 220, 220};   // measured (tools/tune_general.sh)
    if (n < 5) { p->th = 1; p->tl = 1; p->nw = 1; p->kl = n; }       // likelihoods() forced on 2..4 reads
    else { p->th = TH[n]; p->tl = TL[n]; p->nw = NW[n]; p->kl = p->tl == 1 ? 5 : 7; }
    if (const char *e = getenv("LDPGTA_PLAN_NW")) p->nw = atoi(e);          // tuning overrides
    if (const char *e = getenv("LDPGTA_PLAN_TH")) p->th = atoi(e);
    p->kh = n - p->kl;
    uint32_t nmax = 0; int wmax = 0, wsum = 0;
    for (int g = 0; g < c->G; ++g) { nmax = std::max(nmax, c->nhap[g]); wmax = std::max(wmax, c->wp[g]); wsum += c->wp[g]; }
    // joint-frequency tables in shared memory up to 128 KB, otherwise one slice per resident CTA in global memory (L2-resident)
    const size_t entries = (size_t)c->G << n;
    // 16-bit entries whenever the counts fit (groups of at most 65535 haplotypes): the polynomial, which waits on its table
    // rows, reads half the bytes (measured at 13-14 reads: 0.83 -> 0.93, 0.77 -> 0.93 of the roofline; 12-13 reads of
    // two-group models +10 %).  LDPGTA_PLAN_U32=1 restores 32-bit entries below 64 KB for comparison.
    static const bool force_u32 = getenv("LDPGTA_PLAN_U32") != nullptr;
    p->u16 = nmax <= 65535u && n >= 5 && (!force_u32 || entries * 4 > 64 * 1024);
    const size_t ct = p->u16 ? 2 : 4;
    p->table_bytes = entries * ct;
    p->global_table = p->table_bytes > 128 * 1024;
    // distant admixture: e(S) as doubles next to the count tables while both fit (n <= 13 for two groups)
    // (three or more groups only: two groups use the integer form, poly_general.cuh Cubic2)
    p->etab = kind == LDPGTA_DISTANT_ADMIXTURE && c->G > 2 && n >= 5 && !p->global_table &&
              p->table_bytes + ((size_t)8 << n) + (size_t)n * wsum * 8 + 16 * 1024 <= (size_t)c->max_smem_optin;
    const size_t fixed = 128 + (size_t)n * wsum * 8 + (p->global_table ? 0 : p->table_bytes) + (p->etab ? (size_t)8 << n : 0) + 64 + 16 + 16 * 22 * 8;
    const size_t per_word = ((size_t)8 << p->kl) + ((size_t)8 << p->kh);
    size_t budget_kb = (size_t)BUDGET_KB[n];
    // distant admixture with e(S) in a global slice: its polynomial re-reads 128-byte rows through L1, which is what the
    // shared-memory carve-out leaves of 256 KB.  192 KB of chunk tables instead of 220 KB: 16 reads 11.3 -> 9.9 ms, 17 reads
    // 16.8 -> 13.4 ms per 592 draws (the integer models lose 1-2 % with it and keep 220 KB)
    if (kind == LDPGTA_DISTANT_ADMIXTURE && c->G > 2 && n >= 5 && !p->etab) budget_kb = std::min<size_t>(budget_kb, 192);
    if (const char *e = getenv("LDPGTA_PLAN_KB")) budget_kb = (size_t)atoi(e);
    const size_t cap = std::min<size_t>((size_t)c->max_smem_optin, budget_kb * 1024 + (p->etab ? (size_t)8 << n : 0));
    if (fixed + 2 * per_word > (size_t)c->max_smem_optin)
        return fail(LDPGTA_E_UNSUPPORTED, "draws of %d reads with %d group(s) do not fit shared memory (%zu B fixed)", n, c->G, fixed);
    size_t avail = cap > fixed ? cap - fixed : 0;
    int wc = (int)(avail / per_word);
    wc = std::max(wc, 2);
    wc = std::min(wc, wmax);
    // the counting loop folds three words into two POPC.64 (carry-save), so a chunk is a multiple of three words wherever
    // the table is built in several chunks: e.g. 79 words as 15 x 5 + 4 cost 53 POPC.64 per subset, as 14 x 5 + 9 cost 56
    if (wc < wmax && wc >= 3 && !getenv("LDPGTA_PLAN_NO_WC3")) {
        wc -= wc % 3;
        // the fewest multiples of three that still give the same number of chunks
        const int chunks = (wmax + wc - 1) / wc;
        int even = (wmax + chunks - 1) / chunks;
        even += (3 - even % 3) % 3;
        if (even <= wc) wc = even;
    } else {
        const int chunks = (wmax + wc - 1) / wc;             // equalise the chunks: ceil(wmax / ceil(wmax / wc))
        wc = (wmax + chunks - 1) / chunks;
    }
    p->wc = wc;
    PanelView pv = make_view(c, nullptr);
    p->smem = p->u16 ? general_smem_layout<uint16_t>(pv, n, p->kl, p->kh, wc, !p->global_table, p->etab).total
                     : general_smem_layout<uint32_t>(pv, n, p->kl, p->kh, wc, !p->global_table, p->etab).total;
    if (p->smem > (size_t)c->max_smem_optin)
        return fail(LDPGTA_E_UNSUPPORTED, "draws of %d reads need %zu B of shared memory", n, p->smem);
    return 0;
}

template <int TH, int TL, typename CT>
static int launch_general_t(ldpgta_ctx *c, GeneralArgs a, const GeneralPlan &p)
{
    auto kern = k_general<TH, TL, CT>;
    LDPGTA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
    int per_sm = 0;
    LDPGTA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32 * p.nw, p.smem));
    if (per_sm < 1) return fail(LDPGTA_E_UNSUPPORTED, "general kernel does not fit an SM (n=%d, smem=%zu)", a.n, p.smem);
    const int64_t resident = (int64_t)per_sm * c->sm_count;
    const unsigned blocks = (unsigned)std::min<int64_t>(a.n_draws, resident);
    a.table_scratch = nullptr; a.table_stride = 0;
    if (p.global_table) {
        const size_t stride = (p.table_bytes + 255) & ~(size_t)255, need = stride * blocks;
        if (c->table_scratch_bytes < need) {
            LDPGTA_CUDA(cudaStreamSynchronize(c->stream));
            if (c->table_scratch) LDPGTA_CUDA(cudaFree(c->table_scratch));
            c->table_scratch = nullptr; c->table_scratch_bytes = 0;
            LDPGTA_CUDA(cudaMalloc((void **)&c->table_scratch, need));
            c->table_scratch_bytes = need;
        }
        a.table_scratch = c->table_scratch;
        a.table_stride = stride;
    }
    a.etab_scratch = nullptr; a.etab_stride = 0;
    if (a.kind == LDPGTA_DISTANT_ADMIXTURE && a.pv.G > 2 && a.n >= 5 && !p.etab) {
        // e(S) as doubles in a global slice per resident CTA (L2 / HBM): still far cheaper than forming it at every use
        const size_t stride = (size_t)1 << a.n, need = stride * 8 * blocks;
        if (c->etab_scratch_bytes < need) {
            LDPGTA_CUDA(cudaStreamSynchronize(c->stream));
            if (c->etab_scratch) LDPGTA_CUDA(cudaFree(c->etab_scratch));
            c->etab_scratch = nullptr; c->etab_scratch_bytes = 0;
            LDPGTA_CUDA(cudaMalloc((void **)&c->etab_scratch, need));
            c->etab_scratch_bytes = need;
        }
        a.etab = 2;
        a.etab_scratch = c->etab_scratch;
        a.etab_stride = stride;
    }
    kern<<<blocks, 32 * p.nw, p.smem, c->stream>>>(a);
    c->launches++;
    LDPGTA_CUDA(cudaGetLastError());
    return 0;
}

// Colourings (a_h, b_h, c_h) of the h = n - 5 high reads (every read in exactly one part) with b_h < c_h,
// packed a | b << 16 | c << 32.  (3^h - 1) / 2 entries, generated once per h and kept on the device.
int ensure_hi_terms(ldpgta_ctx *c, int h)
{
    if (c->hi_terms[h] || h == 0) return 0;
    std::vector<uint64_t> list;
    const uint32_t fullh = (1u << h) - 1u;
    for (uint32_t a = 0; a <= fullh; ++a) {
        const uint32_t rest = fullh ^ a;
        if (!rest) continue;
        const uint32_t hi = 1u << (31 - __builtin_clz(rest));       // c owns the highest free read: b < c
        const uint32_t sub = rest ^ hi;
        uint32_t b = sub;
        while (true) {
            list.push_back((uint64_t)a | ((uint64_t)b << 16) | ((uint64_t)(rest ^ b) << 32));
            if (!b) break;
            b = (b - 1) & sub;
        }
    }
    LDPGTA_CUDA(cudaMalloc((void **)&c->hi_terms[h], list.size() * 8));
    LDPGTA_CUDA(cudaMemcpyAsync(c->hi_terms[h], list.data(), list.size() * 8, cudaMemcpyHostToDevice, c->stream));
    LDPGTA_CUDA(cudaStreamSynchronize(c->stream));
    c->n_hi_terms[h] = (int64_t)list.size();
    return 0;
}

// The same colourings as runs: a | b_base << 16 | l << 32 stands for b_h = b_base | x over the subsets x of l, where l holds
// the `bits` highest reads outside a_h below the very highest one (which belongs to c_h), or all of them when there are fewer.
int ensure_run_terms(ldpgta_ctx *c, int h, int bits)
{
    if (c->run_terms[h][bits]) return 0;
    std::vector<uint64_t> list;
    const uint32_t fullh = (1u << h) - 1u;
    for (uint32_t a = 0; a < fullh; ++a) {
        const uint32_t rest = fullh ^ a, hi = 1u << (31 - __builtin_clz(rest)), sub = rest ^ hi;
        // the HIGHEST reads of sub go into l: consecutive runs (the lanes of a warp) then differ in the lowest reads outside
        // a_h and their rows spread over the shared-memory banks; with the lowest reads in l every lane of a quarter warp hit
        // the same bank group (16 reads: 3.30 instead of 2.71 ms).  Dealing the runs into groups of eight with every row
        // residue mod 4 twice (2 wavefronts per load, the floor for 32-byte rows, instead of 3.5) gained 1.5 % at 16 reads
        // and lost 4-5 % at 17-18 reads, where the tables are in L2 and neighbouring runs share sectors: not kept.
        uint32_t l = 0, left = sub;
        for (int k = 0; k < bits && left; ++k) { const uint32_t top = 1u << (31 - __builtin_clz(left)); l |= top; left ^= top; }
        const uint32_t upper = sub ^ l;
        uint32_t b = upper;
        while (true) {
            list.push_back((uint64_t)a | ((uint64_t)b << 16) | ((uint64_t)l << 32));
            if (!b) break;
            b = (b - 1) & upper;
        }
    }
    LDPGTA_CUDA(cudaMalloc((void **)&c->run_terms[h][bits], list.size() * 8));
    LDPGTA_CUDA(cudaMemcpyAsync(c->run_terms[h][bits], list.data(), list.size() * 8, cudaMemcpyHostToDevice, c->stream));
    LDPGTA_CUDA(cudaStreamSynchronize(c->stream));
    c->n_run_terms[h][bits] = (int64_t)list.size();
    return 0;
}

int launch_general(ldpgta_ctx *c, GeneralArgs a)
{
    GeneralPlan p;
    if (int rc = plan_general(c, a.n, a.kind, &p)) return rc;
    a.kl = p.kl; a.kh = p.kh; a.wc = p.wc; a.etab = p.etab ? 1 : 0;
    a.terms = nullptr; a.n_terms = 0; a.wide = 0;
    if (a.n >= 5) {
        const int h = a.n - 5;
        if (int rc = ensure_hi_terms(c, h)) return rc;
        a.terms = reinterpret_cast<const u64 *>(c->hi_terms[h]);
        a.n_terms = c->n_hi_terms[h];
        for (int g = 0; g < c->G; ++g) a.wide |= c->nhap[g] > (c->G == 2 ? 11585u : 16383u);   // 32-bit convolution sums: 16 (two groups: 32) products of two counts
    }
    // integer models from 13 reads (enough runs to balance the block): the colourings come in runs that share the X row, as
    // long as the 32-bit sums hold 2^bits convolutions -- 16 (recent admixture: 32) products of two counts each.  Measured
    // against single colourings (LDPGTA_PLAN_NO_RUNS=1), 5008 haplotypes: 13..18 reads +3, +4, +7, +10, +15, +16 % draws/s.
    a.runs = 0;
    static const bool no_runs = getenv("LDPGTA_PLAN_NO_RUNS") != nullptr;
    if (a.n >= 13 && (a.kind != LDPGTA_DISTANT_ADMIXTURE || c->G == 2) && c->G <= 2 && !a.wide && !no_runs) {
        uint64_t nmax = 0;
        for (int g = 0; g < c->G; ++g) nmax = std::max<uint64_t>(nmax, c->nhap[g]);
        const uint64_t per_conv = (c->G == 2 ? 32u : 16u) * nmax * nmax;
        int bits = 0;
        const int want = a.n >= 15 ? 3 : 2;                      // runs of 4 keep the few colourings of 13-14 reads balanced
        while (bits < want && (per_conv << (bits + 1)) <= 0xffffffffull) ++bits;
        if (const char *e = getenv("LDPGTA_PLAN_RUN_BITS")) bits = std::min(bits, atoi(e));
        if (bits >= 1) {
            const int h = a.n - 5;
            if (int rc = ensure_run_terms(c, h, bits)) return rc;
            a.terms = reinterpret_cast<const u64 *>(c->run_terms[h][bits]);
            a.n_terms = c->n_run_terms[h][bits];
            a.runs = 1;
        }
    }
#define LDPGTA_CASE(TH_, TL_)                                                         \
    if (p.th == TH_ && p.tl == TL_)                                                   \
        return p.u16 ? launch_general_t<TH_, TL_, uint16_t>(c, a, p) : launch_general_t<TH_, TL_, uint32_t>(c, a, p);
    LDPGTA_CASE(1, 1) LDPGTA_CASE(2, 1) LDPGTA_CASE(4, 1) LDPGTA_CASE(8, 1) LDPGTA_CASE(2, 4) LDPGTA_CASE(4, 4) LDPGTA_CASE(8, 4)
#undef LDPGTA_CASE
    return fail(LDPGTA_E_UNSUPPORTED, "no kernel instance for n=%d", a.n);
}

}  // namespace ldpgta
