// common.cuh -- context, error plumbing and small device helpers shared by the kernels.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>

#include "../../include/ldpgta.h"

namespace ldpgta {
struct WindowPlan;
}

namespace ldpgta {

constexpr int MAXG = LDPGTA_MAX_GROUPS;
constexpr int MAXN = LDPGTA_MAX_READS_PER_DRAW;

extern thread_local std::string g_last_error;

int fail(int code, const char *fmt, ...) __attribute__((format(printf, 2, 3)));

#define LDPGTA_CUDA(expr)                                                                      \
    do {                                                                                       \
        cudaError_t e__ = (expr);                                                              \
        if (e__ != cudaSuccess)                                                                \
            return ::ldpgta::fail(LDPGTA_E_CUDA, "%s failed: %s (%s:%d)", #expr,               \
                                  cudaGetErrorString(e__), __FILE__, __LINE__);                \
    } while (0)

// Device-side view of the packed read masks of every group (passed by value to kernels).
struct PanelView {
    const uint64_t *masks[MAXG];  // [n_reads][wp[g]] 64-bit words, pad bits/words zero
    const uint32_t *popc[MAXG];   // [n_reads] popcount of each row (the singleton joint frequencies)
    const double *norm[MAXG];     // [nhap+1] x / nhap for x = 0..nhap, IEEE division done once (exact normalisation by lookup)
    int wp[MAXG];                 // padded row length in words (even: rows are 16-byte aligned)
    uint32_t nhap[MAXG];          // haplotypes per group
    double coef[MAXG];            // distant admixture: ancestral proportion of the group
    int G;
};

typedef unsigned long long u64;
typedef unsigned __int128 u128;

__device__ __forceinline__ uint4 ldg128(const uint64_t *p)
{
    return __ldg(reinterpret_cast<const uint4 *>(p));
}

__device__ __forceinline__ u64 shfl_xor_u64(u64 v, int m)
{
    uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
    lo = __shfl_xor_sync(0xffffffffu, lo, m);
    hi = __shfl_xor_sync(0xffffffffu, hi, m);
    return ((u64)hi << 32) | lo;
}

__device__ __forceinline__ u64 warp_sum_u64(u64 v)
{
#pragma unroll
    for (int m = 16; m; m >>= 1) v += shfl_xor_u64(v, m);
    return v;
}

__device__ __forceinline__ double warp_sum_f64(double v)
{
#pragma unroll
    for (int m = 16; m; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    return v;
}

// ---- carry-save popcount of N 64-bit words --------------------------------------------
// POPC issues at 16 lanes/clk/SM, LOP3 at 64 (profiles/INT_PEAKS.json), and the two pipes
// run concurrently.  A carry-save adder (2 LOP3 per 32-bit half) turns three words of weight
// w into one of weight w and one of weight 2w, so popc(x0)+...+popc(x9) needs 5 POPC.64
// instead of 10.  Recursion: groups of three at each weight until at most two words remain.
__device__ __forceinline__ void csa(u64 a, u64 b, u64 c, u64 &sum, u64 &carry)
{
    const u64 ab = a ^ b;
    sum = ab ^ c;
    carry = (a & b) | (ab & c);
}

// The same with the two LOP3s spelled out.  When a, b, c are themselves ANDs, the compiler folds each AND into the
// xor chain ((a0 & a1) ^ x is one LOP3) and then has to form the ANDs again for the majority: 6 to 6.75 LOP3 per 32-bit
// half of a triple instead of 5 (SASS of k_general's main loop).  Opaque lop3 keeps it at 3 ANDs + xor3 + maj.
__device__ __forceinline__ uint32_t lop3_xor3(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t lop3_maj(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xe8;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
// popc(ones) + 2 * popc(twos) of the carry-save sum of three 64-bit words
__device__ __forceinline__ uint32_t csa_count(u64 a, u64 b, u64 c)
{
    const uint32_t al = (uint32_t)a, ah = (uint32_t)(a >> 32), bl = (uint32_t)b, bh = (uint32_t)(b >> 32);
    const uint32_t cl = (uint32_t)c, ch = (uint32_t)(c >> 32);
    const uint32_t ones = __popc(lop3_xor3(al, bl, cl)) + __popc(lop3_xor3(ah, bh, ch));
    const uint32_t twos = __popc(lop3_maj(al, bl, cl)) + __popc(lop3_maj(ah, bh, ch));
    return ones + 2 * twos;
}

// acc += (popc(ones) + 2 * popc(twos)) << SHIFT for the carry-save sum of three 64-bit words.  The doubled terms are
// integer multiply-adds (FMA pipe) so that only one 3-input add per triple lands on the ALU pipe, which also carries
// all the LOP3s.
template <int SHIFT>
__device__ __forceinline__ void csa_accumulate(uint32_t &acc, u64 a, u64 b, u64 c)
{
    const uint32_t al = (uint32_t)a, ah = (uint32_t)(a >> 32), bl = (uint32_t)b, bh = (uint32_t)(b >> 32);
    const uint32_t cl = (uint32_t)c, ch = (uint32_t)(c >> 32);
    const uint32_t o_lo = __popc(lop3_xor3(al, bl, cl)), o_hi = __popc(lop3_xor3(ah, bh, ch));
    const uint32_t t_lo = __popc(lop3_maj(al, bl, cl)), t_hi = __popc(lop3_maj(ah, bh, ch));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(acc) : "r"(t_lo), "n"(2 << SHIFT));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(acc) : "r"(t_hi), "n"(2 << SHIFT));
    if (SHIFT == 0) acc += o_lo + o_hi;
    else {
        asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(acc) : "r"(o_lo), "n"(1 << SHIFT));
        asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(acc) : "r"(o_hi), "n"(1 << SHIFT));
    }
}

template <int N>
struct PopcSum {
    static __device__ __forceinline__ uint32_t run(const u64 *x)
    {
        constexpr int K = N / 3, REM = N % 3;
        u64 ones[K + REM], twos[K];
#pragma unroll
        for (int k = 0; k < K; ++k) csa(x[3 * k], x[3 * k + 1], x[3 * k + 2], ones[k], twos[k]);
#pragma unroll
        for (int r = 0; r < REM; ++r) ones[K + r] = x[3 * K + r];
        return PopcSum<K + REM>::run(ones) + 2u * PopcSum<K>::run(twos);
    }
};
template <> struct PopcSum<0> { static __device__ __forceinline__ uint32_t run(const u64 *) { return 0u; } };
template <> struct PopcSum<1> { static __device__ __forceinline__ uint32_t run(const u64 *x) { return __popcll(x[0]); } };
template <> struct PopcSum<2> { static __device__ __forceinline__ uint32_t run(const u64 *x) { return __popcll(x[0]) + __popcll(x[1]); } };

// the same on 32-bit words (no 64-bit register pairing constraints for the allocator).  DEPTH limits the
// number of carry-save levels: each level trades POPC (XU pipe, 16/clk/SM) for LOP3 (ALU pipe, 64/clk/SM);
// the depth that balances the two pipes is chosen per kernel.
__device__ __forceinline__ void csa32(uint32_t a, uint32_t b, uint32_t c, uint32_t &sum, uint32_t &carry)
{
    sum = lop3_xor3(a, b, c);              // one LOP3 (0x96)
    carry = lop3_maj(a, b, c);             // one LOP3 (0xe8, majority); opaque so that the ANDs feeding a, b, c are formed once
}

template <int N, int DEPTH>
struct PopcSum32 {
    static __device__ __forceinline__ uint32_t run(const uint32_t *x)
    {
        if constexpr (DEPTH == 0 || N < 3) {
            uint32_t s = 0;
#pragma unroll
            for (int i = 0; i < N; ++i) s += __popc(x[i]);
            return s;
        } else {
            constexpr int K = N / 3, REM = N % 3;
            uint32_t ones[K + REM + 1], twos[K + 1];
#pragma unroll
            for (int k = 0; k < K; ++k) csa32(x[3 * k], x[3 * k + 1], x[3 * k + 2], ones[k], twos[k]);
#pragma unroll
            for (int r = 0; r < REM; ++r) ones[K + r] = x[3 * K + r];
            return PopcSum32<K + REM, (DEPTH > 0 ? DEPTH - 1 : 0)>::run(ones) + 2u * PopcSum32<K, (DEPTH > 0 ? DEPTH - 1 : 0)>::run(twos);
        }
    }
};

// coefficients of log_ratio as constant-bank operands (immediates would cost two moves per multiply-add): 1/3 .. 1/21,
// ln 2 low and high parts, sqrt 2
static __device__ __constant__ double c_logr[13] = {1.0 / 3.0, 1.0 / 5.0, 1.0 / 7.0, 1.0 / 9.0, 1.0 / 11.0, 1.0 / 13.0, 1.0 / 15.0, 1.0 / 17.0,
                                                    1.0 / 19.0, 1.0 / 21.0, 1.90821492927058770002e-10, 6.93147180369123816490e-01,
                                                    1.41421356237309504880};

// log(y / x) for positive normal y, x without forming the quotient.  With y = my 2^ey, x = mx 2^ex (my, mx in [1, 2), one
// of them rescaled so that my / mx lies in [1/sqrt 2, sqrt 2]):
//     log(y / x) = k ln 2 + 2 atanh(t),   t = (my - mx) / (my + mx),   |t| <= 0.1716
// my - mx is exact (Sterbenz), so ONE division serves both the quotient and the atanh argument, and the result keeps its
// relative accuracy when y ~ x, where log(fl(y / x)) -- what CPython's math.log(y / x) of ANEUPLOIDY_TEST.py:86 returns --
// carries the 2^-53 rounding of the quotient.  Checked on the host against a long-double reference over 2e7 pairs (equal,
// near, far apart, at the range-reduction boundaries): at most 1.99 ulp from the true value, and never further from
// log(fl(y / x)) than that value's own error bound (1 ulp + 2^-53).  About 50 instructions against ~160 for a library
// division followed by a library log.  All multiply-adds are explicit (the library is built with -fmad=false).
__device__ __forceinline__ double log_ratio(double y, double x)
{
    const int hy = __double2hiint(y), hx = __double2hiint(x);
    int k = (hy >> 20) - (hx >> 20);
    double my = __hiloint2double((hy & 0x000fffff) | 0x3ff00000, __double2loint(y));
    const double mx = __hiloint2double((hx & 0x000fffff) | 0x3ff00000, __double2loint(x));
    if (my > mx * c_logr[12]) { my *= 0.5; k += 1; }
    else if (my * c_logr[12] < mx) { my *= 2.0; k -= 1; }
    const double d = my - mx, s = my + mx;                    // s in [2, 4): no special cases in the reciprocal
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(s));      // 20 bits
    double e = fma(-s, r, 1.0);
    r = fma(r, e, r);
    e = fma(-s, r, 1.0);
    r = fma(r, e, r);
    double t = d * r;
    t = fma(fma(-s, t, d), r, t);                              // one correction step: t = d / s to within an ulp
    const double z = t * t;
    double q = c_logr[9];                                      // atanh(t) = t (1 + z/3 + z^2/5 + ... + z^10/21), z <= 0.0295
#pragma unroll
    for (int i = 8; i >= 0; --i) q = fma(q, z, c_logr[i]);
    const double t2 = t + t, kd = (double)k;
    const double small = fma(t2 * z, q, kd * c_logr[10]);      // ln 2 = hi + lo, hi with 32 trailing zero bits
    return fma(kd, c_logr[11], t2 + small);
}

// LLR(y, x), ANEUPLOIDY_TEST.py:78-90: log(y / x) with the sentinels for vanishing likelihoods
__device__ __forceinline__ double llr_of(double y, double x)
{
    if (x != 0.0 && y != 0.0) {
#ifndef LDPGTA_LIBRARY_LOG
        // both positive, normal and finite (exponent fields 1..2046): the fast form; anything else takes the library's
        const unsigned ey = ((unsigned)__double2hiint(y) >> 20) - 1u, ex = ((unsigned)__double2hiint(x) >> 20) - 1u;
        if (ey < 2046u && ex < 2046u) return log_ratio(y, x);
#endif
        return log(y / x);
    }
    if (x != 0.0) return -1.23456789;
    if (y != 0.0) return +1.23456789;
    return 0.0;
}

// the five LLRs of statistics() (:298) of one draw, res = (monosomy, disomy, SPH, BPH), into llr[5][stride] at column o:
// (BPH,SPH) (disomy,monosomy) (BPH,disomy) (disomy,SPH) (SPH,monosomy).  The draw kernels call this where they store the
// likelihoods, so the reductions start from LLRs that never had to be re-read.
__device__ __forceinline__ void store_llrs(double *__restrict__ llr, int64_t stride, int64_t o, const double *res)
{
    llr[o] = llr_of(res[3], res[2]);
    llr[stride + o] = llr_of(res[1], res[0]);
    llr[2 * stride + o] = llr_of(res[3], res[1]);
    llr[3 * stride + o] = llr_of(res[1], res[2]);
    llr[4 * stride + o] = llr_of(res[2], res[0]);
}

__device__ __forceinline__ double u128_to_double(u128 v)
{
    return (double)(u64)(v >> 64) * 18446744073709551616.0 + (double)(u64)v;
}

}  // namespace ldpgta

// Host-side context (opaque to callers).
struct ldpgta_ctx {
    int device = 0;
    int G = 0;
    uint32_t nhap[ldpgta::MAXG] = {};
    int words[ldpgta::MAXG] = {};   // ceil(nhap/64)
    int wp[ldpgta::MAXG] = {};      // padded to an even number of words
    cudaStream_t stream = nullptr;
    cudaStream_t s_in = nullptr, s_out = nullptr;     // upload / download streams of the pipelined host path
    std::vector<cudaEvent_t> ev_in, ev_k;
    int sm_count = 148;
    int max_smem_optin = 0;

    uint8_t *allele_complement = nullptr;     // per allele row: 1 = stored row is the complement of the panel row
    uint64_t *allele_rows[ldpgta::MAXG] = {};
    int64_t n_alleles[ldpgta::MAXG] = {};
    uint64_t *read_masks[ldpgta::MAXG] = {};
    uint32_t *read_popc[ldpgta::MAXG] = {};   // popcount of every read mask
    double *norm[ldpgta::MAXG] = {};          // x / nhap lookup per group
    // 2-D tensor maps over the read masks ([n_reads][wp] uint64, box = one row) for TMA gather4 (k_small, 4 reads)
    alignas(64) CUtensorMap tmap[ldpgta::MAXG];
    bool tmap_ok[ldpgta::MAXG] = {};
    // two-group panels: both groups' masks of a read in one 80-word row (k_small_both), built on first use and dropped when
    // either group's masks change
    uint64_t *masks_both = nullptr;
    size_t masks_both_bytes = 0;
    bool both_valid = false, both_tmap_ok = false;
    uint64_t masks_epoch = 0;                 // bumped whenever a group's read masks change (copies derived from them are rebuilt)
    alignas(64) CUtensorMap tmap_both;
    bool masks_owned[ldpgta::MAXG] = {};
    int64_t n_reads_g[ldpgta::MAXG] = {};
    int64_t n_reads = 0;

    // growable staging buffers
    void *h_pin = nullptr;  size_t h_pin_bytes = 0;
    void *d_buf = nullptr;  size_t d_buf_bytes = 0;
    // LLRs of the last ldpgta_window_stats call, kept on the device for ldpgta_bin_stats (own allocation)
    double *d_llr = nullptr;  size_t d_llr_bytes = 0;
    int64_t llr_draws = -1, llr_windows = -1;
    const double *llr_ptr = nullptr;          // the resident LLRs ldpgta_bin_stats(x = NULL) reads: d_llr or the plan's buffer
    const int64_t *llr_off_ptr = nullptr;
    const double *llr_lazy_lik = nullptr;     // resident likelihoods [llr_draws][4] from which llr_ptr still has to be formed (k_draw_llr), or nullptr
    int64_t *d_llr_off = nullptr;  size_t d_llr_off_bytes = 0;
    unsigned int *d_flag = nullptr, *h_flag = nullptr;     // device-side index validation of the pipelined host path
    // the pipelined host path as a CUDA graph, replayed while the call signature repeats
    cudaGraphExec_t pipe_exec = nullptr;
    uint64_t pipe_key = 0, pipe_seen = 0, pipe_bad = 0;
    int64_t pipe_launches = 0, pipe_replays = 0;
    cudaEvent_t ev_fork = nullptr, ev_join_in = nullptr, ev_join_out = nullptr;

    // colourings of the high reads used by the three-block polynomial, one list per h = n - 5 (built on first use)
    uint64_t *hi_terms[14] = {};
    int64_t n_hi_terms[14] = {};
    uint64_t *run_terms[14][4] = {};       // the same colourings as runs of 2^bits that share a_h (poly_general.cuh), bits 1..3
    int64_t n_run_terms[14][4] = {};
    // joint-frequency tables that do not fit shared memory (17-18 reads, or two-group models at 16): one slice per resident CTA
    unsigned char *table_scratch = nullptr;
    size_t table_scratch_bytes = 0;
    double *etab_scratch = nullptr;           // e(S) tables of the distant model when they do not fit shared memory
    size_t etab_scratch_bytes = 0;

    // windows + bootstrap plan made resident by ldpgta_set_windows (bootstrap_api.cu)
    ldpgta::WindowPlan *plan = nullptr;
    void *comm = nullptr;                     // peer-memory communicator (comm_api.cu)
    // page-locked, device-mapped block for the inputs and the result of single get_likelihoods calls
    unsigned char *h_io = nullptr, *d_io = nullptr;
    // optional CUDA-event brackets around the draw kernels of the resident pipeline (ldpgta_kernel_timing)
    int time_kernels = 0;            // 1: events around the draw kernels, 2: around the reductions instead (ldpgta_kernel_timing)
    std::vector<cudaEvent_t> t_ev;      // pairs (start, stop), reused
    size_t t_used = 0;
    int t_stride = 1;                // only every t_stride-th bracket is recorded (an event pair costs the stream ~5 us)
    uint64_t t_marks = 0;

    int64_t launches = 0;
};

namespace ldpgta {

inline PanelView make_view(const ldpgta_ctx *c, const double *proportions)
{
    PanelView pv;
    memset(&pv, 0, sizeof pv);
    pv.G = c->G;
    for (int g = 0; g < c->G; ++g) {
        pv.masks[g] = c->read_masks[g];
        pv.popc[g] = c->read_popc[g];
        pv.norm[g] = c->norm[g];
        pv.wp[g] = c->wp[g];
        pv.nhap[g] = c->nhap[g];
        pv.coef[g] = proportions ? proportions[g] : 0.0;
    }
    return pv;
}


struct SmallArgs;
struct GeneralArgs;
struct SampleArgs;
struct WindowPlan;

// helpers of ldpgta_api.cu shared with bootstrap_api.cu
int ensure_dev(ldpgta_ctx *c, size_t bytes);
int ensure_pin(ldpgta_ctx *c, size_t bytes);
int ensure_pipeline(ldpgta_ctx *c, int n_chunks);
bool is_page_locked(const void *p);
int check_kind(const ldpgta_ctx *c, int kind, const double *proportions);
int check_uniform_size(int kind, int64_t n);
int fail_bad_index(const ldpgta_ctx *c, const int32_t *read_idx, int64_t nnz);
// draws of one size, device-resident indices and outputs (dispatch to k_small / k_mid / k_lanes / k_general)
int run_uniform(ldpgta_ctx *c, int kind, int n, const int32_t *idx_dev, int64_t n_draws, const int64_t *pos_dev,
                const double *proportions, double *out_dev, uint32_t *counts_dev, bool force_general = false,
                const PanelView *override_view = nullptr, double *llr_dev = nullptr, int64_t llr_stride = 0);
int launch_sample(ldpgta_ctx *c, int n, const SampleArgs &a);      // pick_reads on the device, n reads per draw
void free_window_plan(ldpgta_ctx *c);
void free_comm(ldpgta_ctx *c);

template <class T>
inline int ensure_own(T **ptr, size_t *have, size_t bytes)
{
    if (*have >= bytes) return 0;
    if (*ptr) LDPGTA_CUDA(cudaFree(*ptr));
    *ptr = nullptr; *have = 0;
    bytes = bytes + bytes / 4 + 4096;
    LDPGTA_CUDA(cudaMalloc((void **)ptr, bytes));
    *have = bytes;
    return 0;
}
// kernel launchers, one translation unit each (small.cu, general.cu)
int launch_small(ldpgta_ctx *c, SmallArgs &a, int n, int kind);
int launch_small_windows(ldpgta_ctx *c, const int32_t *d_idx, int64_t n_draws, double *d_out, const int64_t *d_woff, const int64_t *d_doff,
                         int64_t n_windows, int64_t max_rows, const uint64_t *masks = nullptr, const uint32_t *popc = nullptr, bool local = false);
bool small_windows_ok(const ldpgta_ctx *c, int64_t max_rows);
int make_read_mask_tmap(ldpgta_ctx *c, int g);
int launch_mid(ldpgta_ctx *c, SmallArgs &a, int n);      // 5-6 reads, register kernel; 1 = not applicable
int launch_general(ldpgta_ctx *c, GeneralArgs a);
int launch_lanes(ldpgta_ctx *c, GeneralArgs a);          // 7-10 reads (5-6 for two-group models), register counters; 1 = not applicable
int ensure_run_terms(ldpgta_ctx *c, int h, int bits);    // ... as runs of up to 2^bits colourings with a common a_h
int ensure_hi_terms(ldpgta_ctx *c, int h);               // colouring list of the three-block polynomial for h = n - 5 high reads

}  // namespace ldpgta
