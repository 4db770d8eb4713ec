// common.cuh -- context, error plumbing and small device helpers shared by the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>

#include "../../include/ldpgta.h"

namespace ldpgta {

constexpr int MAXG = LDPGTA_MAX_GROUPS;
constexpr int MAXN = LDPGTA_MAX_READS_PER_DRAW;

extern thread_local std::string g_last_error;

inline int fail(int code, const char *fmt, ...) __attribute__((format(printf, 2, 3)));

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
    int sm_count = 148;
    int max_smem_optin = 0;

    uint64_t *allele_rows[ldpgta::MAXG] = {};
    int64_t n_alleles[ldpgta::MAXG] = {};
    uint64_t *read_masks[ldpgta::MAXG] = {};
    bool masks_owned[ldpgta::MAXG] = {};
    int64_t n_reads_g[ldpgta::MAXG] = {};
    int64_t n_reads = 0;

    // growable staging buffers
    void *h_pin = nullptr;  size_t h_pin_bytes = 0;
    void *d_buf = nullptr;  size_t d_buf_bytes = 0;

    int64_t launches = 0;
};
