// tools/int_peaks.cu -- measures the integer-pipe peaks the likelihood kernels are bounded by:
// POPC.32, LOP3.32, IADD3 and the AND+POPC+ADD "word-op" mix, per SM per clock and per second.
// Writes one JSON object to stdout (bench.py stores it as profiles/INT_PEAKS.json).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/int_peaks tools/int_peaks.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

constexpr int ITERS = 4096;
constexpr int CHAINS = 8;
static double g_clock_khz = 1965000.0;

template <int MODE>
__global__ void __launch_bounds__(256) k_peak(uint32_t *out, long long *cycles, uint32_t seed)
{
    uint32_t r[CHAINS], a = seed ^ threadIdx.x, b = seed * 2654435761u + blockIdx.x;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) r[i] = seed + i * 0x9e3779b9u + threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) {
            if (MODE == 0) {            // POPC.32 only
                asm volatile("popc.b32 %0, %0;" : "+r"(r[i]));
            } else if (MODE == 1) {     // LOP3.32 only
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i]) : "r"(a), "r"(b));
            } else if (MODE == 2) {     // IADD3
                asm volatile("add.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(a));
            } else if (MODE == 3) {     // 1 POPC : 1 LOP3
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;\n\tpopc.b32 %0, %0;" : "+r"(r[i]) : "r"(a), "r"(b));
            } else if (MODE == 4) {     // 1 POPC : 3 LOP3
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;\n\tlop3.b32 %0, %0, %2, %1, 0xe8;\n\tlop3.b32 %0, %0, %1, %2, 0x96;\n\tpopc.b32 %0, %0;"
                             : "+r"(r[i]) : "r"(a), "r"(b));
            } else if (MODE == 5) {     // 1 POPC : 1 IMAD (fma pipe)
                asm volatile("mad.lo.u32 %0, %0, %1, %2;\n\tpopc.b32 %0, %0;" : "+r"(r[i]) : "r"(a), "r"(b));
            } else if (MODE == 6) {     // IMAD.32 only
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(a), "r"(b));
            }
        }
    }
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s ^= r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// the kernels' inner product: acc += popcll(h & l) on 64-bit words held in registers
__global__ void __launch_bounds__(256) k_wordop(uint32_t *out, long long *cycles, uint64_t seed)
{
    uint64_t h[4], l[8];
    uint32_t acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = seed * (i + 3) + threadIdx.x;
#pragma unroll
    for (int j = 0; j < 8; ++j) l[j] = ~seed * (j + 5) + blockIdx.x;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0;
    long long t0 = clock64();
    for (int it = 0; it < ITERS / 4; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] += __popcll(h[i] & l[j]);
#pragma unroll
        for (int j = 0; j < 8; ++j) l[j] = l[j] * 3 + it;     // keep the loop from collapsing
    }
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) s += acc[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// the multiply-adds of the partition polynomials: 64-bit integer a*b+c (k_general's exact class sums) and float64 fma
// (distant admixture), and float64 log (the LLRs of statistics())
template <int MODE>
__global__ void __launch_bounds__(256) k_peak64(double *out, long long *cycles, uint64_t seed)
{
    uint64_t r[CHAINS];
    double d[CHAINS];
    const uint64_t a = seed | 1ull, b = seed * 0x9e3779b97f4a7c15ull + threadIdx.x;
    const double fa = 1.0000001, fb = 1e-9 * (double)(threadIdx.x + 1);
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) { r[i] = seed + i * 0x9e3779b9u + threadIdx.x; d[i] = 1.0 + 1e-3 * (double)(i + threadIdx.x); }
    long long t0 = clock64();
    for (int it = 0; it < ITERS / 4; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) {
            if (MODE == 0) asm volatile("mad.lo.u64 %0, %0, %1, %2;" : "+l"(r[i]) : "l"(a), "l"(b));
            else if (MODE == 1) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(fa), "d"(fb));
            else d[i] = log(d[i]) + 2.5;
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += d[i] + (double)r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <typename F>
static int time_it(F launch, int blocks, double ops_per_thread, const char *name, int sms, long long *d_cyc, bool last)
{
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int w = 0; w < 3; ++w) launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    long long cyc = 0;
    CK(cudaMemcpy(&cyc, d_cyc, sizeof cyc, cudaMemcpyDeviceToHost));
    const double total = ops_per_thread * 256.0 * blocks;
    const double per_s = total / (best * 1e-3);
    // per clock at the maximum SM clock (the device may run a little below it under load)
    const double per_clk_sm = per_s / sms / (g_clock_khz * 1e3);
    printf("  \"%s\": {\"ops_per_s\": %.4e, \"ops_per_clk_per_sm_at_max_clock\": %.2f, \"ms\": %.4f, \"block0_cycles\": %lld}%s\n",
           name, per_s, per_clk_sm, best, cyc, last ? "" : ",");
    return 0;
}

int main()
{
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    g_clock_khz = p.clockRate;
    const int sms = p.multiProcessorCount, blocks = sms * 8;       // 8 x 256 threads = 64 warps per SM
    uint32_t *d_out; long long *d_cyc;
    CK(cudaMalloc(&d_out, (size_t)blocks * 256 * 4));
    CK(cudaMalloc(&d_cyc, (size_t)blocks * 8));
    printf("{\n  \"device\": \"%s\", \"sms\": %d, \"clock_khz_max\": %d,\n", p.name, sms, p.clockRate);
    const double n = (double)ITERS * CHAINS;
    time_it([&] { k_peak<0><<<blocks, 256>>>(d_out, d_cyc, 1u); }, blocks, n, "popc32", sms, d_cyc, false);
    time_it([&] { k_peak<1><<<blocks, 256>>>(d_out, d_cyc, 2u); }, blocks, n, "lop3_32", sms, d_cyc, false);
    time_it([&] { k_peak<2><<<blocks, 256>>>(d_out, d_cyc, 3u); }, blocks, n, "iadd32", sms, d_cyc, false);
    time_it([&] { k_peak<3><<<blocks, 256>>>(d_out, d_cyc, 4u); }, blocks, n, "popc_plus_1lop3_pairs", sms, d_cyc, false);
    time_it([&] { k_peak<4><<<blocks, 256>>>(d_out, d_cyc, 5u); }, blocks, n, "popc_plus_3lop3_groups", sms, d_cyc, false);
    time_it([&] { k_peak<5><<<blocks, 256>>>(d_out, d_cyc, 6u); }, blocks, n, "popc_plus_1imad_pairs", sms, d_cyc, false);
    time_it([&] { k_peak<6><<<blocks, 256>>>(d_out, d_cyc, 8u); }, blocks, n, "imad32", sms, d_cyc, false);
    double *d_out64;
    CK(cudaMalloc(&d_out64, (size_t)blocks * 256 * 8));
    const double n4 = (double)(ITERS / 4) * CHAINS;
    time_it([&] { k_peak64<0><<<blocks, 256>>>(d_out64, d_cyc, 9ull); }, blocks, n4, "imad64_lo", sms, d_cyc, false);
    time_it([&] { k_peak64<1><<<blocks, 256>>>(d_out64, d_cyc, 10ull); }, blocks, n4, "dfma", sms, d_cyc, false);
    time_it([&] { k_peak64<2><<<blocks, 256>>>(d_out64, d_cyc, 11ull); }, blocks, n4, "log_f64", sms, d_cyc, false);
    time_it([&] { k_wordop<<<blocks, 256>>>(d_out, d_cyc, 7ull); }, blocks, (double)(ITERS / 4) * 32, "wordop64_and_popc_add", sms, d_cyc, true);
    printf("}\n");
    return 0;
}
