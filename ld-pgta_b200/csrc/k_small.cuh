// k_small.cuh -- draws of 2, 3 or 4 reads (the CLI default -M 4): everything in registers.
//
// A group of LPD lanes owns one draw.  Each lane loads its share of the NR read masks
// straight from HBM/L2 with 128-bit loads (lane l of the group takes 16-byte slot
// j*LPD + l of the row, so a group reads 16*LPD contiguous bytes per instruction),
// forms every non-empty subset intersection with one AND per subset (parent AND one
// read, joint_frequencies_combo HOMOGENOUES_MODELS.py:129-163), popcounts, and the
// 2^NR-1 partial counts are summed across the group's lanes with xor-shuffles.  The
// closed-form polynomials (poly_closed.cuh) run in the same kernel in float64.
//
// Algorithmic work per draw: G*(2^NR-1)*W word AND+POPC, G*NR*W*8 mask bytes in, 32 B out.
#pragma once
#include "common.cuh"
#include "poly_closed.cuh"

namespace ldpgta {

struct SmallArgs {
    PanelView pv;
    int kind;
    const int32_t *idx;      // [n_draws][NR]
    int64_t n_draws;
    const int64_t *pos;      // optional scatter positions of the outputs
    double *out;             // [*][4]
    uint32_t *counts_out;    // optional [n_draws][G][2^NR]
};

template <int NR, int LPD, int WPL>
__global__ void __launch_bounds__(128) k_small(const SmallArgs a)
{
    static_assert(WPL % 2 == 0, "rows are read in 16-byte slots");
    constexpr int NS = 1 << NR;
    constexpr int SLOTS = WPL / 2;
    const int lane_in = threadIdx.x % LPD;
    int64_t draw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LPD;
    const bool live = draw < a.n_draws;
    if (!live) draw = a.n_draws - 1;        // keep the lanes in the shuffles

    int32_t r[NR];
#pragma unroll
    for (int i = 0; i < NR; ++i) r[i] = __ldg(a.idx + draw * NR + i);

    double f[NS], g2[NS];                   // normalised frequencies (g2: second group of recent admixture)
#pragma unroll
    for (int s = 0; s < NS; ++s) { f[s] = 0.0; g2[s] = 0.0; }

    for (int g = 0; g < a.pv.G; ++g) {
        const int wp = a.pv.wp[g];
        const uint64_t *base = a.pv.masks[g];
        const uint64_t *row[NR];
#pragma unroll
        for (int i = 0; i < NR; ++i) row[i] = base + (int64_t)r[i] * wp;

        uint32_t c[NS];
#pragma unroll
        for (int s = 0; s < NS; ++s) c[s] = 0;

        for (int w0 = 0; w0 < wp; w0 += LPD * WPL) {
            uint4 m[NR][SLOTS];
#pragma unroll
            for (int j = 0; j < SLOTS; ++j) {
                const int w = w0 + (j * LPD + lane_in) * 2;
                const bool in = w < wp;
#pragma unroll
                for (int i = 0; i < NR; ++i) m[i][j] = in ? ldg128(row[i] + w) : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int j = 0; j < SLOTS; ++j) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    u64 x[NS];
#pragma unroll
                    for (int s = 1; s < NS; ++s) {
                        const int top = 31 - __clz(s);          // compile-time after unrolling
                        const int rest = s ^ (1 << top);
                        const uint4 v = m[top][j];
                        const u64 word = h ? (((u64)v.w << 32) | v.z) : (((u64)v.y << 32) | v.x);
                        x[s] = rest ? (x[rest] & word) : word;
                        c[s] += __popcll(x[s]);
                    }
                }
            }
        }
        // sum the partial counts over the LPD lanes of the draw
        if (a.pv.nhap[g] <= 65535u) {       // two 16-bit counters per shuffle
#pragma unroll
            for (int s = 0; s < NS; s += 2) {
                uint32_t v = c[s] | (c[s + 1] << 16);
#pragma unroll
                for (int msk = LPD / 2; msk; msk >>= 1) v += __shfl_xor_sync(0xffffffffu, v, msk);
                c[s] = v & 0xffffu;
                c[s + 1] = v >> 16;
            }
        } else {
#pragma unroll
            for (int s = 1; s < NS; ++s)
#pragma unroll
                for (int msk = LPD / 2; msk; msk >>= 1) c[s] += __shfl_xor_sync(0xffffffffu, c[s], msk);
        }
        if (a.counts_out && live && lane_in == 0) {
            uint32_t *o = a.counts_out + (draw * a.pv.G + g) * NS;
            o[0] = 0;
#pragma unroll
            for (int s = 1; s < NS; ++s) o[s] = c[s];
        }
        const double nh = (double)a.pv.nhap[g];
        if (a.kind == LDPGTA_DISTANT_ADMIXTURE) {
            // effective_joint_frequency, DISTANT_ADMIXTURE_MODELS.py:136-139: sum_g p_g * cnt / N_g
            const double p = a.pv.coef[g];
#pragma unroll
            for (int s = 1; s < NS; ++s) f[s] = f[s] + p * (double)c[s] / nh;
        } else if (g == 0) {
#pragma unroll
            for (int s = 1; s < NS; ++s) f[s] = (double)c[s] / nh;
        } else {
#pragma unroll
            for (int s = 1; s < NS; ++s) g2[s] = (double)c[s] / nh;
        }
    }

    double res[4];
    if (a.kind == LDPGTA_RECENT_ADMIXTURE) closed_recent<NR>(f, g2, res);
    else closed_single<NR>(f, res);

    if (live && lane_in == 0) {
        const int64_t o = a.pos ? a.pos[draw] : draw;
        double2 *dst = reinterpret_cast<double2 *>(a.out + 4 * o);
        dst[0] = make_double2(res[0], res[1]);
        dst[1] = make_double2(res[2], res[3]);
    }
}

}  // namespace ldpgta
