// Production min-sum BP kernel for code-capacity check matrices (BB codes) on sm_100a:
// T LANES PER SHOT, float32 messages resident in shared memory.
//
// Same arithmetic as bp_decode_kernel<float, VAR_MIN_SUM, ..> (bp_kernel.cuh; reference
// rework/decoding.py:5-75), different mapping.  The thread-per-shot kernel is bound by shared-memory
// CAPACITY: [[144,12,12]] needs 2.3 KB of message state per shot, so only 96 shots -- 3 warps -- fit
// on an SM and every dependent instruction is exposed (ncu r1a: 19 % issue utilisation, 43 % stall_wait,
// 33 % short_scoreboard).  Here each shot is decoded by T = 4 or 8 lanes of one warp, so the same
// shared memory feeds T times as many warps:
//   * check pass : lane j of a shot owns checks c = T*i + j (min1 / min2 / sign parity of its RW
//                  incoming messages, written as a 2-word summary),
//   * variable pass: lane j owns variables v = T*i + j (rebuilds the <= 3 check-to-variable messages
//                  from the summaries and the unmodified Q, posterior, hard decision, damped + clipped
//                  Q update),
//   * __syncwarp between the passes; the syndrome of the hard decision is accumulated per lane from
//     packed columns of H and XOR-reduced over the T lanes with shuffles.
// Layout: message rows are numbered ELL-style r = k*m + c (k-th edge of check c) and stored
// [row][slot] with S slots per row, S = G (mod 32), G = 32/T shots per warp, lane = j*G + g.  The G
// lanes of one j read G consecutive words, and rows whose index differs mod T start G banks apart,
// so the check pass (rows k*m + T*i + j, j = 0..T-1) is bank-conflict free and its addresses are
// affine; the variable pass reads rows through a table and is conflict-free whenever the T rows it
// touches differ mod T.
// The kernel is persistent; retired shots are replaced from a global cursor (refill is done by all T
// lanes of the shot and only when at least `refill_min` shots of the warp are idle, or nothing is
// left to wait for).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "bp_kernel.cuh"

namespace qldpc {

struct BPTiledLayout {
    size_t off_vt0, off_vt1, off_colmask, off_prior, off_state;
    size_t tables, per_slot;
    int npad;
};
__host__ __device__ inline BPTiledLayout bp_tiled_layout(const BPGraphDev &g, int T)
{
    BPTiledLayout L;
    L.npad = (g.n + T - 1) / T * T;
    size_t o = 0;
    L.off_vt0 = o;     o += 8 * (size_t)L.npad * 3;
    L.off_vt1 = g.two_tables ? o : L.off_vt0;
    if (g.two_tables) o += 8 * (size_t)L.npad * 3;
    L.off_colmask = o; o += 4 * (size_t)L.npad * g.WM;
    L.off_prior = o;   o += 4 * (size_t)L.npad;
    o = (o + 127) & ~(size_t)127;
    L.off_state = o;
    L.tables = o;
    // rows: RW*m message rows (4 B per slot) + m summary rows (float2 {signed min1, min2}: 8 B per slot)
    // + WN hard-decision rows (4 B per slot)
    L.per_slot = 4 * ((size_t)g.uniform_row_w * g.m + 2 * (size_t)g.m + g.WN);
    return L;
}

// vell tables (global, built by the host): entry (v, t), t < 3: c | (k << 16), or 0xFFFFFFFF when
// variable v has fewer than t+1 edges.  c = check, k = position of v inside row c.
template <int T, int WMS, int RW>
__global__ void __launch_bounds__(576, 1)
bp_tiled_kernel(const BPParams P, const uint32_t *__restrict__ vell0, const uint32_t *__restrict__ vell1, int refill_min)
{
    constexpr int G = 32 / T;                 // shots per warp
    constexpr int VPW = 32 / T;               // variables (or checks) per lane per 32-bit word
    const BPGraphDev &g = P.g;
    const int m = g.m, n = g.n, WN = g.WN;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int j = lane / G, gi = lane % G;
    const int S = (blockDim.x >> 5) * G;      // slots per row
    const int slot = warp * G + gi;
    const unsigned FULL = 0xffffffffu;

    extern __shared__ __align__(16) unsigned char smem[];
    const BPTiledLayout L = bp_tiled_layout(g, T);
    uint2 *vt0 = reinterpret_cast<uint2 *>(smem + L.off_vt0);
    uint2 *vt1 = reinterpret_cast<uint2 *>(smem + L.off_vt1);
    uint32_t *colmask = reinterpret_cast<uint32_t *>(smem + L.off_colmask);
    float *prior = reinterpret_cast<float *>(smem + L.off_prior);
    unsigned char *state = smem + L.off_state;
    const uint32_t rowB = 4u * S;                               // bytes per row
    const uint32_t offM = (uint32_t)RW * m * rowB, offHW = offM + 2u * (uint32_t)m * rowB;   // summaries: 2*rowB per check

    for (int i = threadIdx.x; i < L.npad * 3; i += blockDim.x) {
        const int v = i / 3;
        uint32_t e0 = (v < n) ? vell0[i] : 0xffffffffu, e1 = (v < n) ? vell1[i] : 0xffffffffu;
        auto mk = [&](uint32_t e) {
            if (e == 0xffffffffu) return make_uint2(0xffffffffu, 0u);
            const uint32_t c = e & 0xffffu, k = e >> 16;
            return make_uint2((k * m + c) * rowB, offM + c * 2u * rowB);
        };
        vt0[i] = mk(e0);
        if (g.two_tables) vt1[i] = mk(e1);
    }
    for (int i = threadIdx.x; i < L.npad * WMS; i += blockDim.x) colmask[i] = (i < n * WMS) ? g.colmask[i] : 0u;
    for (int i = threadIdx.x; i < L.npad; i += blockDim.x) prior[i] = (i < n) ? reinterpret_cast<const float *>(P.prior)[i] : 0.f;
    __syncthreads();

    unsigned char *my = state + 4 * slot;                       // this shot's column in the 4-byte rows
    unsigned char *my8 = state + 8 * slot;                      // ... and in the 8-byte summary rows
    const float alpha = (float)P.alpha, damp = (float)P.damping, omd = (float)P.one_minus_damping, clipv = (float)P.clip;
    const int max_iter = P.max_iter;

    uint32_t synd[WMS], acc[WMS];
    long long shot = -1;
    int iter = 0;
    bool active = false, exhausted = false;
    unsigned long long iter_sum = 0;

    while (true) {
        // ---- refill ------------------------------------------------------------------------
        const unsigned idle = __ballot_sync(FULL, !active && !exhausted);       // lanes (all T of a shot together)
        const unsigned busy = __ballot_sync(FULL, active);
        const int idle_shots = __popc(idle) / T;
        if (idle_shots > 0 && (idle_shots >= refill_min || busy == 0)) {
            // one atomic per warp; shot ids are dealt to the idle groups in lane order of sub-lane 0
            const unsigned lead = idle & ((1u << G) - 1u);                      // idle groups, seen at j == 0
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(P.cursor, (unsigned long long)__popc(lead));
            base = __shfl_sync(FULL, base, 0);
            if (!active && !exhausted) {
                shot = (long long)(base + __popc(lead & ((1u << gi) - 1u)));
                if (shot < P.B) {
                    const uint32_t *sp = P.synd + (size_t)shot * WMS;
#pragma unroll
                    for (int w = 0; w < WMS; ++w) synd[w] = sp[w];
                    // Q = where(mask, prior, 0) (decoding.py:21): the T lanes split the variables
                    for (int v = j; v < n; v += T) {
                        const float pv = prior[v] + 0.f;
#pragma unroll
                        for (int t = 0; t < 3; ++t) {
                            *reinterpret_cast<float *>(my + vt1[v * 3 + t].x) = pv;
                        }
                    }
                    iter = 0;
                    active = true;
                } else {
                    exhausted = true;
                }
            }
        }
        const unsigned amask = __ballot_sync(FULL, active);
        if (amask == 0) {
            if (__ballot_sync(FULL, !exhausted) == 0) break;
            continue;
        }
        if (active) {
            __syncwarp(amask);
            // ================= horizontal step: lane j owns checks c = T*i + j ==================
#pragma unroll
            for (int w = 0; w < WMS; ++w) {
                const uint32_t sw = synd[w] >> j;
                const int cnt = min(VPW, (m - 32 * w - j + T - 1) / T);          // checks of this lane in word w
#pragma unroll 2
                for (int ii = 0; ii < cnt; ++ii) {
                    const int c = 32 * w + T * ii + j;
                    const unsigned char *q = my + (uint32_t)c * rowB;
                    float x[RW];
#pragma unroll
                    for (int k = 0; k < RW; ++k) x[k] = *reinterpret_cast<const float *>(q + (uint32_t)k * m * rowB);
                    uint32_t sg = (sw >> (T * ii)) << 31;                          // syndrome bit -> sign bit
                    float min1 = CUDART_INF_F, min2 = CUDART_INF_F;
#pragma unroll
                    for (int k = 0; k < RW; ++k) {
                        sg ^= __float_as_uint(x[k]);
                        const float a = fabsf(x[k]);
                        const float t = fmaxf(min1, a);
                        min1 = fminf(min1, a);
                        min2 = fminf(min2, t);
                    }
                    *reinterpret_cast<float2 *>(my8 + offM + (uint32_t)c * 2u * rowB) =
                        make_float2(__uint_as_float(__float_as_uint(min1) | (sg & 0x80000000u)), min2);
                }
            }
            __syncwarp(amask);

            // ================= vertical step: lane j owns variables v = T*i + j ==================
            const uint2 *vt = (iter == 0) ? vt0 : vt1;
            const bool last = (iter == max_iter - 1);
            const bool wr_llr = (P.llr != nullptr) && (P.llr_mode == LLR_ALL || (P.llr_mode == LLR_FAILED && last));
            float *llr_out = wr_llr ? reinterpret_cast<float *>(P.llr) + (size_t)shot * n : nullptr;
#pragma unroll
            for (int k = 0; k < WMS; ++k) acc[k] = 0;
            for (int wv = 0; wv < WN; ++wv) {
                uint32_t hw = 0;
                const int cnt = min(VPW, (n - 32 * wv - j + T - 1) / T);         // variables of this lane in word wv
#pragma unroll 2
                for (int ii = 0; ii < cnt; ++ii) {
                    const int b = T * ii + j;
                    const int v = 32 * wv + b;
                    const uint2 *ent = vt + v * 3;
                    uint2 e[3];
                    float qo[3], r[3];
#pragma unroll
                    for (int t = 0; t < 3; ++t) e[t] = ent[t];
#pragma unroll
                    for (int t = 0; t < 3; ++t) {
                        const float q = *reinterpret_cast<const float *>(my + e[t].x);
                        const float2 s12 = *reinterpret_cast<const float2 *>(my8 + e[t].y);
                        const float a1 = fabsf(s12.x);
                        const float mag = (fabsf(q) == a1) ? s12.y : a1;                          // decoding.py:51-53
                        r[t] = __uint_as_float(__float_as_uint(__fmul_rn(alpha, mag)) ^
                                               ((__float_as_uint(s12.x) ^ __float_as_uint(q)) & 0x80000000u)); // :55
                        qo[t] = q;
                    }
                    const float val = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), r[2]), prior[v]);    // :61-62
                    const bool hd = val < 0.f;
                    hw |= (uint32_t)hd << b;
                    if (wr_llr) llr_out[v] = val;
                    if (hd) {
#pragma unroll
                        for (int k = 0; k < WMS; ++k) acc[k] ^= colmask[v * WMS + k];
                    }
#pragma unroll
                    for (int t = 0; t < 3; ++t) {
                        float qn = __fsub_rn(val, r[t]);                                           // :63
                        qn = bp_damp(damp, qn, omd, qo[t]);                                        // :65
                        qn = fminf(fmaxf(qn, -clipv), clipv);                                      // :66
                        *reinterpret_cast<float *>(my + e[t].x) = qn;
                    }
                }
                // the T lanes of the shot hold disjoint bits of this word
#pragma unroll
                for (int o = G; o < 32; o <<= 1) hw |= __shfl_xor_sync(amask, hw, o);
                if (j == 0) *reinterpret_cast<uint32_t *>(my + offHW + (uint32_t)wv * rowB) = hw;
            }

            // ================= syndrome of the hard decision ======================================
            bool conv = true;
#pragma unroll
            for (int k = 0; k < WMS; ++k) {
                uint32_t a = acc[k];
#pragma unroll
                for (int o = G; o < 32; o <<= 1) a ^= __shfl_xor_sync(amask, a, o);
                conv = conv && (a == synd[k]);
            }

            __syncwarp(amask);          // hard-decision words written by sub-lane 0 are read by all T lanes
            if (conv || last) {
                uint32_t *ho = P.hard + (size_t)shot * WN;
                for (int w = j; w < WN; w += T) ho[w] = *reinterpret_cast<const uint32_t *>(my + offHW + (uint32_t)w * rowB);
                if (j == 0) {
                    P.conv[shot] = conv ? 1 : 0;
                    if (P.iters) P.iters[shot] = iter;
                    if (!conv && P.fail_idx) P.fail_idx[atomicAdd(P.fail_count, 1u)] = (int32_t)shot;
                    iter_sum += (unsigned long long)(iter + 1);
                }
                active = false;
            } else {
                ++iter;
            }
        }
    }

    if (P.iter_total) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) iter_sum += __shfl_xor_sync(FULL, iter_sum, o);
        if (lane == 0 && iter_sum) atomicAdd(P.iter_total, iter_sum);
    }
}

}  // namespace qldpc
