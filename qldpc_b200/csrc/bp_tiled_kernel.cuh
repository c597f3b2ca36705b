// Production BP kernel for code-capacity check matrices (BB codes) on sm_100a:
// TL LANES PER SHOT, message state resident in shared memory.
//
// Same arithmetic as bp_decode_kernel (bp_kernel.cuh) -- min-sum (rework/decoding.py:5-75), sum-product
// (decoding/beliefPropagation.py:88-144) and its damped/scaled/clipped form (rework/decoding.py:131-191), in
// float32 or float64 -- with a different mapping; results are bit-identical to the thread-per-shot kernel.
//
// The thread-per-shot kernel is bound by shared-memory CAPACITY: [[144,12,12]] needs 2.3 KB of float32 message
// state per shot, so only 96 shots -- 3 warps -- fit on an SM and every dependent instruction is exposed (ncu
// r1a: 19 % issue utilisation, 43 % stall_wait, 33 % short_scoreboard).  Here each shot is decoded by TL = 4 or
// 8 lanes of one warp, so the same shared memory feeds TL times as many warps:
//   * check pass : lane j of a shot owns checks c = TL*i + j (min1 / min2 / sign parity -- or the tanh product --
//                  of its RW incoming messages, written as a per-check summary),
//   * variable pass: lane j owns variables v = TL*i + j (rebuilds the 3 check-to-variable messages from the
//                  summaries and the unmodified Q, posterior, hard decision, Q update),
//   * __syncwarp between the passes; the syndrome of the hard decision is accumulated per lane from packed
//     columns of H and XOR-reduced over the TL lanes with shuffles.
// Layout: message rows are numbered ELL-style r = k*m + c (k-th edge of check c) and stored [row][slot] with S
// slots per row, S = G (mod 32), G = 32/TL shots per warp, lane = j*G + g.  The G lanes of one j read G
// consecutive elements, and rows whose index differs mod TL start G banks apart, so the check pass (rows
// k*m + TL*i + j, j = 0..TL-1) is bank-conflict free and its addresses are affine; the variable pass reads rows
// through a table and is conflict-free whenever the TL rows it touches differ mod TL.
// The kernel is persistent; retired shots are replaced from a global cursor (refill is done by all TL lanes of
// the shot and only when at least `refill_min` shots of the warp are idle, or nothing is left to wait for).
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
// tsize: bytes per message (4 / 8); summary words per check: 2 for min-sum, 1 for sum-product
__host__ __device__ inline BPTiledLayout bp_tiled_layout(const BPGraphDev &g, int TL, int tsize, int variant)
{
    BPTiledLayout L;
    const int nsum = (variant == VAR_MIN_SUM) ? 2 : 1;
    L.npad = (g.n + TL - 1) / TL * TL;
    size_t o = 0;
    L.off_vt0 = o;     o += 32 * (size_t)L.npad;                // per position: {q0,q1,q2,v} {m0,m1,m2,prior bits}
    L.off_vt1 = g.two_tables ? o : L.off_vt0;
    if (g.two_tables) o += 32 * (size_t)L.npad;
    L.off_colmask = o; o += 4 * (size_t)L.npad * g.WM;
    o = (o + 7) & ~(size_t)7;
    L.off_prior = o;   o += (size_t)tsize * L.npad;
    o = (o + 127) & ~(size_t)127;
    L.off_state = o;
    L.tables = o;
    // rows: RW*m message rows + m summary rows (nsum messages wide) + WN hard-decision rows (4 B per slot)
    L.per_slot = (size_t)tsize * ((size_t)g.uniform_row_w * g.m + (size_t)nsum * g.m) + 4 * (size_t)g.WN;
    return L;
}

// per-check summary, moved with one vector access: {signed min1, min2} for min-sum, the signed product for sum-product
template <typename T, int NS> struct alignas(NS * sizeof(T)) SummaryT { T v[NS]; };

// vell tables (global, built by the host): 4 words per POSITION p (the variable processed by lane p % TL at step
// p / TL): entries t < 3 are c | (k << 16) (c = check, k = position of the variable inside row c) in the order the
// messages are added, word 3 is the variable index v.  Positions permute variables only inside a 32-variable word
// (so hard-decision bits stay in their word); the permutation is chosen on the host to spread the rows touched by
// the TL lanes of a step over distinct shared-memory banks.  Every variable has exactly 3 edges (all BB codes).
template <typename T, int VAR, int TL, int WMS, int RW>
__global__ void __launch_bounds__(576, 1)
bp_tiled_kernel(const BPParams P, const uint32_t *__restrict__ vell0, const uint32_t *__restrict__ vell1, int refill_min)
{
    typedef Num<T> N;
    typedef typename N::bits_t bits_t;
    constexpr int NS = (VAR == VAR_MIN_SUM) ? 2 : 1;
    typedef SummaryT<T, NS> sum_t;
    constexpr int G = 32 / TL;                // shots per warp
    constexpr int VPW = 32 / TL;              // variables (or checks) per lane per 32-bit word
    constexpr uint32_t TS = sizeof(T);
    const BPGraphDev &g = P.g;
    const int m = g.m, n = g.n, WN = g.WN;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int j = lane / G, gi = lane % G;
    const int S = (blockDim.x >> 5) * G;      // slots per row
    const int slot = warp * G + gi;
    const unsigned FULL = 0xffffffffu;

    extern __shared__ __align__(16) unsigned char smem[];
    const BPTiledLayout L = bp_tiled_layout(g, TL, (int)TS, VAR);
    uint4 *vt0 = reinterpret_cast<uint4 *>(smem + L.off_vt0);
    uint4 *vt1 = reinterpret_cast<uint4 *>(smem + L.off_vt1);
    uint32_t *colmask = reinterpret_cast<uint32_t *>(smem + L.off_colmask);
    T *prior = reinterpret_cast<T *>(smem + L.off_prior);
    unsigned char *state = smem + L.off_state;
    const uint32_t rowQ = TS * S;                               // bytes per message row
    const uint32_t rowM = NS * TS * S;                          // bytes per summary row
    const uint32_t offM = (uint32_t)RW * m * rowQ, offHW = offM + (uint32_t)m * rowM;

    for (int p = threadIdx.x; p < L.npad; p += blockDim.x) {
        auto mk = [&](const uint32_t *ve, uint4 *vt) {
            if (p >= n) {                                            // padding position: never visited by the passes
                vt[2 * p] = make_uint4(0u, 0u, 0u, 0xffffffffu);
                vt[2 * p + 1] = make_uint4(offM, offM, offM, 0u);
                return;
            }
            uint32_t qq[3], mm[3];
            for (int t = 0; t < 3; ++t) {
                const uint32_t e = ve[4 * p + t], c = e & 0xffffu, k = e >> 16;
                qq[t] = (k * m + c) * rowQ;
                mm[t] = offM + c * rowM;
            }
            const uint32_t v = ve[4 * p + 3];
            const uint32_t pb = (sizeof(T) == 4 && v < (uint32_t)n) ? __float_as_uint((float)reinterpret_cast<const T *>(P.prior)[v]) : 0u;
            vt[2 * p] = make_uint4(qq[0], qq[1], qq[2], v);
            vt[2 * p + 1] = make_uint4(mm[0], mm[1], mm[2], pb);
        };
        mk(vell0, vt0);
        if (g.two_tables) mk(vell1, vt1);
    }
    for (int i = threadIdx.x; i < L.npad * WMS; i += blockDim.x) colmask[i] = (i < n * WMS) ? g.colmask[i] : 0u;
    for (int i = threadIdx.x; i < L.npad; i += blockDim.x) prior[i] = (i < n) ? reinterpret_cast<const T *>(P.prior)[i] : (T)0;
    __syncthreads();

    unsigned char *myQ = state + TS * slot;                     // this shot's column in the message rows
    unsigned char *myM = state + NS * TS * slot;                // ... in the summary rows
    unsigned char *myH = state + 4 * slot;                      // ... in the hard-decision rows
    const T alpha = (T)P.alpha, damp = (T)P.damping, omd = (T)P.one_minus_damping, clipv = (T)P.clip;
    const int max_iter = P.max_iter;
    const bool sym = P.sym != 0;
    const bool slot_is_tanh = (VAR == VAR_SUM_PRODUCT) && !sym;

    uint32_t synd[WMS], acc[WMS];
    long long shot = -1;
    int iter = 0;
    bool active = false, exhausted = false;
    unsigned long long iter_sum = 0;

    while (true) {
        // ---- refill ------------------------------------------------------------------------
        const unsigned idle = __ballot_sync(FULL, !active && !exhausted);       // lanes (all TL of a shot together)
        const unsigned busy = __ballot_sync(FULL, active);
        const int idle_shots = __popc(idle) / TL;
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
                    // Q = where(mask, prior, 0) (decoding.py:21 / beliefPropagation.py:107): the lanes split the variables
                    for (int p = j; p < n; p += TL) {
                        const uint4 ea = vt1[2 * p];
                        if (ea.w >= (uint32_t)n) continue;                          // padding position
                        T pv = bp_canon(prior[ea.w]);
                        if (slot_is_tanh) pv = N::tanh_(N::mul(pv, (T)0.5));
                        *reinterpret_cast<T *>(myQ + ea.x) = pv;
                        *reinterpret_cast<T *>(myQ + ea.y) = pv;
                        *reinterpret_cast<T *>(myQ + ea.z) = pv;
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
            // ================= horizontal step: lane j owns checks c = TL*i + j ==================
#pragma unroll
            for (int w = 0; w < WMS; ++w) {
                const uint32_t sw = synd[w] >> j;
                const int cnt = min(VPW, (m - 32 * w - j + TL - 1) / TL);        // checks of this lane in word w
#pragma unroll 2
                for (int ii = 0; ii < cnt; ++ii) {
                    const int c = 32 * w + TL * ii + j;
                    const unsigned char *q = myQ + (uint32_t)c * rowQ;
                    T x[RW];
#pragma unroll
                    for (int k = 0; k < RW; ++k) x[k] = *reinterpret_cast<const T *>(q + (uint32_t)k * m * rowQ);
                    const bits_t sbit = (bits_t)((sw >> (TL * ii)) & 1u) << (8 * sizeof(bits_t) - 1);   // syndrome bit -> sign bit
                    sum_t sv;
                    if (VAR == VAR_MIN_SUM) {
                        // decoding.py:28-53: sign product, min1, min2
                        bits_t sg = sbit;
                        T min1 = N::inf(), min2 = N::inf();
#pragma unroll
                        for (int k = 0; k < RW; ++k) {
                            sg ^= N::bits(x[k]);
                            const T a = fabs(x[k]);
                            const T t = fmax(min1, a);
                            min1 = fmin(min1, a);
                            min2 = fmin(min2, t);
                        }
                        sv.v[NS - 1] = min2;
                        sv.v[0] = N::from_bits(N::bits(min1) | (sg & N::SIGN));           // signed min1 (written last: NS == 2 here)
                    } else {
                        // beliefPropagation.py:114-118: row product of tanh(Q/2), ascending column order
                        T prod = (T)1;
#pragma unroll
                        for (int k = 0; k < RW; ++k) {
                            const T t = slot_is_tanh ? x[k] : N::tanh_(N::mul(x[k], (T)0.5));
                            prod = N::mul(prod, t);
                        }
                        sv.v[0] = N::from_bits(N::bits(prod) ^ sbit);                             // * (1 - 2 s)
                    }
                    *reinterpret_cast<sum_t *>(myM + offM + (uint32_t)c * rowM) = sv;
                }
            }
            __syncwarp(amask);

            // ================= vertical step: lane j owns variables v = TL*i + j ==================
            const uint4 *vt = (iter == 0) ? vt0 : vt1;
            const bool last = (iter == max_iter - 1);
            const bool wr_llr = (P.llr != nullptr) && (P.llr_mode == LLR_ALL || (P.llr_mode == LLR_FAILED && last));
            T *llr_out = wr_llr ? reinterpret_cast<T *>(P.llr) + (size_t)shot * n : nullptr;
#pragma unroll
            for (int k = 0; k < WMS; ++k) acc[k] = 0;
            for (int wv = 0; wv < WN; ++wv) {
                uint32_t hw = 0;
                const int cnt = min(VPW, (n - 32 * wv - j + TL - 1) / TL);       // variables of this lane in word wv
#pragma unroll 2
                for (int ii = 0; ii < cnt; ++ii) {
                    const int p = 32 * wv + TL * ii + j;                           // position
                    const uint4 ea = vt[2 * p], eb = vt[2 * p + 1];
                    const uint32_t eq[3] = {ea.x, ea.y, ea.z}, em[3] = {eb.x, eb.y, eb.z};
                    const int v = (int)ea.w;                                          // variable at this position (same word)
                    const int b = v & 31;
                    T qo[3], r[3];
#pragma unroll
                    for (int t = 0; t < 3; ++t) {
                        const T q = *reinterpret_cast<const T *>(myQ + eq[t]);
                        const sum_t s12 = *reinterpret_cast<const sum_t *>(myM + em[t]);
                        const T s1 = s12.v[0];
                        if (VAR == VAR_MIN_SUM) {
                            const T s2 = s12.v[NS - 1];
                            const T a1 = fabs(s1);
                            const T mag = (fabs(q) == a1) ? s2 : a1;                              // decoding.py:51-53
                            r[t] = N::from_bits(N::bits(N::mul(alpha, mag)) ^ ((N::bits(s1) ^ N::bits(q)) & N::SIGN)); // :55
                        } else {
                            const T tq = slot_is_tanh ? q : N::tanh_(N::mul(q, (T)0.5));
                            const T ts = (fabs(tq) < (T)1e-15) ? (T)1e-15 : tq;                   // beliefPropagation.py:122
                            T rr = bp_sp_r(N::div(s1, ts));                                      // :125-126
                            if (sym) rr = N::mul(rr, alpha);                                     // decoding.py:171
                            r[t] = rr;
                        }
                        qo[t] = q;
                    }
                    const T pr = (sizeof(T) == 4) ? (T)__uint_as_float(eb.w) : prior[v];
                    const T val = N::add(N::add(N::add(r[0], r[1]), r[2]), pr);                  // values = R_sum + prior
                    const bool hd = val < (T)0;
                    hw |= (uint32_t)hd << b;
                    if (wr_llr) llr_out[v] = val;
                    if (hd) {
#pragma unroll
                        for (int k = 0; k < WMS; ++k) acc[k] ^= colmask[v * WMS + k];
                    }
#pragma unroll
                    for (int t = 0; t < 3; ++t) {
                        T qn = N::sub(val, r[t]);                                                 // Q_new = values - R
                        if (VAR == VAR_MIN_SUM || sym) {
                            qn = bp_damp(damp, qn, omd, qo[t]);                                   // decoding.py:65 / :179
                            qn = fmin(fmax(qn, -clipv), clipv);                                   // :66 / :181
                            qn = bp_canon(qn);
                        } else {
                            qn = N::tanh_(N::mul(qn, (T)0.5));                                    // the slot holds tanh(Q/2)
                        }
                        *reinterpret_cast<T *>(myQ + eq[t]) = qn;
                    }
                }
                // the TL lanes of the shot hold disjoint bits of this word
#pragma unroll
                for (int o = G; o < 32; o <<= 1) hw |= __shfl_xor_sync(amask, hw, o);
                if (j == 0) *reinterpret_cast<uint32_t *>(myH + offHW + (uint32_t)wv * (4u * S)) = hw;
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

            __syncwarp(amask);          // hard-decision words written by sub-lane 0 are read by all TL lanes
            if (conv || last) {
                uint32_t *ho = P.hard + (size_t)shot * WN;
                for (int w = j; w < WN; w += TL) ho[w] = *reinterpret_cast<const uint32_t *>(myH + offHW + (uint32_t)w * (4u * S));
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
