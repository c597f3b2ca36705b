// BP for check matrices too large for one warp (space-time / detector-error-model H), float32 or float64, min-sum or
// sum-product in the reference's own arithmetic -- ONE CTA PER SHOT, EDGE MESSAGES STAGED IN GLOBAL MEMORY.
//
// The variable-to-check messages Q live in a per-CTA array in global memory and are streamed through the SM every
// iteration with 128-bit, fully coalesced accesses -- read by the check pass, read again by the update pass, written once:
// 12 E bytes (float32) per shot-iteration, the algorithmic figure of SURVEY.md section 8d -- while everything a check or a
// variable reduces to stays on chip:
//   * check pass   : a lane owns check slots (the labelling of bp_cta_kernel.cuh / bp_warp_layout.h: 3 per lane, 12 warps
//                    for the 864 x 2592 space-time matrix); it loads the RW incoming messages of a check as RW * sizeof(T)
//                    / 16 vector loads, forms the RW outgoing messages R (min-sum: prefix / suffix minima, signs as bits;
//                    sum-product: the tanh product, division by the own factor, 2 atanh -- beliefPropagation.py:114-126)
//                    and scatters them into the shared-memory column of their variable (plane = position in the
//                    variable's addition order); nothing is kept in registers across the pass;
//   * variable pass: the owner of a variable adds its column in the reference's order, adds the prior, publishes the
//                    posterior (shared memory);
//   * update pass  : the check's owner reloads the message row, re-reads the R it scattered and the posteriors of its
//                    variables from shared memory, Q <- clip(damping * (posterior - R) + (1 - damping) * Q), stores the
//                    row; the syndrome test of the hard decision is the xor of the posterior signs it just read.
// Per-check summaries never exist in memory at all (the thread-per-shot staged kernel of bp_kernel.cuh writes {min1, min2}
// per check to HBM and re-reads them per edge: 1.6x the algorithmic traffic, 0.24-0.31 of the HBM roofline).  The staging
// array belongs to the CTA, not to the shot (E_pad * sizeof(T) = 36 / 72 KB): a few hundred CTAs keep it resident in the
// 126 MB L2, so DRAM sees almost none of the algorithmic traffic.  Registers hold one check at a time, which is what lets
// the float64 and the exact (tanh-domain) sum-product instantiations run with 12 warps per CTA.
// Arithmetic per type is that of the other kernels (Num<T>, bp_damp, bp_canon, bp_sp_r): results equal the thread-per-shot
// kernels bit for bit for min-sum; for sum-product the row product is taken in slot order, not in column order (last-ulp
// differences in the tanh product).
#pragma once
#include "sp_math.cuh"
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "bp_kernel.cuh"
#include "bp_warp_kernel.cuh"

namespace qldpc {

// shared memory of a CTA: planes [3][VPL][32] + dump row + posteriors [VPL][32] + inf row, VPL = NW * SV
__host__ __device__ inline size_t bp_stage_smem(int VPL, int tsize) { return (size_t)tsize * 32 * (4 * VPL + 2); }
// staging array of a CTA: [NW * SC check slots][RW][32 lanes] messages
__host__ __device__ inline size_t bp_stage_gstate(int NW, int SC, int RW, int tsize) { return (size_t)tsize * NW * SC * RW * 32; }

template <typename T> struct StageVec;
template <> struct StageVec<float> { typedef float4 type; static constexpr int N = 4; };
template <> struct StageVec<double> { typedef double2 type; static constexpr int N = 2; };

template <typename T> __device__ __forceinline__ T stage_ldb(const T *base, uint32_t byte_off4)
{
    // table entries are byte offsets of 4-byte elements: scale to sizeof(T)
    return *reinterpret_cast<const T *>(reinterpret_cast<const unsigned char *>(base) + (size_t)byte_off4 * (sizeof(T) / 4));
}
template <typename T> __device__ __forceinline__ void stage_stb(T *base, uint32_t byte_off4, T v)
{
    *reinterpret_cast<T *>(reinterpret_cast<unsigned char *>(base) + (size_t)byte_off4 * (sizeof(T) / 4)) = v;
}

// VAR: 0 min-sum (rework/decoding.py:5-75), 1 sum-product (beliefPropagation.py:88-144; the staged word is tanh(Q / 2)),
// 2 sum-product with alpha, damping and clipping (rework/decoding.py:131-191; the staged word is Q)
// TABREG: the scatter / gather offsets of the owned edge slots live in registers (2 * SC * RW of them: one CTA per SM), or
// are re-read from the (L1-resident) tables at every use, which leaves room for two CTAs per SM -- the choice for the
// FP64-latency-bound float64 instantiations.
// sum-product arithmetic of this kernel: float64 takes the branch-free FP64-pipe versions of sp_math.cuh (tanh(q/2): 23 FP64
// instructions instead of 34 + the library's special cases, 2 atanh: 19 instead of 76, division by the own factor: 4 instead of
// ~20), float32 the math library.
__device__ __forceinline__ double stage_tanh_half(double q) { return spm_tanh_half(q); }
__device__ __forceinline__ float stage_tanh_half(float q) { return tanhf(__fmul_rn(q, 0.5f)); }
__device__ __forceinline__ double stage_sp_r(double x) { return spm_2atanh_clipped(x); }
__device__ __forceinline__ float stage_sp_r(float x) { return bp_sp_r(x); }
// prod / ts for 1e-15 <= |ts| <= 1
__device__ __forceinline__ double stage_div(double prod, double ts)
{
    const double q = spm_div(prod, fabs(ts));
    return spm_make(spm_hi(q) ^ (spm_hi(ts) & 0x80000000u), spm_lo(q));
}
__device__ __forceinline__ float stage_div(float prod, float ts) { return __fdiv_rn(prod, ts); }

template <typename T, int VAR, int SC, int SV, int RW, bool TWO, bool TABREG>
__global__ void __launch_bounds__(SC >= 3 ? 384 : 576, TABREG ? 1 : 2)
bp_stage_kernel(const BPParams P, const BPWarpTables W, int VPL)
{
    typedef Num<T> N;
    typedef typename N::bits_t bits_t;
    typedef typename StageVec<T>::type vec_t;
    constexpr int VN = StageVec<T>::N, NG = RW / VN;
    static_assert(RW % VN == 0, "a message row is a whole number of 128-bit words");
    const int n = P.g.n, WN = P.g.WN, WM = P.g.WM;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NW = blockDim.x >> 5;
    const unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char smem[];
    T *Rbuf = reinterpret_cast<T *>(smem);                     // [3][VPL][32] + dump row
    T *Vbuf = Rbuf + 32 * (3 * VPL + 1);                       // [VPL][32] + inf row
    __shared__ long long s_next;
    // staging array of this CTA: row (check slot ig, group g) = 32 lanes x 128 bits
    vec_t *Qg = reinterpret_cast<vec_t *>(reinterpret_cast<unsigned char *>(P.gstate) + (size_t)blockIdx.x * bp_stage_gstate(NW, SC, RW, (int)sizeof(T)));
    auto row_ptr_of = [&](int i, int g) { return Qg + ((size_t)((warp * SC + i) * NG + g) * 32 + lane); };

    // ---- per-lane tables into registers (BYTE offsets of 4-byte elements) ----------------------------
    uint32_t sidx[TABREG ? SC : 1][TABREG ? RW : 1], vidx[TABREG ? SC : 1][TABREG ? RW : 1], cinfo[SC];
    auto SI = [&](int i, int k) -> uint32_t { return TABREG ? sidx[TABREG ? i : 0][TABREG ? k : 0] : __ldg(W.sidx + ((warp * SC + i) * RW + k) * 32 + lane); };
    auto VI = [&](int i, int k) -> uint32_t { return TABREG ? vidx[TABREG ? i : 0][TABREG ? k : 0] : __ldg(W.vidx + ((warp * SC + i) * RW + k) * 32 + lane); };
    T prior[SV];
#pragma unroll
    for (int i = 0; i < SV; ++i) {
        const uint32_t v = W.vorig[(warp * SV + i) * 32 + lane];
        prior[i] = (v != 0xffffffffu) ? N::add(reinterpret_cast<const T *>(P.prior)[v], (T)0) : (T)0;     // (+ 0: -0.0 -> +0.0)
    }
#pragma unroll
    for (int i = 0; i < SC; ++i) {
        const int ig = warp * SC + i;
        cinfo[i] = W.cinfo[ig * 32 + lane];
        if (TABREG) {
#pragma unroll
            for (int k = 0; k < RW; ++k) {
                vidx[TABREG ? i : 0][TABREG ? k : 0] = W.vidx[(ig * RW + k) * 32 + lane];
                sidx[TABREG ? i : 0][TABREG ? k : 0] = W.sidx[(ig * RW + k) * 32 + lane];
            }
        }
    }
    uint32_t padmask = 0;                                          // edge slots that read the +inf row (SC * RW <= 32 bits)
#pragma unroll
    for (int i = 0; i < SC; ++i)
#pragma unroll
        for (int k = 0; k < RW; ++k) padmask |= (VI(i, k) >= 4u * 32u * (uint32_t)VPL ? 1u : 0u) << (i * RW + k);
    for (int r = tid; r < 32 * (3 * VPL + 1); r += blockDim.x) Rbuf[r] = (T)0;       // columns of padding positions stay zero
    if (tid < 32) Vbuf[VPL * 32 + tid] = N::inf();

    const T alpha = (T)P.alpha, damp = (T)P.damping, omd = (T)P.one_minus_damping, clipv = (T)P.clip;
    const T qpad = (T)P.qpad;
    const int max_iter = P.max_iter;
    unsigned long long iter_sum = 0;
    // value of a padding edge slot: never a minimum / a factor 1 of the product, never a sign
    const T pad_word = (VAR == 0) ? qpad : (VAR == 1 ? (T)1 : N::inf());

    auto load_synd = [&](long long sh, uint32_t (&w)[SC]) {
#pragma unroll
        for (int i = 0; i < SC; ++i) w[i] = (sh < P.B && cinfo[i] != 0xffffffffu) ? P.synd[(size_t)sh * WM + (cinfo[i] >> 5)] : 0u;
    };
    auto load_row = [&](int i, T (&q)[RW]) {
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            const vec_t v = *row_ptr_of(i, g);
            const T *e = reinterpret_cast<const T *>(&v);
#pragma unroll
            for (int x = 0; x < VN; ++x) q[g * VN + x] = e[x];
        }
    };
    auto store_row = [&](int i, const T (&q)[RW]) {
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            vec_t v;
            T *e = reinterpret_cast<T *>(&v);
#pragma unroll
            for (int x = 0; x < VN; ++x) e[x] = q[g * VN + x];
            *row_ptr_of(i, g) = v;
        }
    };

    if (tid == 0) s_next = (long long)atomicAdd(P.cursor, 1ull);
    __syncthreads();
    long long shot = s_next;
    uint32_t sw[SC];
    load_synd(shot, sw);
    __syncthreads();

    while (shot < P.B) {
        if (tid == 0) s_next = (long long)atomicAdd(P.cursor, 1ull);     // read by everyone after iteration 0
        long long next_shot = 0;
        bits_t sbit[SC];                                            // syndrome bit of each owned check, at the sign-bit position
#pragma unroll
        for (int i = 0; i < SC; ++i) sbit[i] = ((sw[i] >> (cinfo[i] & 31u)) & 1u) ? N::SIGN : (bits_t)0;
        // Q = prior along the edges (decoding.py:21 / beliefPropagation.py:107): publish the priors, gather, stage
#pragma unroll
        for (int i = 0; i < SV; ++i) Vbuf[(warp * SV + i) * 32 + lane] = prior[i];
        __syncthreads();
#pragma unroll
        for (int i = 0; i < SC; ++i) {
            T q[RW];
#pragma unroll
            for (int k = 0; k < RW; ++k) {
                T x = stage_ldb(Vbuf, VI(i, k));
                if (VAR == 1) x = stage_tanh_half(x);
                q[k] = ((padmask >> (i * RW + k)) & 1u) ? pad_word : x;
            }
            store_row(i, q);
        }
        // (a lane only ever touches its own staged rows: no barrier needed between this store and the loads below)

        int iter = 0;
        bool conv = false;
        for (;; ++iter) {
            // ================= check pass: R of every edge into the column of its variable =================
#pragma unroll
            for (int i = 0; i < SC; ++i) {
                T q[RW], r[RW];
                load_row(i, q);
                if (VAR == 0) {
                    // R[k] = alpha * (-1)^s * prod_{j != k} sign(Q[j]) * min_{j != k} |Q[j]| (decoding.py:41-55): the minimum
                    // over the others IS min1, or min2 at the arg-min (ties included)
                    T pre[RW], suf[RW];
                    pre[1] = fabs(q[0]);
                    suf[RW - 2] = fabs(q[RW - 1]);
#pragma unroll
                    for (int k = 2; k < RW; ++k) pre[k] = fmin(pre[k - 1], fabs(q[k - 1]));
#pragma unroll
                    for (int k = RW - 3; k >= 0; --k) suf[k] = fmin(suf[k + 1], fabs(q[k + 1]));
                    bits_t sgall = sbit[i];
#pragma unroll
                    for (int k = 0; k < RW; ++k) sgall ^= N::bits(q[k]);
#pragma unroll
                    for (int k = 0; k < RW; ++k) {
                        const T o = (k == 0) ? suf[0] : (k == RW - 1) ? pre[RW - 1] : fmin(pre[k], suf[k]);
                        r[k] = N::from_bits(N::bits(N::mul(alpha, o)) ^ ((sgall ^ N::bits(q[k])) & N::SIGN));     // :55
                    }
                } else {
                    // beliefPropagation.py:114-126 / decoding.py:157-171: t = tanh(Q / 2), row product, division by the own
                    // factor (t_safe), 2 atanh(clip); `* (1 - 2 s)` on the product
                    T t[RW];
                    T prod = (T)1;
#pragma unroll
                    for (int k = 0; k < RW; ++k) {
                        t[k] = (VAR == 1) ? q[k] : stage_tanh_half(q[k]);
                        prod = N::mul(prod, t[k]);
                    }
                    prod = N::from_bits(N::bits(prod) ^ sbit[i]);
#pragma unroll
                    for (int k = 0; k < RW; ++k) {
                        const T ts = (fabs(t[k]) < (T)1e-15) ? (T)1e-15 : t[k];
                        T rr = stage_sp_r(stage_div(prod, ts));
                        if (VAR == 2) rr = N::mul(rr, alpha);                                                   // decoding.py:171
                        r[k] = rr;
                    }
                }
#pragma unroll
                for (int k = 0; k < RW; ++k) {
                    if (TWO && iter == 0) stage_stb(Rbuf, __ldg(W.sidx0 + ((warp * SC + i) * RW + k) * 32 + lane), r[k]);
                    else stage_stb(Rbuf, SI(i, k), r[k]);           // (padding slots deliver into the dump row)
                }
            }
            __syncthreads();

            // ================= variable pass: posteriors of the owned variables =====================
            const bool last = (iter == max_iter - 1);
#pragma unroll
            for (int i = 0; i < SV; ++i) {
                const int ig = warp * SV + i;
                const T r0 = Rbuf[(0 * VPL + ig) * 32 + lane], r1 = Rbuf[(1 * VPL + ig) * 32 + lane], r2 = Rbuf[(2 * VPL + ig) * 32 + lane];
                Vbuf[ig * 32 + lane] = N::add(N::add(N::add(r0, r1), r2), prior[i]);
            }
            __syncthreads();

            // ================= update pass: reload the row, Q update, syndrome of the hard decision =================
            bool ok = true;
#pragma unroll
            for (int i = 0; i < SC; ++i) {
                T q[RW];
                load_row(i, q);
                bits_t par = sbit[i];
#pragma unroll
                for (int k = 0; k < RW; ++k) {
                    const T val = stage_ldb(Vbuf, VI(i, k));
                    const T rk = stage_ldb(Rbuf, (TWO && iter == 0) ? __ldg(W.sidx0 + ((warp * SC + i) * RW + k) * 32 + lane) : SI(i, k));
                    par ^= N::bits(val);                               // (+inf of a padding slot: sign 0)
                    T qn = N::sub(val, rk);                                                                     // :63 / :133
                    if (VAR != 1) {
                        qn = bp_damp(damp, qn, omd, q[k]);                                                     // :65 / :179
                        qn = fmin(fmax(qn, -clipv), clipv);                                                    // :66 / :181
                        qn = bp_canon(qn);
                    } else {
                        qn = stage_tanh_half(qn);
                    }
                    // padding slots: min-sum lets them settle at +clip (>= every real |Q|); the sum-product ones stay neutral
                    if (VAR != 0 && ((padmask >> (i * RW + k)) & 1u)) qn = pad_word;
                    q[k] = qn;
                }
                store_row(i, q);
                ok = ok && (cinfo[i] == 0xffffffffu || (par & N::SIGN) == 0);
            }
            conv = __syncthreads_and(ok) != 0;
            if (iter == 0) {
                next_shot = s_next;
                load_synd(next_shot, sw);
            }
            if (conv || last) break;
        }

        // ---- retire the shot: hard decision = sign of the posteriors, in the order of H ----------------------
        const bool wr_llr = P.llr != nullptr && (P.llr_mode == LLR_ALL || (P.llr_mode == LLR_FAILED && !conv));
        for (int i = warp; i < WN; i += NW) {
            const bool valid = lane + 32 * i < n;
            const T val = valid ? stage_ldb(Vbuf, __ldg(W.vpos + i * 32 + lane)) : (T)0;
            const uint32_t w = __ballot_sync(FULL, valid && (val < (T)0));
            if (lane == 0) P.hard[(size_t)shot * WN + i] = w;
            if (wr_llr && valid) reinterpret_cast<T *>(P.llr)[(size_t)shot * n + lane + 32 * i] = val;
        }
        if (tid == 0) {
            P.conv[shot] = conv ? 1 : 0;
            if (P.iters) P.iters[shot] = iter;
            if (!conv && P.fail_idx) P.fail_idx[atomicAdd(P.fail_count, 1u)] = (int32_t)shot;
            iter_sum += (unsigned long long)(iter + 1);
        }
        shot = next_shot;
        __syncthreads();                           // the posteriors are overwritten by the next shot's priors
    }
    if (P.iter_total && tid == 0 && iter_sum) atomicAdd(P.iter_total, iter_sum);
}

}  // namespace qldpc
