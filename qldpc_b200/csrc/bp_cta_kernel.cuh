// Min-sum BP, float32, messages in registers, for check matrices too large for one warp: ONE CTA PER SHOT.
//
// The mapping of bp_warp_kernel.cuh (a lane owns checks with their incoming messages in registers, scatters the outgoing
// messages into the columns of their variables, owns variables, gathers posteriors) spread over the NW warps of a CTA:
// warp w owns check slots [w*SC, (w+1)*SC) and variable slots [w*SV, (w+1)*SV) of the host-built labelling
// (bp_warp_layout.h, which also serves rows with fewer than RW edges: their spare edge slots read a row of +inf as
// "posterior", start at qpad >= every real |Q| and settle at +clip, so they never change a sign or a minimum; the
// sum-product variants pin them at +inf instead, psi(inf) = 0).  The
// three __syncwarp of the warp kernel become block barriers, the convergence vote a __syncthreads_and.
// Built for the space-time matrices of the reference (spaceTime.py: 864 x 2592, row weight 7-8, column weight <= 3 =>
// 12 warps x (3 check slots, 7 variable slots), 43 KB of shared memory per shot): the message state that the HBM-staged
// kernel streams through global memory every iteration (12 E bytes per shot-iteration) never leaves the SM.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "bp_kernel.cuh"
#include "bp_warp_kernel.cuh"

namespace qldpc {

// shared memory of a CTA: planes [3][VPL][32] + dump row + posteriors [VPL][32] + inf row, VPL = NW * SV
__host__ __device__ inline size_t bp_cta_smem(int VPL) { return 4 * (size_t)32 * (4 * VPL + 2); }

// VAR as in bp_warp_kernel: 0 min-sum, 1 sum-product, 2 symmetric sum-product (psi domain)
template <int SC, int SV, int RW, bool TWO, int VAR>
__global__ void __launch_bounds__(SC >= 3 ? 384 : 576, 1)
bp_cta_kernel(const BPParams P, const BPWarpTables W, int VPL)
{
    const int n = P.g.n, WN = P.g.WN, WM = P.g.WM;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NW = blockDim.x >> 5;
    const unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char smem[];
    float *Rbuf = reinterpret_cast<float *>(smem);                 // [3][VPL][32] + dump row
    float *Vbuf = Rbuf + 32 * (3 * VPL + 1);                       // [VPL][32] + inf row
    __shared__ long long s_next;

    // ---- per-lane tables into registers (BYTE offsets) ------------------------------------------------
    uint32_t sidx[SC][RW], vidx[SC][RW], cinfo[SC];
    float prior[SV];
#pragma unroll
    for (int i = 0; i < SV; ++i) {
        const uint32_t v = W.vorig[(warp * SV + i) * 32 + lane];
        prior[i] = (v != 0xffffffffu) ? reinterpret_cast<const float *>(P.prior)[v] + 0.f : 0.f;
    }
#pragma unroll
    for (int i = 0; i < SC; ++i) {
        const int ig = warp * SC + i;
        cinfo[i] = W.cinfo[ig * 32 + lane];
#pragma unroll
        for (int k = 0; k < RW; ++k) {
            vidx[i][k] = W.vidx[(ig * RW + k) * 32 + lane];
            sidx[i][k] = W.sidx[(ig * RW + k) * 32 + lane];
        }
    }
    uint32_t padmask = 0;                                          // edge slots that read the +inf row (SC * RW <= 32 bits)
#pragma unroll
    for (int i = 0; i < SC; ++i)
#pragma unroll
        for (int k = 0; k < RW; ++k) padmask |= (vidx[i][k] >= 4u * 32u * (uint32_t)VPL ? 1u : 0u) << (i * RW + k);
    for (int r = tid; r < 32 * (3 * VPL + 1); r += blockDim.x) Rbuf[r] = 0.f;       // columns of padding positions stay zero
    if (tid < 32) Vbuf[VPL * 32 + tid] = CUDART_INF_F;

    const float alpha = (float)P.alpha, damp = (float)P.damping, omd = (float)P.one_minus_damping, clipv = (float)P.clip;
    const float qpad = (float)P.qpad;
    const float prior0 = reinterpret_cast<const float *>(P.prior)[0] + 0.f;
    const int max_iter = P.max_iter;
    unsigned long long iter_sum = 0;

    auto load_synd = [&](long long sh, uint32_t (&w)[SC]) {
#pragma unroll
        for (int i = 0; i < SC; ++i) w[i] = (sh < P.B && cinfo[i] != 0xffffffffu) ? P.synd[(size_t)sh * WM + (cinfo[i] >> 5)] : 0u;
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
        uint32_t sbit[SC];
        float salpha[SC];
#pragma unroll
        for (int i = 0; i < SC; ++i) {
            sbit[i] = ((sw[i] >> (cinfo[i] & 31u)) & 1u) << 31;
            salpha[i] = __uint_as_float(__float_as_uint(VAR == 1 ? 1.f : alpha) ^ sbit[i]);
        }
        // Q = prior along the edges (decoding.py:21); padding slots: qpad (min-sum) or +inf (sum-product)
        float Q[SC][RW];
        if (P.prior_uniform) {
#pragma unroll
            for (int i = 0; i < SC; ++i)
#pragma unroll
                for (int k = 0; k < RW; ++k) Q[i][k] = ((padmask >> (i * RW + k)) & 1u) ? (VAR == 0 ? qpad : CUDART_INF_F) : prior0;
        } else {
#pragma unroll
            for (int i = 0; i < SV; ++i) Vbuf[(warp * SV + i) * 32 + lane] = prior[i] + 0.f;
            __syncthreads();
#pragma unroll
            for (int i = 0; i < SC; ++i)
#pragma unroll
                for (int k = 0; k < RW; ++k) Q[i][k] = (VAR == 0) ? fminf(ldb(Vbuf, vidx[i][k]), qpad) : ldb(Vbuf, vidx[i][k]);
        }

        int iter = 0;
        bool conv = false;
        for (;; ++iter) {
            // ================= horizontal step (lane-local; see bp_warp_kernel.cuh) =================
            float R[SC][RW];
#pragma unroll
            for (int i = 0; i < SC; ++i) {
                float pre[RW], suf[RW];
                if (VAR == 0) {
                    pre[1] = Q[i][0];
                    suf[RW - 2] = Q[i][RW - 1];
#pragma unroll
                    for (int k = 2; k < RW; ++k) pre[k] = bpw_xmin(pre[k - 1], Q[i][k - 1]);
#pragma unroll
                    for (int k = RW - 3; k >= 0; --k) suf[k] = bpw_xmin(suf[k + 1], Q[i][k + 1]);
                } else {
                    float ps[RW];                                      // psi(+inf) = 0: padding slots drop out of the sums
#pragma unroll
                    for (int k = 0; k < RW; ++k) ps[k] = bpw_psi(fabsf(Q[i][k]));
                    pre[1] = ps[0];
                    suf[RW - 2] = ps[RW - 1];
#pragma unroll
                    for (int k = 2; k < RW; ++k) pre[k] = pre[k - 1] + ps[k - 1];
#pragma unroll
                    for (int k = RW - 3; k >= 0; --k) suf[k] = suf[k + 1] + ps[k + 1];
                }
                uint32_t sgall = 0;
                if (VAR != 0) {
#pragma unroll
                    for (int k = 0; k < RW; ++k) sgall ^= __float_as_uint(Q[i][k]);
                }
                float o[RW];
#pragma unroll
                for (int k = 0; k < RW; ++k) {
                    if (VAR == 0) {
                        o[k] = (k == 0) ? suf[0] : (k == RW - 1) ? pre[RW - 1] : bpw_xmin(pre[k], suf[k]);
                    } else {
                        const float sk = (k == 0) ? suf[0] : (k == RW - 1) ? pre[RW - 1] : pre[k] + suf[k];
                        const float mag = fminf(bpw_psi(sk), 16.811242831518264f);
                        o[k] = __uint_as_float(__float_as_uint(mag) | ((sgall ^ __float_as_uint(Q[i][k])) & 0x80000000u));
                    }
                }
#pragma unroll
                for (int k = 0; k < RW; k += 2) {                      // packed float32 pairs, see bp_warp_kernel.cuh
                    bpw_mul2(o[k], o[k + 1], salpha[i], salpha[i], R[i][k], R[i][k + 1]);
#pragma unroll
                    for (int kk = k; kk < k + 2; ++kk) {
                        if (TWO && iter == 0) stb(Rbuf, __ldg(W.sidx0 + ((warp * SC + i) * RW + kk) * 32 + lane), R[i][kk]);
                        else stb(Rbuf, sidx[i][kk], R[i][kk]);
                    }
                }
            }
            __syncthreads();

            // ================= vertical step: posteriors of the owned variables =====================
            const bool last = (iter == max_iter - 1);
#pragma unroll
            for (int i = 0; i + 1 < SV; i += 2) {
                const int ig = warp * SV + i;
                float a0 = Rbuf[(0 * VPL + ig) * 32 + lane], a1 = Rbuf[(0 * VPL + ig + 1) * 32 + lane];
                bpw_add2(a0, a1, Rbuf[(1 * VPL + ig) * 32 + lane], Rbuf[(1 * VPL + ig + 1) * 32 + lane], a0, a1);
                bpw_add2(a0, a1, Rbuf[(2 * VPL + ig) * 32 + lane], Rbuf[(2 * VPL + ig + 1) * 32 + lane], a0, a1);
                bpw_add2(a0, a1, prior[i], prior[i + 1], a0, a1);
                Vbuf[ig * 32 + lane] = a0;
                Vbuf[(ig + 1) * 32 + lane] = a1;
            }
            if (SV & 1) {
                constexpr int i = SV - 1;
                const int ig = warp * SV + i;
                const float r0 = Rbuf[(0 * VPL + ig) * 32 + lane], r1 = Rbuf[(1 * VPL + ig) * 32 + lane], r2 = Rbuf[(2 * VPL + ig) * 32 + lane];
                Vbuf[ig * 32 + lane] = __fadd_rn(__fadd_rn(__fadd_rn(r0, r1), r2), prior[i]);
            }
            __syncthreads();

            // ================= Q update in registers + syndrome of the hard decision =================
            bool ok = true;
#pragma unroll
            for (int i = 0; i < SC; ++i) {
                uint32_t par = sbit[i];
#pragma unroll
                for (int k = 0; k < RW; k += 2) {
                    const float v0 = ldb(Vbuf, vidx[i][k]), v1 = ldb(Vbuf, vidx[i][k + 1]);
                    par ^= __float_as_uint(v0) ^ __float_as_uint(v1);   // (+inf of a padding slot: sign 0)
                    float q0, q1;
                    bpw_sub2(v0, v1, R[i][k], R[i][k + 1], q0, q1);
                    if (VAR != 1) {
                        float t0, t1;
                        bpw_mul2(omd, omd, Q[i][k], Q[i][k + 1], t0, t1);
                        bpw_fma2(damp, damp, q0, q1, t0, t1, q0, q1);
                        q0 = bpw_xmin(q0, clipv);     // np.clip(q, -c, c) for c >= 0: sign(q) * min(|q|, c), ONE FMNMX.XORSIGN
                        q1 = bpw_xmin(q1, clipv);
                    }
                    // sum-product: a padding slot must stay at +inf (psi = 0); at +clip it would add psi(clip) to the sums
                    if (VAR != 0 && ((padmask >> (i * RW + k)) & 1u)) q0 = CUDART_INF_F;
                    if (VAR != 0 && ((padmask >> (i * RW + k + 1)) & 1u)) q1 = CUDART_INF_F;
                    Q[i][k] = q0;
                    Q[i][k + 1] = q1;
                }
                ok = ok && (cinfo[i] == 0xffffffffu || (int)par >= 0);
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
            const float val = valid ? ldb(Vbuf, __ldg(W.vpos + i * 32 + lane)) : 0.f;
            const uint32_t w = __ballot_sync(FULL, valid && (val < 0.f));
            if (lane == 0) P.hard[(size_t)shot * WN + i] = w;
            if (wr_llr && valid) reinterpret_cast<float *>(P.llr)[(size_t)shot * n + lane + 32 * i] = val;
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
