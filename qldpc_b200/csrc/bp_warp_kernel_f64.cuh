// Float64 instantiation of the warp-per-shot min-sum kernel (bp_warp_kernel.cuh): the bit-exact parity mode.
//
// Same mapping, labelling and tables; differences: no xorsign-min in float64, so the check pass keeps prefix / suffix minima
// of |Q| and handles the signs as bits; the damping is three roundings (NumPy evaluates the two products and the sum as
// separate ufuncs, bp_damp(double)); -0.0 messages are canonicalised (sign(0) = + in the reference).  Values equal the
// reference's (and the T-lanes-per-shot / thread-per-shot float64 kernels') bit for bit: the minimum over the other edges is
// min1, or min2 at the arg-min, ties included.  One 8-warp CTA per SM (about 170 registers per thread).
#pragma once
#include "bp_warp_kernel.cuh"

namespace qldpc {

__device__ __forceinline__ double ldbd(const double *base, uint32_t byte_off)
{
    return *reinterpret_cast<const double *>(reinterpret_cast<const unsigned char *>(base) + byte_off);
}
__device__ __forceinline__ void stbd(double *base, uint32_t byte_off, double v)
{
    *reinterpret_cast<double *>(reinterpret_cast<unsigned char *>(base) + byte_off) = v;
}

template <int CPL, int VPL, int RW, bool TWO>
__global__ void __launch_bounds__(BPW_WARPS * 32, 1)
bp_warp_kernel_f64(const BPParams P, const BPWarpTables W)
{
    const int n = P.g.n, WN = P.g.WN, WM = P.g.WM;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char smem[];
    double *Rbuf = reinterpret_cast<double *>(smem + 2 * bp_warp_smem_per_warp(VPL) * warp);     // [3][VPL][32] + dump row
    double *Vbuf = Rbuf + 32 * (3 * VPL + 1);

    // ---- per-lane tables into registers (BYTE offsets into the R / posterior buffers) -------------
    uint32_t sidx[CPL][RW], vidx[CPL][RW], cinfo[CPL];
    double prior[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const uint32_t v = W.vorig[i * 32 + lane];
        prior[i] = (v != 0xffffffffu) ? reinterpret_cast<const double *>(P.prior)[v] + 0.0 : 0.0;    // (+ 0: a -0.0 prior becomes +0.0)
    }
#pragma unroll
    for (int i = 0; i < CPL; ++i) {
        cinfo[i] = W.cinfo[i * 32 + lane];
#pragma unroll
        for (int k = 0; k < RW; ++k) {
            vidx[i][k] = 2u * W.vidx[(i * RW + k) * 32 + lane];        // (tables hold byte offsets of 4-byte elements)
            sidx[i][k] = 2u * W.sidx[(i * RW + k) * 32 + lane];
        }
    }
#pragma unroll
    for (int r = 0; r < 3 * VPL + 1; ++r) Rbuf[r * 32 + lane] = 0.0;      // columns of padding positions stay zero for ever
    Vbuf[VPL * 32 + lane] = CUDART_INF;

    const double alpha = P.alpha, damp = P.damping, omd = P.one_minus_damping, clipv = P.clip;
    const int max_iter = P.max_iter;
    unsigned long long iter_sum = 0;

    // The shot index (global cursor) and the syndrome words of the NEXT shot are fetched while the current one is being
    // decoded, so that neither the atomic nor the load latency is exposed between two shots.
    auto load_synd = [&](long long sh, uint32_t (&w)[CPL]) {
#pragma unroll
        for (int i = 0; i < CPL; ++i) w[i] = (sh < P.B && cinfo[i] != 0xffffffffu) ? P.synd[(size_t)sh * WM + (cinfo[i] >> 5)] : 0u;
    };
    unsigned long long s0 = 0;
    if (lane == 0) s0 = atomicAdd(P.cursor, 1ull);
    long long shot = (long long)__shfl_sync(FULL, s0, 0);
    uint32_t sw[CPL];
    load_synd(shot, sw);

    while (shot < P.B) {
        if (lane == 0) s0 = atomicAdd(P.cursor, 1ull);       // consumed after iteration 0
        long long next_shot = 0;
        uint32_t sbit[CPL];                      // syndrome bit of each owned check, moved to the sign-bit position
#pragma unroll
        for (int i = 0; i < CPL; ++i) {
            sbit[i] = ((sw[i] >> (cinfo[i] & 31u)) & 1u) << 31;
        }
        // Q = where(mask, prior, 0) (decoding.py:21): publish the priors, gather them along the edges
        __syncwarp();
#pragma unroll
        for (int i = 0; i < VPL; ++i) Vbuf[i * 32 + lane] = prior[i];
        __syncwarp();
        double Q[CPL][RW];
#pragma unroll
        for (int i = 0; i < CPL; ++i)
#pragma unroll
            for (int k = 0; k < RW; ++k) Q[i][k] = ldbd(Vbuf, vidx[i][k]);     // (every check has RW edges here: only whole padding
                                                                                 //  lanes read the +inf row, and nothing reads them)

        int iter = 0;
        bool conv = false;
        for (;; ++iter) {
            // ================= horizontal step (lane-local) ========================================
            // R[k] = alpha * (-1)^s * prod_{j != k} sign(Q[j]) * min_{j != k} |Q[j]| (decoding.py:41-55).  The minimum over
            // the others IS min1, or min2 at the arg-min (ties included), so the values equal the reference's
            // where(|Q| == min1, min2, min1) selection exactly.
            double R[CPL][RW];
#pragma unroll
            for (int i = 0; i < CPL; ++i) {
                // magnitudes: prefix / suffix minima of |Q| ("all but k"); signs: xor of all sign bits, then of the own one
                double pre[RW], suf[RW];
                uint32_t sgall = sbit[i];
                pre[1] = fabs(Q[i][0]);
                suf[RW - 2] = fabs(Q[i][RW - 1]);
#pragma unroll
                for (int k = 2; k < RW; ++k) pre[k] = fmin(pre[k - 1], fabs(Q[i][k - 1]));
#pragma unroll
                for (int k = RW - 3; k >= 0; --k) suf[k] = fmin(suf[k + 1], fabs(Q[i][k + 1]));
#pragma unroll
                for (int k = 0; k < RW; ++k) sgall ^= (uint32_t)__double2hiint(Q[i][k]);
#pragma unroll
                for (int k = 0; k < RW; ++k) {
                    const double o = (k == 0) ? suf[0] : (k == RW - 1) ? pre[RW - 1] : fmin(pre[k], suf[k]);
                    const double am = __dmul_rn(alpha, o);                                           // :55
                    const uint32_t sg = (sgall ^ (uint32_t)__double2hiint(Q[i][k])) & 0x80000000u;
                    const double r = __hiloint2double(__double2hiint(am) ^ (int)sg, __double2loint(am));
                    R[i][k] = r;
                    if (TWO && iter == 0) stbd(Rbuf, 2u * __ldg(W.sidx0 + (i * RW + k) * 32 + lane), r);
                    else stbd(Rbuf, sidx[i][k], r);       // (padding lanes write garbage into the dump row)
                }
            }
            __syncwarp();

            // ================= vertical step: posteriors of the owned variables =====================
            const bool last = (iter == max_iter - 1);
#pragma unroll
            for (int i = 0; i < VPL; ++i) {
                const double r0 = Rbuf[(0 * VPL + i) * 32 + lane], r1 = Rbuf[(1 * VPL + i) * 32 + lane], r2 = Rbuf[(2 * VPL + i) * 32 + lane];
                Vbuf[i * 32 + lane] = __dadd_rn(__dadd_rn(__dadd_rn(r0, r1), r2), prior[i]);        // :61-62
            }
            __syncwarp();

            // ================= Q update in registers + syndrome of the hard decision =================
            // The check is satisfied by the hard decisions iff the xor of the posterior sign bits equals its syndrome bit.
            bool ok = true;
#pragma unroll
            for (int i = 0; i < CPL; ++i) {
                uint32_t par = sbit[i];
#pragma unroll
                for (int k = 0; k < RW; ++k) {
                    const double val = ldbd(Vbuf, vidx[i][k]);
                    par ^= (uint32_t)__double2hiint(val);   // sign bit == hard decision (a sum with a non-zero prior is never -0.0)
                    double qn = __dsub_rn(val, R[i][k]);                                          // :63
                    qn = bp_damp(damp, qn, omd, Q[i][k]);                                         // :65 (three roundings, like NumPy)
                    qn = fmin(fmax(qn, -clipv), clipv);                                           // :66
                    Q[i][k] = bp_canon(qn);                                                       // sign(0) = + (decoding.py:30)
                }
                ok = ok && (cinfo[i] == 0xffffffffu || (int)par >= 0);
            }
            conv = __all_sync(FULL, ok);
            if (iter == 0) {
                next_shot = (long long)__shfl_sync(FULL, s0, 0);
                load_synd(next_shot, sw);
            }
            if (conv || last) break;
        }

        // ---- retire the shot: hard decision = sign of the posteriors still in the buffer, in the order of H ----
        uint32_t myw = 0;
        const bool wr_llr = P.llr != nullptr && (P.llr_mode == LLR_ALL || (P.llr_mode == LLR_FAILED && !conv));
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            if (i < WN) {                          // (uniform) VPL >= WN: a labelling may use more slots than ceil(n / 32)
                const bool valid = lane + 32 * i < n;
                const double val = valid ? ldbd(Vbuf, 2u * __ldg(W.vpos + i * 32 + lane)) : 0.0;
                const uint32_t w = __ballot_sync(FULL, valid && (val < 0.0));
                if (lane == i) myw = w;
                if (wr_llr && valid) reinterpret_cast<double *>(P.llr)[(size_t)shot * n + lane + 32 * i] = val;
            }
        }
        if (lane < WN) P.hard[(size_t)shot * WN + lane] = myw;
        if (lane == 0) {
            P.conv[shot] = conv ? 1 : 0;
            if (P.iters) P.iters[shot] = iter;
            if (!conv && P.fail_idx) P.fail_idx[atomicAdd(P.fail_count, 1u)] = (int32_t)shot;
            iter_sum += (unsigned long long)(iter + 1);
        }
        shot = next_shot;
    }
    if (P.iter_total && lane == 0 && iter_sum) atomicAdd(P.iter_total, iter_sum);
}

}  // namespace qldpc
