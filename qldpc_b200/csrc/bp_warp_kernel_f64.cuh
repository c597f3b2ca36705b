// Float64 instantiation of the warp-per-shot min-sum kernel (bp_warp_kernel.cuh): the BIT-EXACT parity mode
// (rework/decoding.py:5-75 performs exactly these float64 operations in exactly this order).
//
// Same mapping as the float32 kernel -- a lane owns checks with their incoming messages in registers, scatters the
// outgoing messages into the columns of their variables, owns variables, gathers posteriors -- written for the FP64 pipe
// of sm_100a (16 lanes per SM sub-partition: a DP instruction occupies the pipe for two issue cycles, so everything that
// is not an IEEE operation of the reference is kept OFF that pipe or made as cheap as possible):
//   * there is no 64-bit FMNMX: "the message with the smaller magnitude" is one DSETP on |a|, |b| (operand modifiers, no
//     extra instruction) plus two SEL on the register halves -- 3 issue slots instead of the 7 of fmin(fabs, fabs); the
//     chain carries whatever sign the winner had, magnitudes are taken by the |.| modifier of the consumer (DSETP, DMUL);
//   * signs are bits: xor of the high words, put on alpha * min with one LOP3;
//   * the clip is one DSETP on |q| and two SEL against (+-clip) -- not fmin(fmax());
//   * the damping keeps the reference's three roundings (NumPy evaluates d * Q_new, (1 - d) * Q_old and their sum as separate
//     ufuncs): DMUL, DMUL, DADD, no FMA contraction;
//   * no canonicalisation of -0.0: with damping > 0 a message can only become -0.0 by the underflow of damping * x for a
//     non-zero x below 2^-1074 / damping, and |x| >= ulp(prior) * 0.3^iterations here (the launcher sends damping <= 0,
//     negative clips and max_iter > 500 to the thread-per-shot kernel, which canonicalises);
//   * shared memory holds messages and posteriors as 64-bit words.  A 64-bit access is served one half-warp at a time and two
//     lanes of a half collide when their columns are equal mod 16 (same pair of banks), which the float32 labelling does
//     not exclude: the kernel runs under its own labelling, built for 16-lane conflict domains (bp_warp_layout.h, lanes =
//     16).  Measured on [[144,12,12]]: 1.46 G shot-iterations/s with the float32 labelling, 1.54 G with split planes of
//     low / high words (conflict-free 32-bit accesses, twice the LSU instructions), 1.74 G with the half-warp labelling.
// 477 issue slots and 159 DP instructions per shot-iteration of [[144,12,12]] against 896 / 195 in the first version
// (0.79 G shot-iterations/s); 4-warp CTAs, three per SM.  Results equal the thread-per-shot / tiled float64 kernels bit
// for bit (tests).
//
// VAR = 1 / 2: the reference's SUM-PRODUCT decoders in float64 on the same mapping (VAR 1: performBeliefPropagationFast,
// beliefPropagation.py:88-144; VAR 2: performBeliefPropagation_Symmetric, rework/decoding.py:131-191 -- alpha, damping, clip).
// The check update runs in the reference's tanh domain: t = tanh(Q / 2) per edge, the product of the OTHER five factors by a
// prefix and a suffix chain (the reference divides the row product by the own factor, the same number up to rounding; a
// check that holds a factor below the reference's 1e-15 guard takes the division path so that its `tanh_Q_safe` semantics
// are kept), clip to +-0.9999999 and 2 atanh -- with the branch-free float64 tanh / atanh of sp_math.cuh (42 FP64
// instructions per edge instead of the math library's 130 + ~100 integer ones).  Not bit-exact (NumPy's tanh / arctanh differ
// from any other implementation in the last place); held to 1e-7 relative on the golden shots, north_star bar 1e-4.
#pragma once
#include "bp_warp_kernel.cuh"
#include "sp_math.cuh"

namespace qldpc {

constexpr int BPW64_WARPS = 4;          // warps per CTA of the float64 kernel

// shared memory per warp: message planes [3][VPL][32] + dump row, posteriors [VPL][32] + inf row, 8 bytes per entry
__host__ __device__ inline size_t bp_warp64_smem_per_warp(int VPL) { return 8 * (size_t)32 * (4 * VPL + 2); }

__device__ __forceinline__ uint32_t d_hi(double x) { return (uint32_t)__double2hiint(x); }
__device__ __forceinline__ uint32_t d_lo(double x) { return (uint32_t)__double2loint(x); }
__device__ __forceinline__ double d_make(uint32_t hi, uint32_t lo) { return __hiloint2double((int)hi, (int)lo); }

// the operand with the smaller magnitude (sign carried along, NaN-free inputs): DSETP.LT |a|, |b| + 2 SEL
__device__ __forceinline__ double bpw_absmin(double a, double b)
{
    const bool lt = fabs(a) < fabs(b);
    return d_make(lt ? d_hi(a) : d_hi(b), lt ? d_lo(a) : d_lo(b));
}

// np.clip(q, -c, c) for c >= 0 and finite q: one DSETP on |q| and two SEL against +-c
__device__ __forceinline__ double bpw_clip(double q, double c, uint32_t c_hi, uint32_t c_lo)
{
    const uint32_t sc_hi = (d_hi(q) & 0x80000000u) | c_hi;
    uint32_t hi, lo;
    asm("{\n\t.reg .pred p;\n\t.reg .f64 a;\n\tabs.f64 a, %2;\n\tsetp.gt.f64 p, a, %3;\n\t"
        "selp.b32 %0, %4, %5, p;\n\tselp.b32 %1, %6, %7, p;\n\t}"
        : "=r"(hi), "=r"(lo) : "d"(q), "d"(c), "r"(sc_hi), "r"(d_hi(q)), "r"(c_lo), "r"(d_lo(q)));
    return d_make(hi, lo);
}

// Accessors of the per-warp buffers.  `off`: byte offset of an 8-byte element (twice the table entry, which addresses 4-byte
// elements).
struct BPW64Mem {
    __device__ static __forceinline__ uint32_t scale(uint32_t off) { return 2u * off; }
    __device__ static __forceinline__ double ld(const unsigned char *base, uint32_t off) { return *reinterpret_cast<const double *>(base + off); }
    __device__ static __forceinline__ void st(unsigned char *base, uint32_t off, double v) { *reinterpret_cast<double *>(base + off) = v; }
    // element `row * 32 + lane` (the lane's own column)
    __device__ static __forceinline__ double ld_own(const unsigned char *base, int row, int lane) { return ld(base, 8u * (uint32_t)(row * 32 + lane)); }
    __device__ static __forceinline__ void st_own(unsigned char *base, int row, int lane, double v) { st(base, 8u * (uint32_t)(row * 32 + lane), v); }
};

// ZSC: zero-syndrome shortcut compiled in (see bp_warp_kernel.cuh; chosen by the launcher at low error rates)
template <int CPL, int VPL, int RW, bool TWO, int VAR = 0, bool ZSC = false>
__global__ void __launch_bounds__(BPW64_WARPS * 32, (CPL * RW > 18) ? 2 : 3)
bp_warp_kernel_f64(const BPParams P, const BPWarpTables W)
{
    const int n = P.g.n, WN = P.g.WN, WM = P.g.WM;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned FULL = 0xffffffffu;
    constexpr int RROWS = 3 * VPL + 1, VROWS = VPL + 1;                // message planes + dump row; posteriors + inf row
    typedef BPW64Mem RM;
    typedef BPW64Mem VM;
    extern __shared__ __align__(16) unsigned char smem[];
    unsigned char *Rbuf = smem + bp_warp64_smem_per_warp(VPL) * warp;
    unsigned char *Vbuf = Rbuf + 8 * 32 * RROWS;

    // ---- per-lane tables into registers (BYTE offsets into the message / posterior buffers) -------------
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
            vidx[i][k] = VM::scale(W.vidx[(i * RW + k) * 32 + lane]);
            sidx[i][k] = RM::scale(W.sidx[(i * RW + k) * 32 + lane]);
        }
    }
#pragma unroll
    for (int r = 0; r < RROWS; ++r) RM::st_own(Rbuf, r, lane, 0.0);      // columns of padding positions stay zero for ever
    VM::st_own(Vbuf, VPL, lane, CUDART_INF);
    const double prior0 = reinterpret_cast<const double *>(P.prior)[0] + 0.0;

    const double alpha = P.alpha, damp = P.damping, omd = P.one_minus_damping, clipv = P.clip;
    const uint32_t clip_hi = d_hi(clipv), clip_lo = d_lo(clipv);
    const int max_iter = P.max_iter;
    unsigned long long iter_sum = 0;

    // Shot indices (global cursor, BPW_GRAB shots per atomic) and syndrome words are fetched two shots ahead, as in the
    // float32 kernel.
    auto load_synd = [&](long long sh, uint32_t (&w)[CPL]) {
#pragma unroll
        for (int i = 0; i < CPL; ++i) w[i] = (sh < P.B && cinfo[i] != 0xffffffffu) ? P.synd[(size_t)sh * WM + (cinfo[i] >> 5)] : 0u;
    };
    unsigned long long s0 = 0;
    if (lane == 0) s0 = atomicAdd(P.cursor, (unsigned long long)BPW_GRAB);
    long long shot = (long long)__shfl_sync(FULL, s0, 0);
    long long next_shot = shot + 1, grp_next = shot + 2, grp_end = shot + BPW_GRAB;
    uint32_t sw[CPL], swn[CPL];
    load_synd(shot, sw);
    load_synd(next_shot, swn);

    while (shot < P.B) {
        const bool need_grab = grp_next >= grp_end;            // (warp-uniform) shot k + 2 starts a new group
        if (need_grab && lane == 0) s0 = atomicAdd(P.cursor, (unsigned long long)BPW_GRAB);   // consumed after iteration 0
        long long next2_shot = 0;
        uint32_t sbit[CPL];                      // syndrome bit of each owned check, at the sign-bit position
#pragma unroll
        for (int i = 0; i < CPL; ++i) sbit[i] = ((sw[i] >> (cinfo[i] & 31u)) & 1u) << 31;
        // all-zero syndrome, positive priors: the reference returns the all-zero correction at its first check (decoding.py:69-73)
        if (ZSC && P.zero_ok && !(P.llr != nullptr && P.llr_mode == LLR_ALL)) {
            uint32_t anyb = 0;
#pragma unroll
            for (int i = 0; i < CPL; ++i) anyb |= sbit[i];
            if (__all_sync(FULL, anyb == 0)) {
                if (need_grab) {
                    grp_next = (long long)__shfl_sync(FULL, s0, 0);
                    grp_end = grp_next + BPW_GRAB;
                }
                next2_shot = grp_next++;
#pragma unroll
                for (int i = 0; i < CPL; ++i) sw[i] = swn[i];
                load_synd(next2_shot, swn);
                if (lane < WN) P.hard[(size_t)shot * WN + lane] = 0u;
                if (lane == 0) {
                    P.conv[shot] = 1;
                    if (P.iters) P.iters[shot] = 0;
                }
                shot = next_shot;
                next_shot = next2_shot;
                continue;
            }
        }
        // Q = where(mask, prior, 0) (decoding.py:21): one value when the prior is uniform, else publish the priors and
        // gather them along the edges
        double Q[CPL][RW];
        if (P.prior_uniform) {
#pragma unroll
            for (int i = 0; i < CPL; ++i)
#pragma unroll
                for (int k = 0; k < RW; ++k) Q[i][k] = prior0;
        } else {
            __syncwarp();
#pragma unroll
            for (int i = 0; i < VPL; ++i) VM::st_own(Vbuf, i, lane, prior[i]);
            __syncwarp();
#pragma unroll
            for (int i = 0; i < CPL; ++i)
#pragma unroll
                for (int k = 0; k < RW; ++k) Q[i][k] = VM::ld(Vbuf, vidx[i][k]);   // (every check has RW edges here: only whole
                                                                                     //  padding lanes read the +inf row)
        }

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
                double pre[RW], suf[RW];
                if constexpr (VAR == 0) {
                    pre[1] = Q[i][0];
                    suf[RW - 2] = Q[i][RW - 1];
#pragma unroll
                    for (int k = 2; k < RW; ++k) pre[k] = bpw_absmin(pre[k - 1], Q[i][k - 1]);
#pragma unroll
                    for (int k = RW - 3; k >= 0; --k) suf[k] = bpw_absmin(suf[k + 1], Q[i][k + 1]);
                    uint32_t sgall = sbit[i];
#pragma unroll
                    for (int k = 0; k < RW; ++k) sgall ^= d_hi(Q[i][k]);
#pragma unroll
                    for (int k = 0; k < RW; ++k) {
                        const double o = (k == 0) ? suf[0] : (k == RW - 1) ? pre[RW - 1] : bpw_absmin(pre[k], suf[k]);
                        const double am = __dmul_rn(alpha, fabs(o));                                     // :55
                        R[i][k] = d_make(d_hi(am) ^ ((sgall ^ d_hi(Q[i][k])) & 0x80000000u), d_lo(am));
                    }
                } else {
                    // beliefPropagation.py:114-126 / decoding.py:157-166
                    double t[RW], o[RW];
                    uint32_t tmin = 0x7fffffffu;
#pragma unroll
                    for (int k = 0; k < RW; ++k) {
                        t[k] = spm_tanh_half(Q[i][k]);
                        tmin = min(tmin, d_hi(t[k]) & 0x7fffffffu);
                    }
                    pre[1] = t[0];
                    suf[RW - 2] = t[RW - 1];
#pragma unroll
                    for (int k = 2; k < RW; ++k) pre[k] = __dmul_rn(pre[k - 1], t[k - 1]);
#pragma unroll
                    for (int k = RW - 3; k >= 0; --k) suf[k] = __dmul_rn(suf[k + 1], t[k + 1]);
#pragma unroll
                    for (int k = 0; k < RW; ++k) o[k] = (k == 0) ? suf[0] : (k == RW - 1) ? pre[RW - 1] : __dmul_rn(pre[k], suf[k]);
                    if (tmin <= 0x3cd203afu) {                       // (high word of 1e-15) some |t| may be below the guard: the
                        const double prod = __dmul_rn(pre[RW - 1], t[RW - 1]);   // reference's row product / tanh_Q_safe, literally
#pragma unroll
                        for (int k = 0; k < RW; ++k) o[k] = __ddiv_rn(prod, fabs(t[k]) < 1e-15 ? 1e-15 : t[k]);
                    }
#pragma unroll
                    for (int k = 0; k < RW; ++k) {
                        const double x = d_make(d_hi(o[k]) ^ sbit[i], d_lo(o[k]));                       // * syndrome_sign
                        const double r = spm_2atanh_clipped(x);
                        R[i][k] = (VAR == 2) ? __dmul_rn(r, alpha) : r;                                  // decoding.py:171
                    }
                }
#pragma unroll
                for (int k = 0; k < RW; ++k) {
                    if (TWO && iter == 0) RM::st(Rbuf, RM::scale(__ldg(W.sidx0 + (i * RW + k) * 32 + lane)), R[i][k]);
                    else RM::st(Rbuf, sidx[i][k], R[i][k]);       // (padding lanes write garbage into the dump row)
                }
            }
            __syncwarp();

            // ================= vertical step: posteriors of the owned variables =====================
            const bool last = (iter == max_iter - 1);
#pragma unroll
            for (int i = 0; i < VPL; ++i) {
                const double r0 = RM::ld_own(Rbuf, 0 * VPL + i, lane), r1 = RM::ld_own(Rbuf, 1 * VPL + i, lane),
                             r2 = RM::ld_own(Rbuf, 2 * VPL + i, lane);
                VM::st_own(Vbuf, i, lane, __dadd_rn(__dadd_rn(__dadd_rn(r0, r1), r2), prior[i]));    // :61-62
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
                    const double val = VM::ld(Vbuf, vidx[i][k]);
                    par ^= d_hi(val);   // sign bit == hard decision (a sum with a canonical prior is never -0.0)
                    double qn = __dsub_rn(val, R[i][k]);                                          // :63
                    if constexpr (VAR == 1) {
                        Q[i][k] = qn;                                                             // beliefPropagation.py:133
                    } else {
                        qn = __dadd_rn(__dmul_rn(damp, qn), __dmul_rn(omd, Q[i][k]));             // :65 (three roundings, like NumPy)
                        Q[i][k] = bpw_clip(qn, clipv, clip_hi, clip_lo);                          // :66
                    }
                }
                ok = ok && (cinfo[i] == 0xffffffffu || (int)par >= 0);
            }
            conv = __all_sync(FULL, ok);
            if (iter == 0) {
                if (need_grab) {
                    grp_next = (long long)__shfl_sync(FULL, s0, 0);
                    grp_end = grp_next + BPW_GRAB;
                }
                next2_shot = grp_next++;
#pragma unroll
                for (int i = 0; i < CPL; ++i) sw[i] = swn[i];         // (issued one shot ago: landed)
                load_synd(next2_shot, swn);
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
                const double val = valid ? VM::ld(Vbuf, VM::scale(__ldg(W.vpos + i * 32 + lane))) : 0.0;
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
        next_shot = next2_shot;
    }
    if (P.iter_total && lane == 0 && iter_sum) atomicAdd(P.iter_total, iter_sum);
}

}  // namespace qldpc
