// Min-sum BP for code-capacity check matrices (BB codes) on sm_100a: ONE WARP PER SHOT, the variable-to-check
// messages live in REGISTERS of the lane that owns their check.
//
// Same arithmetic, in the same order, as bp_decode_kernel<float, VAR_MIN_SUM> / bp_tiled_kernel (reference
// rework/decoding.py:5-75): results are bit-identical.  Different mapping, chosen after the ncu profile of the tiled
// kernel (profiles/r1e: shared-memory wavefronts at 84 % of peak, issue slots 76 % busy):
//   * a lane owns up to CPL checks and keeps their RW incoming messages Q[i][k] in registers, so the whole check pass
//     is lane-local: a prefix/suffix chain of FMNMX.XORSIGN (min of magnitudes, xor of signs in one instruction) gives
//     the RW outgoing messages R, which the lane SCATTERS into the columns of their destination variables
//     (plane t = position of the message in the variable's addition order);
//   * a lane also owns up to VPL variables: it reads its own column (three planes, conflict-free, no index), adds in
//     the reference's order, adds the prior and publishes the posterior;
//   * each check-owner lane gathers the posteriors of its RW variables, updates Q in registers (damping against the Q it
//     still holds, clip) and xors the sign bits it just read: the syndrome test of the hard decision is lane-local,
//     followed by one __all_sync.
// Which lane owns what, and in which register slot an edge sits, is chosen by the host (bp_warp_layout.h) so that
// scatter and gather are bank-conflict free: CPL*RW + 4*VPL + CPL*RW wavefronts per shot-iteration (56 for
// [[144,12,12]]) and ~225 warp-instructions (the float32 adds / multiplies / FMAs of neighbouring edges are issued as packed
// pairs, FADD2 / FMUL2 / FFMA2), against ~150 wavefronts per shot and ~560 instructions in the tiled kernel.
// A warp decodes exactly one shot, retires it and fetches the next one from the global cursor (four shots per atomic, two
// shots ahead): no divergence.  The same kernel runs the reference's sum-product in the psi domain (VAR = 1, 2).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "bp_kernel.cuh"

namespace qldpc {

constexpr int BPW_WARPS = 8;            // warps per CTA
constexpr int BPW_GRAB = 4;             // shots a warp takes from the global cursor per atomic

__device__ __forceinline__ float ldb(const float *base, uint32_t byte_off)
{
    return *reinterpret_cast<const float *>(reinterpret_cast<const unsigned char *>(base) + byte_off);
}

// Packed float32 pairs (FADD2 / FMUL2 / FFMA2 on sm_100: two IEEE-rounded float32 operations per issue slot).  The kernel is
// bound by issue slots, not by the FMA pipe, so pairing the message updates of neighbouring edges is free throughput.
// Each component is rounded exactly like the scalar instruction: results do not change.
__device__ __forceinline__ void bpw_sub2(float a0, float a1, float b0, float b1, float &r0, float &r1)
{
    asm("{\n\t.reg .b64 a, b, c;\n\tmov.b64 a, {%2, %3};\n\tmov.b64 b, {%4, %5};\n\tsub.rn.f32x2 c, a, b;\n\tmov.b64 {%0, %1}, c;\n\t}"
        : "=f"(r0), "=f"(r1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
__device__ __forceinline__ void bpw_add2(float a0, float a1, float b0, float b1, float &r0, float &r1)
{
    asm("{\n\t.reg .b64 a, b, c;\n\tmov.b64 a, {%2, %3};\n\tmov.b64 b, {%4, %5};\n\tadd.rn.f32x2 c, a, b;\n\tmov.b64 {%0, %1}, c;\n\t}"
        : "=f"(r0), "=f"(r1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
__device__ __forceinline__ void bpw_mul2(float a0, float a1, float b0, float b1, float &r0, float &r1)
{
    asm("{\n\t.reg .b64 a, b, c;\n\tmov.b64 a, {%2, %3};\n\tmov.b64 b, {%4, %5};\n\tmul.rn.f32x2 c, a, b;\n\tmov.b64 {%0, %1}, c;\n\t}"
        : "=f"(r0), "=f"(r1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
__device__ __forceinline__ void bpw_fma2(float a0, float a1, float b0, float b1, float c0, float c1, float &r0, float &r1)
{
    asm("{\n\t.reg .b64 a, b, c, d;\n\tmov.b64 a, {%2, %3};\n\tmov.b64 b, {%4, %5};\n\tmov.b64 c, {%6, %7};\n\tfma.rn.f32x2 d, a, b, c;\n\tmov.b64 {%0, %1}, d;\n\t}"
        : "=f"(r0), "=f"(r1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1), "f"(c0), "f"(c1));
}

// per-warp shared memory: message planes [3][VPL][32] + one dump row [32] (padding lanes) + posteriors [VPL][32] + a row of
// +inf (what padding edge slots read as their "posterior")
__host__ __device__ inline size_t bp_warp_smem_per_warp(int VPL) { return 4 * (size_t)32 * (4 * VPL + 2); }

__device__ __forceinline__ void stb(float *base, uint32_t byte_off, float v)
{
    *reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(base) + byte_off) = v;
}

// FMNMX.XORSIGN: magnitude min(|a|, |b|), sign = sign(a) ^ sign(b).  A chain of these over a set of messages yields
// the min-sum check output (product of the signs, minimum of the magnitudes) in ONE ALU-pipe instruction per message.
__device__ __forceinline__ float bpw_xmin(float a, float b)
{
    float d;
    asm("min.xorsign.abs.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
    return d;
}

// psi(a) = -log(tanh(a / 2)) = log((1 + e^-a) / (1 - e^-a)) for a >= 0, its own inverse.  The sum-product check update in
// the psi domain: |R_k| = 2 atanh(prod_{j != k} tanh(|Q_j| / 2)) = psi(sum_{j != k} psi(|Q_j|)).  Unlike the tanh domain,
// float32 keeps its RELATIVE accuracy over the whole range here: near saturation the tanh product sits within 1e-7 of 1,
// below the float32 resolution, while psi of a large argument is a tiny number with a full mantissa.
//   a >= 1: u = e^-a <= 0.37, psi = 2 atanh(u) by its odd series (truncation < 5e-7 relative);
//   a <  1: 1 - e^-a = -expm1(-a) by its series (no cancellation), psi = log((1 + u) / (1 - u)) >= 0.77.
// psi(0) is clamped to psi at tanh = 1e-15, the reference's guard (beliefPropagation.py:122).
__device__ __forceinline__ float bpw_psi(float a)
{
    const float u = __expf(-a);
    const float u2 = u * u;
    float big = fmaf(u2, 1.f / 11.f, 1.f / 9.f);
    big = fmaf(u2, big, 1.f / 7.f);
    big = fmaf(u2, big, 1.f / 5.f);
    big = fmaf(u2, big, 1.f / 3.f);
    big = fmaf(u2, big, 1.f);
    big = 2.f * u * big;
    const float z = -fminf(a, 1.f);
    float e = fmaf(z, 1.f / 362880.f, 1.f / 40320.f);          // expm1(z) / z, degree 8 in z, |z| <= 1
    e = fmaf(z, e, 1.f / 5040.f);
    e = fmaf(z, e, 1.f / 720.f);
    e = fmaf(z, e, 1.f / 120.f);
    e = fmaf(z, e, 1.f / 24.f);
    e = fmaf(z, e, 1.f / 6.f);
    e = fmaf(z, e, 0.5f);
    e = fmaf(z, e, 1.f);
    const float d = -z * e;                                     // 1 - e^-a
    const float small = __logf(__fdividef(2.f - d, d));
    return fminf((a >= 1.f) ? big : small, 34.538776394910684f);
}

// Tables (global, built by the host -- bp_warp_layout.h -- for a labelling "position = slot * 32 + lane" of the checks
// and of the variables that it is free to choose):
//   sidx  [CPL*RW][32] byte offset in the message planes where edge slot k of the check at (i, lane) delivers its
//                      message in iterations >= 1 (plane t, column of the variable; the dump row for padding);
//                      sidx0: the same for iteration 0 (read only when TWO: the reference's NumPy reduction order
//                      differs between iteration 0 and the later ones, graph.py)
//   vidx  [CPL*RW][32] byte offset in the posterior buffer of the variable of edge slot k of the check at (i, lane)
//   cinfo [CPL][32]    index of the check at (i, lane) in H, 0xffffffff for padding
//   vorig [VPL][32]    index of the variable at (i, lane) in H, 0xffffffff for padding
//   vpos  [VPL][32]    byte offset in the posterior buffer of variable 32 i + lane of H
struct BPWarpTables {
    const uint32_t *sidx, *sidx0, *vidx, *cinfo, *vorig, *vpos;
};

// VAR: 0 = normalised / damped / clipped min-sum (rework/decoding.py:5-75), 1 = sum-product (beliefPropagation.py:88-144),
// 2 = sum-product with alpha, damping and clipping (rework/decoding.py:131-191); 1 and 2 in the psi domain (above).
// ZSC: the zero-syndrome shortcut below is compiled in (its own instantiation: it costs three registers, which the [[144,12,12]]
// kernel at its 128-register cap pays with a spill and 11 % at p = 0.05 -- the launcher picks it when the priors say that at least a
// tenth of the shots have an all-zero syndrome)
template <int CPL, int VPL, int RW, bool TWO, int VAR, bool ZSC = false>
__global__ void __launch_bounds__(BPW_WARPS * 32, (CPL * RW + VPL * 4 > 44) ? 1 : 2)
bp_warp_kernel(const BPParams P, const BPWarpTables W)
{
    const int n = P.g.n, WN = P.g.WN, WM = P.g.WM;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char smem[];
    float *Rbuf = reinterpret_cast<float *>(smem + bp_warp_smem_per_warp(VPL) * warp);     // [3][VPL][32] + dump row
    float *Vbuf = Rbuf + 32 * (3 * VPL + 1);

    // ---- per-lane tables into registers (BYTE offsets into the R / posterior buffers) -------------
    uint32_t sidx[CPL][RW], vidx[CPL][RW], cinfo[CPL];
    float prior[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const uint32_t v = W.vorig[i * 32 + lane];
        prior[i] = (v != 0xffffffffu) ? reinterpret_cast<const float *>(P.prior)[v] + 0.f : 0.f;    // (+ 0: a -0.0 prior becomes +0.0)
    }
#pragma unroll
    for (int i = 0; i < CPL; ++i) {
        cinfo[i] = W.cinfo[i * 32 + lane];
#pragma unroll
        for (int k = 0; k < RW; ++k) {
            vidx[i][k] = W.vidx[(i * RW + k) * 32 + lane];
            sidx[i][k] = W.sidx[(i * RW + k) * 32 + lane];
        }
    }
#pragma unroll
    for (int r = 0; r < 3 * VPL + 1; ++r) Rbuf[r * 32 + lane] = 0.f;      // columns of padding positions stay zero for ever
    Vbuf[VPL * 32 + lane] = CUDART_INF_F;
    const float prior0 = reinterpret_cast<const float *>(P.prior)[0] + 0.f;

    const float alpha = (float)P.alpha, damp = (float)P.damping, omd = (float)P.one_minus_damping, clipv = (float)P.clip;
    const int max_iter = P.max_iter;
    unsigned long long iter_sum = 0;

    // Shot indices (global cursor) and syndrome words are fetched TWO shots ahead: the atomic of shot k + 2 is issued at the
    // top of shot k and read after its first iteration, when the syndrome loads of k + 2 are issued; they are consumed at
    // the top of shot k + 2.  At low error rates a shot lasts one or two iterations, less than a global-memory round trip.
    auto load_synd = [&](long long sh, uint32_t (&w)[CPL]) {
#pragma unroll
        for (int i = 0; i < CPL; ++i) w[i] = (sh < P.B && cinfo[i] != 0xffffffffu) ? P.synd[(size_t)sh * WM + (cinfo[i] >> 5)] : 0u;
    };
    // (the cursor is advanced BPW_GRAB shots at a time: one atomic per shot on a single address saturates the L2 at about
    //  1.4e9 shots/s, which short shots -- low error rates -- reach)
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
        uint32_t sbit[CPL];                      // syndrome bit of each owned check, moved to the sign-bit position
        float salpha[CPL];                       // (-1)^s * alpha
#pragma unroll
        for (int i = 0; i < CPL; ++i) {
            sbit[i] = ((sw[i] >> (cinfo[i] & 31u)) & 1u) << 31;
            salpha[i] = __uint_as_float(__float_as_uint(VAR == 1 ? 1.f : alpha) ^ sbit[i]);
        }
        // All-zero syndrome (a quarter to a third of the shots at p = 0.01): with positive priors every message and posterior of
        // iteration 0 is positive, the hard decision is all-zero and reproduces the syndrome -- the reference returns at its first
        // check (decoding.py:69-73).  The shot is retired with exactly those outputs and costs only its bookkeeping.
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
        float Q[CPL][RW];
        if (P.prior_uniform) {
#pragma unroll
            for (int i = 0; i < CPL; ++i)
#pragma unroll
                for (int k = 0; k < RW; ++k) Q[i][k] = prior0;
        } else {
            __syncwarp();
#pragma unroll
            for (int i = 0; i < VPL; ++i) Vbuf[i * 32 + lane] = prior[i] + 0.f;
            __syncwarp();
#pragma unroll
            for (int i = 0; i < CPL; ++i)
#pragma unroll
                for (int k = 0; k < RW; ++k) Q[i][k] = ldb(Vbuf, vidx[i][k]);  // (every check has RW edges here: only whole padding
                                                                                 //  lanes read the +inf row, and nothing reads them)
        }

        int iter = 0;
        bool conv = false;
        for (;; ++iter) {
            // ================= horizontal step (lane-local) ========================================
            // R[k] = alpha * (-1)^s * prod_{j != k} sign(Q[j]) * min_{j != k} |Q[j]| (decoding.py:41-55).  "All but k" is
            // prefix (x) suffix of the xorsign-min; the minimum over the others IS min1, or min2 at the arg-min (ties
            // included), so the values equal the reference's where(|Q| == min1, min2, min1) selection exactly.
            float R[CPL][RW];
#pragma unroll
            for (int i = 0; i < CPL; ++i) {
                float pre[RW], suf[RW];
                if (VAR == 0) {
                    pre[1] = Q[i][0];
                    suf[RW - 2] = Q[i][RW - 1];
#pragma unroll
                    for (int k = 2; k < RW; ++k) pre[k] = bpw_xmin(pre[k - 1], Q[i][k - 1]);
#pragma unroll
                    for (int k = RW - 3; k >= 0; --k) suf[k] = bpw_xmin(suf[k + 1], Q[i][k + 1]);
                } else {
                    // psi domain: "all but k" = prefix + suffix SUMS of psi(|Q|) (no subtraction: no cancellation)
                    float ps[RW];
#pragma unroll
                    for (int k = 0; k < RW; ++k) ps[k] = bpw_psi(fabsf(Q[i][k]));
                    pre[1] = ps[0];
                    suf[RW - 2] = ps[RW - 1];
#pragma unroll
                    for (int k = 2; k < RW; ++k) pre[k] = pre[k - 1] + ps[k - 1];
#pragma unroll
                    for (int k = RW - 3; k >= 0; --k) suf[k] = suf[k + 1] + ps[k + 1];
                }
                uint32_t sgall = 0;                                    // xor of all sign bits (sum-product)
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
                        // 2 atanh(clip(prod, 0.9999999)) (beliefPropagation.py:125-126): saturates at 16.81
                        const float mag = fminf(bpw_psi(sk), 16.811242831518264f);
                        o[k] = __uint_as_float(__float_as_uint(mag) | ((sgall ^ __float_as_uint(Q[i][k])) & 0x80000000u));
                    }
                }
#pragma unroll
                for (int k = 0; k < RW; k += 2) {
                    // (+-alpha) * (+-magnitude): same rounding as alpha * magnitude; (-1)^s only for plain sum-product
                    bpw_mul2(o[k], o[k + 1], salpha[i], salpha[i], R[i][k], R[i][k + 1]);
#pragma unroll
                    for (int kk = k; kk < k + 2; ++kk) {
                        if (TWO && iter == 0) stb(Rbuf, __ldg(W.sidx0 + (i * RW + kk) * 32 + lane), R[i][kk]);
                        else stb(Rbuf, sidx[i][kk], R[i][kk]);       // (padding lanes write garbage into the dump row)
                    }
                }
            }
            __syncwarp();

            // ================= vertical step: posteriors of the owned variables =====================
            const bool last = (iter == max_iter - 1);
#pragma unroll
            for (int i = 0; i + 1 < VPL; i += 2) {                                                  // two slots per packed add
                float a0 = Rbuf[(0 * VPL + i) * 32 + lane], a1 = Rbuf[(0 * VPL + i + 1) * 32 + lane];
                bpw_add2(a0, a1, Rbuf[(1 * VPL + i) * 32 + lane], Rbuf[(1 * VPL + i + 1) * 32 + lane], a0, a1);
                bpw_add2(a0, a1, Rbuf[(2 * VPL + i) * 32 + lane], Rbuf[(2 * VPL + i + 1) * 32 + lane], a0, a1);
                bpw_add2(a0, a1, prior[i], prior[i + 1], a0, a1);                                   // :61-62
                Vbuf[i * 32 + lane] = a0;
                Vbuf[(i + 1) * 32 + lane] = a1;
            }
            if (VPL & 1) {
                constexpr int i = VPL - 1;
                const float r0 = Rbuf[(0 * VPL + i) * 32 + lane], r1 = Rbuf[(1 * VPL + i) * 32 + lane], r2 = Rbuf[(2 * VPL + i) * 32 + lane];
                Vbuf[i * 32 + lane] = __fadd_rn(__fadd_rn(__fadd_rn(r0, r1), r2), prior[i]);
            }
            __syncwarp();

            // ================= Q update in registers + syndrome of the hard decision =================
            // The check is satisfied by the hard decisions iff the xor of the posterior sign bits equals its syndrome bit
            // (3-input LOP3s: the ALU pipe has room since the check pass moved to FMNMX.XORSIGN).
            bool ok = true;
#pragma unroll
            for (int i = 0; i < CPL; ++i) {
                uint32_t par = sbit[i];
                static_assert(RW % 2 == 0, "edge slots are updated in pairs");
#pragma unroll
                for (int k = 0; k < RW; k += 2) {
                    const float v0 = ldb(Vbuf, vidx[i][k]), v1 = ldb(Vbuf, vidx[i][k + 1]);
                    par ^= __float_as_uint(v0) ^ __float_as_uint(v1);   // sign bit == hard decision (a sum with a non-zero prior is never -0.0)
                    float q0, q1;
                    bpw_sub2(v0, v1, R[i][k], R[i][k + 1], q0, q1);                              // :63
                    if (VAR != 1) {                                                               // (plain sum-product: Q = values - R)
                        float t0, t1;
                        bpw_mul2(omd, omd, Q[i][k], Q[i][k + 1], t0, t1);                         // :65, same roundings as bp_damp(float)
                        bpw_fma2(damp, damp, q0, q1, t0, t1, q0, q1);
                        q0 = bpw_xmin(q0, clipv);     // np.clip(q, -c, c) for c >= 0: sign(q) * min(|q|, c), ONE FMNMX.XORSIGN                                     // :66
                        q1 = bpw_xmin(q1, clipv);
                    }
                    Q[i][k] = q0;
                    Q[i][k + 1] = q1;
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
                const float val = valid ? ldb(Vbuf, __ldg(W.vpos + i * 32 + lane)) : 0.f;
                const uint32_t w = __ballot_sync(FULL, valid && (val < 0.f));
                if (lane == i) myw = w;
                if (wr_llr && valid) reinterpret_cast<float *>(P.llr)[(size_t)shot * n + lane + 32 * i] = val;
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
