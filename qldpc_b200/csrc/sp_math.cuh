// Float64 tanh(q / 2) and 2 atanh(x) for the sum-product check update (beliefPropagation.py:114-126,
// rework/decoding.py:157-166), written for the FP64 pipe of sm_100a.
//
// The CUDA math library's tanh / atanh / IEEE division cost 34 + 76 + 20 FP64 instructions per edge plus ~100 integer
// instructions of special-case handling (denormals, infinities, slow division paths) that the check update never needs:
// its arguments are bounded (|x| <= 0.9999999 by the reference's own clip; tanh saturates to 1.0 above |q| = 38).  These
// versions take 23 + 19 FP64 instructions, one MUFU each, no branches:
//   * tanh(q/2) = -E / (2 + E) with E = expm1(-|q|) = 2^k expm1(r) + (2^k - 1): no cancellation for small |q| (k = 0 gives
//     E = expm1(r) with full relative accuracy);
//   * 2 atanh(x) = log((1 + x) / (1 - x)) = k ln 2 + 2 atanh(s), s = (a - b') / (a + b') with a = 1 + |x|, b' = 2^k (1 - |x|)
//     and k chosen from the exponent / leading mantissa bits so that a / b' lies in [0.75, 1.5) -- the quotient (1 + x) /
//     (1 - x) itself is never formed, one division instead of two; for k = 0 (|x| < 0.2) s = |x| exactly;
//   * divisions: MUFU.RCP64H seed (20 bits) + one cubic Newton step (2^-60) + the product.
// Relative error <= 4e-16 on both (tests/test_host.py checks them against 50-digit references; tests/cpp/sp_math_check.cpp is
// the host build of this file).  The reference's NumPy functions are themselves only accurate to ~1 ulp and the parity
// bar for sum-product is 1e-4 relative (north_star); the float64 kernels are held to 1e-7 on every golden shot.
//
// The header compiles as plain C++ (host: fma() and `/`) for those tests.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define SPM_FN __host__ __device__ __forceinline__
#else
#define SPM_FN static inline
#endif

namespace qldpc {

SPM_FN uint32_t spm_hi(double x)
{
#if defined(__CUDA_ARCH__)
    return (uint32_t)__double2hiint(x);
#else
    uint64_t u; memcpy(&u, &x, 8); return (uint32_t)(u >> 32);
#endif
}
SPM_FN uint32_t spm_lo(double x)
{
#if defined(__CUDA_ARCH__)
    return (uint32_t)__double2loint(x);
#else
    uint64_t u; memcpy(&u, &x, 8); return (uint32_t)u;
#endif
}
SPM_FN double spm_make(uint32_t hi, uint32_t lo)
{
#if defined(__CUDA_ARCH__)
    return __hiloint2double((int)hi, (int)lo);
#else
    uint64_t u = ((uint64_t)hi << 32) | lo; double x; memcpy(&x, &u, 8); return x;
#endif
}

// a / b for a normal b > 0 far from the ends of the exponent range; relative error <= 2^-52
SPM_FN double spm_div(double a, double b)
{
#if defined(__CUDA_ARCH__)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
    const double e = fma(-b, r, 1.0);
    const double t = fma(e, e, e);
    r = fma(r, t, r);
    return a * r;
#elif defined(SPM_EMULATE_RCP)                               // host model of the device sequence (tests)
    double r = 1.0 / b;
    uint64_t u; memcpy(&u, &r, 8); u &= 0xffffffff00000000ull; memcpy(&r, &u, 8);     // MUFU.RCP64H fills the upper word only
    const double e = fma(-b, r, 1.0);
    const double t = fma(e, e, e);
    r = fma(r, t, r);
    return a * r;
#else
    return a / b;
#endif
}

// tanh(q / 2)
SPM_FN double spm_tanh_half(double q)
{
    const double L2E = 0x1.71547652b82fep+0, LN2_HI = 0x1.62e42fefa39efp-1, LN2_LO = 0x1.abc9e3b39803fp-56;
    const double MAGIC = 0x1.8p+52;
    // |q| clamped near 80 on the high word (integer min: off the FP64 pipe): tanh(40) rounds to 1, and 2^k stays normal
    const uint32_t qh = spm_hi(q) & 0x7fffffffu;
    const double ax = spm_make(qh < 0x40540000u ? qh : 0x40540000u, spm_lo(q));
    const double km = fma(-ax, L2E, MAGIC);                  // the low word of km is k = rint(-|q| / ln 2) (two's complement)
    const double kd = km - MAGIC;
    double r = fma(kd, -LN2_HI, -ax);
    r = fma(kd, -LN2_LO, r);                                 // |r| <= ln 2 / 2
    double p = 0x1.af4dea43cc39ep-26;                        // expm1(r) = r + r^2 P(r), near-minimax, 3.6e-17 relative
    p = fma(p, r, 0x1.2891861bf46efp-22);
    p = fma(p, r, 0x1.71de02312a642p-19);
    p = fma(p, r, 0x1.a019b90c7771cp-16);
    p = fma(p, r, 0x1.a01a01abe824dp-13);
    p = fma(p, r, 0x1.6c16c1788c0eep-10);
    p = fma(p, r, 0x1.11111111100dbp-7);
    p = fma(p, r, 0x1.5555555553d62p-5);
    p = fma(p, r, 0x1.5555555555557p-3);
    p = fma(p, r, 0x1.0000000000001p-1);
    p = fma(r * r, p, r);
    const double s = spm_make((uint32_t)(1023 + (int)spm_lo(km)) << 20, 0u);      // 2^k
    const double e = fma(s, p, s - 1.0);                     // expm1(-|q|) in (-1, 0]
    const double t = spm_div(fabs(e), 2.0 + e);              // (e <= 0; |.| keeps tanh(+0) = +0)
    return spm_make(spm_hi(t) | (spm_hi(q) & 0x80000000u), spm_lo(t));
}

// 2 atanh(clip(x, -0.9999999, 0.9999999))
SPM_FN double spm_2atanh_clipped(double x)
{
    const double CLIP = 0.9999999, LN2 = 0x1.62e42fefa39efp-1;
    // np.clip(|x|, CLIP): the bit patterns of non-negative doubles order like the values (integer compare: off the FP64 pipe)
    const uint64_t xb = ((uint64_t)(spm_hi(x) & 0x7fffffffu) << 32) | spm_lo(x);
    const uint64_t cb = ((uint64_t)spm_hi(CLIP) << 32) | spm_lo(CLIP);
    const bool below = xb < cb;
    const double ax = spm_make(below ? (uint32_t)(xb >> 32) : (uint32_t)(cb >> 32), below ? (uint32_t)xb : (uint32_t)cb);
    const double a = 1.0 + ax, b = 1.0 - ax;                 // [1, 2), [1e-7, 1]
    const uint32_t ha = spm_hi(a), hb = spm_hi(b);
    const uint32_t fa = ha & 0xfffffu, fb = hb & 0xfffffu;   // leading 20 fraction bits
    // j: a / mantissa(b) >= 1.5 -> +1, < 0.75 -> -1 (decided on the truncated mantissas: the polynomial's range has the slack)
    const int j = (2u * fa >= (1u << 20) + 3u * fb) ? 1 : (((1u << 20) + 4u * fa < 3u * fb) ? -1 : 0);
    const int k = j + 1023 - (int)(hb >> 20);                // >= 0; y = (1 + |x|) / (1 - |x|) = 2^k m, m in [0.75, 1.5)
    const double bs = spm_make(hb + ((uint32_t)k << 20), spm_lo(b));
    const bool k0 = (k == 0);                                // |x| < 0.2: s = |x| exactly
    const double num = k0 ? ax : a - bs, den = k0 ? 1.0 : a + bs;
    const double s = spm_div(num, den);
    const double z = s * s;
    double p = 0x1.3657e05f4f6ecp-3;                         // 2 atanh(s) = 2 s + s z P(z), |s| <= 0.2, 6.5e-17 relative
    p = fma(p, z, 0x1.38ee287226b79p-3);
    p = fma(p, z, 0x1.746c9439449d6p-3);
    p = fma(p, z, 0x1.c71c38b631aecp-3);
    p = fma(p, z, 0x1.2492495680d42p-2);
    p = fma(p, z, 0x1.9999999978e1cp-2);
    p = fma(p, z, 0x1.5555555555571p-1);
    double r = fma(s * z, p, s + s);
    r = fma((double)k, LN2, r);
    return spm_make(spm_hi(r) | (spm_hi(x) & 0x80000000u), spm_lo(r));
}

}  // namespace qldpc
