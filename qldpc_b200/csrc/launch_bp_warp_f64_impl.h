// Launchers of the warp-per-shot float64 kernels (bp_warp_kernel_f64.cuh): min-sum, the bit-exact parity mode 
// (launch_bp_warp_f64.cu, QLDPC_F64_VAR = 0) and the two sum-product variants (launch_bp_warp_f64_sp.cu, QLDPC_F64_VAR = 1):
// one translation unit each, so that they compile in parallel.
#include "capi_internal.h"

#if !defined(QLDPC_F64_VAR) || !defined(QLDPC_F64_ENTRY)
#error "define QLDPC_F64_VAR (0: min-sum, 1: the sum-product variants) and QLDPC_F64_ENTRY before including this file"
#endif

template <int CPL, int VPL, bool TWO, int VAR, bool ZSC = false>
static cudaError_t launch_f64_inst(const qldpc_code *c, const BPParams &P, const BPGeom &G, cudaStream_t st)
{
    auto kern = bp_warp_kernel_f64<CPL, VPL, 6, TWO, VAR, ZSC>;
    if (G.smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G.smem);
        if (e != cudaSuccess) return e;
    }
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, G.threads, G.smem);
    const long long grid = std::max<long long>(1, std::min<long long>((long long)c->num_sms * std::max(1, occ), (P.B + BPW64_WARPS - 1) / BPW64_WARPS));
    kern<<<(int)grid, G.threads, G.smem, st>>>(P, c->wtab64);      // (labelling for 16-lane conflict domains)
    return cudaGetLastError();
}

template <int CPL, int VPL>
static cudaError_t launch_f64(const qldpc_code *c, const BPParams &P, const BPGeom &G, cudaStream_t st)
{
    // (uniform prior: the iteration-0 addition order cannot change a bit -- see launch_bp_warp.cu)
    // (sum-product too: every check of these codes has the same weight, so the messages of iteration 0 share one magnitude)
    const bool two = c->two_tables && !P.prior_uniform;
#if QLDPC_F64_VAR == 0
    if (P.zero_ok == 2)             // many all-zero syndromes expected (low error rates): the instantiation with the shortcut
        return two ? launch_f64_inst<CPL, VPL, true, 0, true>(c, P, G, st) : launch_f64_inst<CPL, VPL, false, 0, true>(c, P, G, st);
    return two ? launch_f64_inst<CPL, VPL, true, 0>(c, P, G, st) : launch_f64_inst<CPL, VPL, false, 0>(c, P, G, st);
#else
    if (P.zero_ok == 2) {
        if (G.warp_var == 4)
            return two ? launch_f64_inst<CPL, VPL, true, 1, true>(c, P, G, st) : launch_f64_inst<CPL, VPL, false, 1, true>(c, P, G, st);
        return two ? launch_f64_inst<CPL, VPL, true, 2, true>(c, P, G, st) : launch_f64_inst<CPL, VPL, false, 2, true>(c, P, G, st);
    }
    if (G.warp_var == 4)
        return two ? launch_f64_inst<CPL, VPL, true, 1>(c, P, G, st) : launch_f64_inst<CPL, VPL, false, 1>(c, P, G, st);
    return two ? launch_f64_inst<CPL, VPL, true, 2>(c, P, G, st) : launch_f64_inst<CPL, VPL, false, 2>(c, P, G, st);
#endif
}

#define QLDPC_WARP_SHAPES(F)                                                  \
    if (c->WM == 2 && c->WN == 3) return F<2, 3>(c, P, G, st);                \
    if (c->WM == 2 && c->WN == 4) return F<2, 4>(c, P, G, st);                \
    if (c->WM == 3 && c->WN == 5) return F<3, 5>(c, P, G, st);                \
    if (c->WM == 5 && c->WN == 9) return F<5, 9>(c, P, G, st);                \
    return cudaErrorInvalidValue

cudaError_t QLDPC_F64_ENTRY(const qldpc_code *c, const BPParams &P, const BPGeom &G, cudaStream_t st)
{
    QLDPC_WARP_SHAPES(launch_f64);
}
