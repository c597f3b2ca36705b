// Launchers of the warp-per-shot float32 min-sum kernels (bp_warp_kernel.cuh) and the dispatch to its other variants.
#include "capi_internal.h"

template <int CPL, int VPL, bool TWO, int VAR, bool ZSC = false>
static cudaError_t launch_bp_warp_inst3(const qldpc_code *c, const BPParams &P, const BPGeom &G, cudaStream_t st)
{
    auto kern = bp_warp_kernel<CPL, VPL, 6, TWO, VAR, ZSC>;
    if (G.smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G.smem);
        if (e != cudaSuccess) return e;
    }
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, G.threads, G.smem);
    const long long grid = std::max<long long>(1, std::min<long long>((long long)c->num_sms * std::max(1, occ), (P.B + BPW_WARPS - 1) / BPW_WARPS));
    kern<<<(int)grid, G.threads, G.smem, st>>>(P, c->wtab);
    return cudaGetLastError();
}

template <int CPL, int VPL>
static cudaError_t launch_ms(const qldpc_code *c, const BPParams &P, const BPGeom &G, cudaStream_t st)
{
    // The iteration-0 addition order only matters for non-uniform priors: with one prior value every message of iteration 0
    // has the same magnitude a = alpha * prior, and (+-a +- a) +- a rounds the same in any order (2a and 0 are exact).
    const bool two = c->two_tables && !P.prior_uniform;
    if (P.zero_ok == 2)             // many all-zero syndromes expected (low error rates): the instantiation with the shortcut
        return two ? launch_bp_warp_inst3<CPL, VPL, true, 0, true>(c, P, G, st) : launch_bp_warp_inst3<CPL, VPL, false, 0, true>(c, P, G, st);
    return two ? launch_bp_warp_inst3<CPL, VPL, true, 0>(c, P, G, st) : launch_bp_warp_inst3<CPL, VPL, false, 0>(c, P, G, st);
}

#define QLDPC_WARP_SHAPES(F)                                                  \
    if (c->WM == 2 && c->WN == 3) return F<2, 3>(c, P, G, st);                \
    if (c->WM == 2 && c->WN == 4) return F<2, 4>(c, P, G, st);                \
    if (c->WM == 3 && c->WN == 5) return F<3, 5>(c, P, G, st);                \
    if (c->WM == 5 && c->WN == 9) return F<5, 9>(c, P, G, st);                \
    return cudaErrorInvalidValue

// G.warp_var: 0 min-sum, 1 sum-product, 2 symmetric sum-product (float32); 3 float64 min-sum (the bit-exact parity mode),
// 4 / 5 float64 sum-product / symmetric sum-product
cudaError_t launch_bp_warp(const qldpc_code *c, const BPParams &P, const BPGeom &G, cudaStream_t st)
{
    if (G.warp_var == 3) return launch_bp_warp_f64(c, P, G, st);
    if (G.warp_var >= 4) return launch_bp_warp_f64_sp(c, P, G, st);
    if (G.warp_var != 0) return launch_bp_warp_sp(c, P, G, st);
    QLDPC_WARP_SHAPES(launch_ms);
}
