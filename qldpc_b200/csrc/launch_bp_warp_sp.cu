// Launchers of the warp-per-shot float32 sum-product kernels (psi domain, bp_warp_kernel.cuh VAR = 1, 2).
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
static cudaError_t launch_sp(const qldpc_code *c, const BPParams &P, const BPGeom &G, cudaStream_t st)
{
    if (P.zero_ok == 2) {           // many all-zero syndromes expected (low error rates): the instantiations with the shortcut
        if (G.warp_var == 1)
            return c->two_tables ? launch_bp_warp_inst3<CPL, VPL, true, 1, true>(c, P, G, st) : launch_bp_warp_inst3<CPL, VPL, false, 1, true>(c, P, G, st);
        return c->two_tables ? launch_bp_warp_inst3<CPL, VPL, true, 2, true>(c, P, G, st) : launch_bp_warp_inst3<CPL, VPL, false, 2, true>(c, P, G, st);
    }
    if (G.warp_var == 1)
        return c->two_tables ? launch_bp_warp_inst3<CPL, VPL, true, 1>(c, P, G, st) : launch_bp_warp_inst3<CPL, VPL, false, 1>(c, P, G, st);
    return c->two_tables ? launch_bp_warp_inst3<CPL, VPL, true, 2>(c, P, G, st) : launch_bp_warp_inst3<CPL, VPL, false, 2>(c, P, G, st);
}

#define QLDPC_WARP_SHAPES(F)                                                  \
    if (c->WM == 2 && c->WN == 3) return F<2, 3>(c, P, G, st);                \
    if (c->WM == 2 && c->WN == 4) return F<2, 4>(c, P, G, st);                \
    if (c->WM == 3 && c->WN == 5) return F<3, 5>(c, P, G, st);                \
    if (c->WM == 5 && c->WN == 9) return F<5, 9>(c, P, G, st);                \
    return cudaErrorInvalidValue

cudaError_t launch_bp_warp_sp(const qldpc_code *c, const BPParams &P, const BPGeom &G, cudaStream_t st)
{
    QLDPC_WARP_SHAPES(launch_sp);
}
