// Launcher of the CTA-per-shot kernels (bp_cta_kernel.cuh): space-time sized check matrices.
#include "capi_internal.h"

cudaError_t launch_bp_cta(const qldpc_code *c, const BPParams &P, const BPGeom &G, cudaStream_t st)
{
    const int var = G.warp_var;
    void (*kern)(const BPParams, const BPWarpTables, int) = nullptr;
#define QLDPC_CTA_PICK(SC, SV)                                                                                              \
    kern = c->two_tables ? (var == 0 ? bp_cta_kernel<SC, SV, 8, true, 0> : var == 1 ? bp_cta_kernel<SC, SV, 8, true, 1>     \
                                                                                    : bp_cta_kernel<SC, SV, 8, true, 2>)    \
                         : (var == 0 ? bp_cta_kernel<SC, SV, 8, false, 0> : var == 1 ? bp_cta_kernel<SC, SV, 8, false, 1>   \
                                                                                     : bp_cta_kernel<SC, SV, 8, false, 2>)
    if (c->cta_sc == 2) { QLDPC_CTA_PICK(2, 5); } else { QLDPC_CTA_PICK(3, 7); }
#undef QLDPC_CTA_PICK
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G.smem);
    if (e != cudaSuccess) return e;
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, G.threads, G.smem);
    const long long grid = std::max<long long>(1, std::min<long long>((long long)c->num_sms * std::max(1, occ), P.B));
    kern<<<(int)grid, G.threads, G.smem, st>>>(P, c->ctab, c->cta_sv * c->cta_nw);
    return cudaGetLastError();
}
