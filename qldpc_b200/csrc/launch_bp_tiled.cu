// Launchers of the T-lanes-per-shot kernels (bp_tiled_kernel.cuh).
#include "capi_internal.h"

template <typename T, int VAR, int TL, int WMS>
static cudaError_t launch_bp_tiled_inst(const qldpc_code *c, const BPParams &P, const BPGeom &G, cudaStream_t st)
{
    auto kern = bp_tiled_kernel<T, VAR, TL, WMS, 6>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G.smem);
    if (e != cudaSuccess) return e;
    const int ti = (sizeof(T) == 8) ? 0 : (TL == 4 ? 1 : 2);     // float64 keeps identity positions
    kern<<<G.grid, G.threads, G.smem, st>>>(P, c->d_vell0[ti], c->d_vell1[ti], G.refill_min);
    return cudaGetLastError();
}

template <typename T, int VAR, int TL>
static cudaError_t launch_bp_tiled_w(const qldpc_code *c, const BPParams &P, const BPGeom &G, cudaStream_t st)
{
    switch (P.g.WM) {
    case 2: return launch_bp_tiled_inst<T, VAR, TL, 2>(c, P, G, st);
    case 3: return launch_bp_tiled_inst<T, VAR, TL, 3>(c, P, G, st);
    case 5: return launch_bp_tiled_inst<T, VAR, TL, 5>(c, P, G, st);
    default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_bp_tiled(const qldpc_code *c, const BPParams &P, const BPGeom &G, int precision, int kv, cudaStream_t st)
{
    if (precision == 32 && kv == VAR_MIN_SUM)
        return G.tiled_T == 4 ? launch_bp_tiled_w<float, VAR_MIN_SUM, 4>(c, P, G, st) : launch_bp_tiled_w<float, VAR_MIN_SUM, 8>(c, P, G, st);
    if (G.tiled_T != 8) return cudaErrorInvalidValue;
    if (precision == 32) return launch_bp_tiled_w<float, VAR_SUM_PRODUCT, 8>(c, P, G, st);
    if (kv == VAR_MIN_SUM) return launch_bp_tiled_w<double, VAR_MIN_SUM, 8>(c, P, G, st);
    return launch_bp_tiled_w<double, VAR_SUM_PRODUCT, 8>(c, P, G, st);
}
