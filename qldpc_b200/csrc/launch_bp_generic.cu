// Launchers of the thread-per-shot kernels (bp_kernel.cuh): on-chip and HBM-staged state.
#include "capi_internal.h"

template <typename T, int VAR, int WMS, bool SMEM>
static cudaError_t launch_bp_inst(const BPParams &P, const BPGeom &G, cudaStream_t st)
{
    auto kern = bp_decode_kernel<T, VAR, WMS, SMEM>;
    if (SMEM) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G.smem);
        if (e != cudaSuccess) return e;
    }
    kern<<<G.grid, G.threads, G.smem, st>>>(P);
    return cudaGetLastError();
}

template <typename T, int VAR>
static cudaError_t launch_bp_tv(const BPParams &P, const BPGeom &G, cudaStream_t st)
{
    if (G.staged) return launch_bp_inst<T, VAR, 0, false>(P, G, st);
    switch (P.g.WM) {
    case 1: return launch_bp_inst<T, VAR, 1, true>(P, G, st);
    case 2: return launch_bp_inst<T, VAR, 2, true>(P, G, st);
    case 3: return launch_bp_inst<T, VAR, 3, true>(P, G, st);
    case 4: case 5: {
        // WM == 4 runs the 5-word instantiation on a 5-word view?  No: keep exact strides.
        if (P.g.WM == 5) return launch_bp_inst<T, VAR, 5, true>(P, G, st);
        return launch_bp_inst<T, VAR, 4, true>(P, G, st);
    }
    default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_bp_generic(const BPParams &P, const BPGeom &G, int precision, int kv, cudaStream_t st)
{
    if (precision == 64)
        return (kv == VAR_MIN_SUM) ? launch_bp_tv<double, VAR_MIN_SUM>(P, G, st) : launch_bp_tv<double, VAR_SUM_PRODUCT>(P, G, st);
    return (kv == VAR_MIN_SUM) ? launch_bp_tv<float, VAR_MIN_SUM>(P, G, st) : launch_bp_tv<float, VAR_SUM_PRODUCT>(P, G, st);
}
