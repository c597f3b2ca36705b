// Launcher of the CTA-per-shot kernels with the edge messages staged in global memory (bp_stage_kernel.cuh).
#include "capi_internal.h"
#include "bp_stage_kernel.cuh"

template <typename T, int VAR, int SC, int SV, bool TWO>
static cudaError_t launch_stage_inst(const qldpc_code *c, const BPParams &P, const BPGeom &G, cudaStream_t st)
{
    // float32: offsets in registers, one CTA per SM (measured 9.3-9.9e7 shot-iterations/s against 7.9-8.3e7 with two CTAs per SM
    // and the offsets re-read from the tables); float64: the FP64-latency-bound arithmetic wants the second CTA
    constexpr bool TABREG = sizeof(T) == 4;
    auto kern = bp_stage_kernel<T, VAR, SC, SV, 8, TWO, TABREG>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G.smem);
    if (e != cudaSuccess) return e;
    // float64: 64-bit shared-memory words are served per half-warp -- the labelling built for 16-lane conflict domains
    const BPWarpTables &tab = (sizeof(T) == 8 && c->cta64_ok) ? c->ctab64 : c->ctab;
    kern<<<G.grid, G.threads, G.smem, st>>>(P, tab, c->cta_sv * c->cta_nw);
    return cudaGetLastError();
}

template <typename T, int SC, int SV>
static cudaError_t launch_stage_v(const qldpc_code *c, const BPParams &P, const BPGeom &G, cudaStream_t st)
{
    const int var = G.warp_var;           // 0 min-sum, 1 sum-product, 2 symmetric sum-product
    if (c->two_tables)
        return var == 0 ? launch_stage_inst<T, 0, SC, SV, true>(c, P, G, st)
             : var == 1 ? launch_stage_inst<T, 1, SC, SV, true>(c, P, G, st) : launch_stage_inst<T, 2, SC, SV, true>(c, P, G, st);
    return var == 0 ? launch_stage_inst<T, 0, SC, SV, false>(c, P, G, st)
         : var == 1 ? launch_stage_inst<T, 1, SC, SV, false>(c, P, G, st) : launch_stage_inst<T, 2, SC, SV, false>(c, P, G, st);
}

// occupancy of the kernel that launch_bp_stage would pick (CTAs per SM), for the geometry
int bp_stage_occupancy(const qldpc_code *c, int precision, int threads, size_t smem)
{
    int occ = 1;
    // (all instantiations of one type share the launch bounds; the min-sum one stands for them)
    if (precision == 64) {
        auto k = c->cta_sc == 2 ? bp_stage_kernel<double, 0, 2, 5, 8, false, false> : bp_stage_kernel<double, 0, 3, 7, 8, false, false>;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, threads, smem);
    } else {
        auto k = c->cta_sc == 2 ? bp_stage_kernel<float, 0, 2, 5, 8, false, true> : bp_stage_kernel<float, 0, 3, 7, 8, false, true>;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, threads, smem);
    }
    return occ < 1 ? 1 : occ;
}

cudaError_t launch_bp_stage(const qldpc_code *c, const BPParams &P, const BPGeom &G, int precision, cudaStream_t st)
{
    if (precision == 64)
        return c->cta_sc == 2 ? launch_stage_v<double, 2, 5>(c, P, G, st) : launch_stage_v<double, 3, 7>(c, P, G, st);
    return c->cta_sc == 2 ? launch_stage_v<float, 2, 5>(c, P, G, st) : launch_stage_v<float, 3, 7>(c, P, G, st);
}
