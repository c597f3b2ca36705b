// Launchers of the OSD-0 kernels (osd_kernel.cuh): column-major / row-major warp kernels, block-per-shot kernels.
#include "capi_internal.h"

template <typename K, int WM>
static cudaError_t launch_osd_inst(const qldpc_code *c, const OSDParams &P, long long count_hint, cudaStream_t st)
{
    auto kern = osd0_kernel<K, WM>;
    const size_t smem = osd_smem_colmask(P.n, P.WM) + osd_smem_per_warp<K>(P.n) * OSD_WARPS;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, OSD_WARPS * 32, smem);
    long long grid = (long long)c->num_sms * std::max(1, occ);
    if (count_hint >= 0) grid = std::max<long long>(1, std::min<long long>(grid, (count_hint + OSD_WARPS - 1) / OSD_WARPS));
    kern<<<(int)grid, OSD_WARPS * 32, smem, st>>>(P);
    return cudaGetLastError();
}

template <typename K, int WM, int NS>
static cudaError_t launch_osd_fast_inst(const qldpc_code *c, const OSDParams &P, long long count_hint, cudaStream_t st)
{
    auto kern = osd0_fast_kernel<K, WM, NS>;
    const size_t smem = osd_smem_colmask(P.n, P.WM) + osd_smem_per_warp<double>(P.n) * OSD_WARPS;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, OSD_WARPS * 32, smem);
    long long grid = (long long)c->num_sms * std::max(1, occ);
    if (count_hint >= 0) grid = std::max<long long>(1, std::min<long long>(grid, (count_hint + OSD_WARPS - 1) / OSD_WARPS));
    kern<<<(int)grid, OSD_WARPS * 32, smem, st>>>(P);
    return cudaGetLastError();
}

template <typename K, int WM, int NS2>
static cudaError_t launch_osd_fast2_inst(const qldpc_code *c, const OSDParams &P, long long count_hint, cudaStream_t st)
{
    auto kern = osd0_fast2_kernel<K, WM, NS2>;
    const size_t smem = osd_smem_colmask(P.n, P.WM) + osd_smem_per_warp<double>(P.n) * OSD_WARPS * 2;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, OSD_WARPS * 32, smem);
    long long grid = (long long)c->num_sms * std::max(1, occ);
    if (count_hint >= 0) grid = std::max<long long>(1, std::min<long long>(grid, (count_hint + 2 * OSD_WARPS - 1) / (2 * OSD_WARPS)));
    kern<<<(int)grid, OSD_WARPS * 32, smem, st>>>(P);
    return cudaGetLastError();
}

// column-major kernel: the shapes of the reference's codes (m <= 160, n <= 288); cudaErrorNotSupported otherwise
template <typename K>
static cudaError_t launch_osd_fast(const qldpc_code *c, const OSDParams &P, long long count_hint, cudaStream_t st)
{
    const int NS = (P.n + 31) / 32;
    static const bool one_per_warp = getenv("QLDPC_OSD_ONE_SHOT_PER_WARP") != nullptr;    // test / comparison hook
    if (!one_per_warp) {                       // two shots per warp where the registers allow it
        const int NS2 = (P.n + 15) / 16;
        if (P.WM == 2 && NS2 == 5) return launch_osd_fast2_inst<K, 2, 5>(c, P, count_hint, st);
        if (P.WM == 2 && NS2 == 6) return launch_osd_fast2_inst<K, 2, 6>(c, P, count_hint, st);
        if (P.WM == 2 && NS2 == 7) return launch_osd_fast2_inst<K, 2, 7>(c, P, count_hint, st);
        if (P.WM == 3 && NS2 == 9) return launch_osd_fast2_inst<K, 3, 9>(c, P, count_hint, st);
    }
    if (P.WM == 2 && NS == 3) return launch_osd_fast_inst<K, 2, 3>(c, P, count_hint, st);
    if (P.WM == 2 && NS == 4) return launch_osd_fast_inst<K, 2, 4>(c, P, count_hint, st);
    if (P.WM == 3 && NS == 5) return launch_osd_fast_inst<K, 3, 5>(c, P, count_hint, st);
    if (P.WM == 5 && NS == 9) return launch_osd_fast_inst<K, 5, 9>(c, P, count_hint, st);
    return cudaErrorNotSupported;
}

template <typename K>
static cudaError_t launch_osd_k(const qldpc_code *c, const OSDParams &P, long long count_hint, cudaStream_t st)
{
    static const bool force_rowmajor = getenv("QLDPC_OSD_FORCE_ROWMAJOR") != nullptr;    // test hook
    if (!force_rowmajor) {
        const cudaError_t e = launch_osd_fast<K>(c, P, count_hint, st);
        if (e != cudaErrorNotSupported) return e;
    }
    switch (P.WM) {
    case 1: return launch_osd_inst<K, 1>(c, P, count_hint, st);
    case 2: return launch_osd_inst<K, 2>(c, P, count_hint, st);
    case 3: return launch_osd_inst<K, 3>(c, P, count_hint, st);
    case 4: return launch_osd_inst<K, 4>(c, P, count_hint, st);
    case 5: return launch_osd_inst<K, 5>(c, P, count_hint, st);
    default: return cudaErrorInvalidValue;
    }
}

template <typename K>
static cudaError_t launch_osd_block(const qldpc_code *c, const OSDBlockParams &P, long long count_hint, cudaStream_t st)
{
    auto kern = osd0_block_kernel<K>;
    const size_t smem = osdb_smem_bytes<K>(P.m, P.n);
    if (smem > (size_t)c->smem_optin) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, OSDB_THREADS, smem);
    long long grid = (long long)c->num_sms * std::max(1, occ);
    if (count_hint >= 0) grid = std::max<long long>(1, std::min<long long>(grid, count_hint));
    kern<<<(int)grid, OSDB_THREADS, smem, st>>>(P);
    return cudaGetLastError();
}

template <typename K>
static cudaError_t launch_osd_block_fast(const qldpc_code *c, const OSDBlockParams &P, long long count_hint, cudaStream_t st)
{
    auto kern = P.colpack ? (P.WM == 27 ? osd0_block_fast_kernel<K, true, 27> : osd0_block_fast_kernel<K, true, 0>)
                          : (P.WM == 27 ? osd0_block_fast_kernel<K, false, 27> : osd0_block_fast_kernel<K, false, 0>);   // 27 words: [[144,12,12]] x 12 rounds
    const size_t smem = osdbf_smem_bytes<K>(P.m, P.n);
    if (smem > (size_t)c->smem_optin) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, OSDBF_THREADS, smem);
    long long grid = (long long)c->num_sms * std::max(1, occ);
    if (count_hint >= 0) grid = std::max<long long>(1, std::min<long long>(grid, count_hint));
    kern<<<(int)grid, OSDBF_THREADS, smem, st>>>(P);
    return cudaGetLastError();
}

bool osd_use_block(const qldpc_code *c)
{
    static const bool force = getenv("QLDPC_OSD_FORCE_BLOCK") != nullptr;    // test hook
    return force || c->WM > 5 || c->n > 65535;
}

int osd_launch(qldpc_code *c, OSDParams &P, int llr_f64, long long count_hint, cudaStream_t st, DevBuf *redo)
{
    P.m = c->m; P.n = c->n; P.WM = c->WM; P.WN = c->WN;
    P.rank = c->rank;
    P.colmask = c->d_colmask;
    if (osd_use_block(c)) {
        if (c->n > 65534 || c->m > 32767) return qldpc_fail(QLDPC_ERR_UNSUPPORTED, "OSD: more than 65534 columns or 32767 rows");
        OSDBlockParams Q;
        Q.m = c->m; Q.n = c->n; Q.WM = c->WM; Q.WN = c->WN; Q.rank = c->rank; Q.max_col_w = c->max_col_w;
        Q.var_ptr = c->d_var_ptr; Q.vtab = c->d_vtab1; Q.colpack = c->d_colpack;
        Q.idx = P.idx; Q.count_dev = P.count_dev; Q.count_host = P.count_host;
        Q.synd = P.synd; Q.llr = P.llr; Q.hard = P.hard; Q.out = P.out; Q.valid = P.valid;
        Q.redo_idx = nullptr; Q.redo_count = nullptr;
        static const bool force_rowmajor = getenv("QLDPC_OSD_FORCE_ROWMAJOR") != nullptr;    // test hook
        const long long cap = P.count_dev ? P.cap : P.count_host;
        if (!force_rowmajor && cap > 0 && redo && c->WM <= 32) {
            // column-major kernel; the shots it flags as inconsistent are redone by the row-major one
            CK(redo->reserve(sizeof(int32_t) * (size_t)cap + 16));
            Q.redo_count = redo->as<unsigned int>();
            Q.redo_idx = redo->as<int32_t>() + 4;
            CK(cudaMemsetAsync(Q.redo_count, 0, sizeof(unsigned int), st));
            cudaError_t e = llr_f64 ? launch_osd_block_fast<double>(c, Q, count_hint, st) : launch_osd_block_fast<float>(c, Q, count_hint, st);
            if (e != cudaSuccess)
                return qldpc_fail(e == cudaErrorInvalidValue ? QLDPC_ERR_UNSUPPORTED : QLDPC_ERR_CUDA,
                            std::string("osd0_block_fast_kernel launch (check matrix too large for shared memory?): ") + cudaGetErrorString(e));
            Q.idx = Q.redo_idx; Q.count_dev = Q.redo_count; Q.count_host = 0;
            e = llr_f64 ? launch_osd_block<double>(c, Q, -1, st) : launch_osd_block<float>(c, Q, -1, st);
            if (e != cudaSuccess) return qldpc_fail(QLDPC_ERR_CUDA, std::string("osd0_block_kernel (redo) launch: ") + cudaGetErrorString(e));
            return QLDPC_OK;
        }
        cudaError_t e = llr_f64 ? launch_osd_block<double>(c, Q, count_hint, st) : launch_osd_block<float>(c, Q, count_hint, st);
        if (e != cudaSuccess)
            return qldpc_fail(e == cudaErrorInvalidValue ? QLDPC_ERR_UNSUPPORTED : QLDPC_ERR_CUDA,
                        std::string("osd0_block_kernel launch (check matrix too large for shared memory?): ") + cudaGetErrorString(e));
        return QLDPC_OK;
    }
    cudaError_t e = llr_f64 ? launch_osd_k<double>(c, P, count_hint, st) : launch_osd_k<float>(c, P, count_hint, st);
    if (e != cudaSuccess) return qldpc_fail(QLDPC_ERR_CUDA, std::string("osd0_kernel launch: ") + cudaGetErrorString(e));
    return QLDPC_OK;
}
