// Internal declarations shared by the translation units of libqldpc_b200.so: the code handle, device workspaces, launch
// geometry and the per-kernel-family launchers (one .cu per family so that nvcc compiles them in parallel; the kernel
// headers are templates, a translation unit only pays for what it instantiates).
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/qldpc_b200.h"
#include "bp_kernel.cuh"
#include "bp_tiled_kernel.cuh"
#include "bp_warp_kernel.cuh"
#include "bp_cta_kernel.cuh"
#include "bp_warp_kernel_f64.cuh"
#include "bp_warp_layout.h"
#include "host_pack.h"
#include "osd_kernel.cuh"

using namespace qldpc;

int qldpc_fail(int code, const std::string &msg);      // records the message behind qldpc_last_error(), returns code
#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return qldpc_fail(QLDPC_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_) + \
                                                  " (" __FILE__ ":" + std::to_string(__LINE__) + ")");  \
    } while (0)

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { want = bytes; e = cudaMalloc(&p, want); }
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

struct HostBuf {                    // pinned host staging buffer
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        const cudaError_t e = cudaHostAlloc(&p, bytes + bytes / 8, cudaHostAllocDefault);
        if (e == cudaSuccess) cap = bytes + bytes / 8;
        return e;
    }
    void release()
    {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

struct Ctrl {                       // device control block, zeroed before every BP launch
    unsigned long long cursor;
    unsigned long long iter_total;
    unsigned int fail_count;
    unsigned int pad;
};

struct qldpc_code {
    int m = 0, n = 0, E = 0, k = 0, WM = 0, WN = 0;
    int uniform_row_w = 0, max_col_w = 0, two_tables = 0;
    int rank = 0;                                       // GF(2) rank of H
    int num_sms = 0, smem_optin = 0;
    int32_t *d_row_ptr = nullptr, *d_col_idx = nullptr, *d_var_ptr = nullptr;
    uint32_t *d_vtab0 = nullptr, *d_vtab1 = nullptr, *d_colmask = nullptr, *d_Lrows = nullptr, *d_Hrows = nullptr;
    uint32_t *d_colpack = nullptr;                      // [n] checks of a column packed 3 x 10 bits + count << 30 (m <= 1024, column weight <= 3; block OSD)
    // T-lanes-per-shot kernel tables: 4 words per position {e0, e1, e2, v}, e = check | k << 16.  [0]: identity positions
    // (float64: the reference's addition order is kept as is), [1]/[2]: positions optimised for TL = 4 / 8 (float32)
    uint32_t *d_vell0[3] = {nullptr, nullptr, nullptr}, *d_vell1[3] = {nullptr, nullptr, nullptr};
    double tiled_conflict_cost[3][2] = {{0, 0}, {0, 0}, {0, 0}};   // modelled wavefronts per shot-iteration: before / after
    uint32_t *d_wtab = nullptr;                                    // warp-per-shot kernel: the six tables of BPWarpTables, back to back
    BPWarpTables wtab = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    WarpLayoutBuilder *wlayout = nullptr;                          // labelling of checks / variables / edge slots (host)
    // float64 warp kernel: the same tables for a labelling with 16-lane conflict domains (64-bit shared-memory words)
    uint32_t *d_wtab64 = nullptr;
    BPWarpTables wtab64 = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    WarpLayoutBuilder *wlayout64 = nullptr;
    int warp64_cost[3] = {0, 0, 0};
    int warp_cost[3] = {0, 0, 0};                                  // gather wavefronts per shot-iteration: natural, current, floor
    bool warp_ok = false;
    // CTA-per-shot kernel (bp_cta_kernel.cuh): labelling with NW * 3 check slots and NW * 7 variable slots
    uint32_t *d_ctab = nullptr;
    BPWarpTables ctab = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int cta_nw = 0, cta_sc = 3, cta_sv = 7, cta_cost[3] = {0, 0, 0};
    bool cta_ok = false;
    // the same shape labelled for 16-lane conflict domains (64-bit shared-memory words: float64 bp_stage_kernel)
    uint32_t *d_ctab64 = nullptr;
    BPWarpTables ctab64 = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int cta64_cost[3] = {0, 0, 0};
    bool cta64_ok = false;
    int max_row_w = 0;
    double prior_max = 0.0;
    bool prior_uniform = false;
    bool prior_positive = false;                                   // every prior > 0, also after rounding to float32
    double zero_frac = 0.0;                                        // prod_i (1 - p_i), p_i = 1 / (1 + e^prior_i): the share of error-free shots the priors imply
    bool tiled_ok = false;
    std::vector<double> prior_cache;
    DevBuf prior32, prior64, ctrl, gstate;
    DevBuf ws_redo;
    DevBuf ws_synd, ws_hard, ws_err, ws_conv, ws_iters, ws_llr, ws_fail, ws_valid, ws_u8a, ws_u8b, ws_flags,
        ws_weight, ws_cnt, ws_llr_in, ws_rec, ws_inv;
    // Three-stage pipeline of the host-pointer decode call: a copy-in stream, a compute stream and a copy-out stream,
    // chained per chunk by events; chunk buffers rotate over NSLOT slots.  Kernels of different chunks never share the
    // GPU (each runs at full speed), the copies of the neighbouring chunks run under them.
    struct Slot {
        DevBuf ctrl, gstate, u8in, u8out, synd, hard, conv, iters, llr, fail, redo, valid, inv;
        HostBuf h_synd, h_hard;       // host-side packing: pinned bit-packed rows of the chunk (host_pack.h)
        cudaEvent_t ev_in = nullptr, ev_comp = nullptr, ev_out0 = nullptr, ev_out = nullptr;   // (ev_out0 .. ev_out: the copy-out, timed)
        bool used = false;
        long long pend_o = 0, pend_b = 0;   // host-side packing: the chunk whose corrections still sit in h_hard
        bool pend = false;
        bool host_mode = false;             // the slot's current chunk is packed by host threads
        bool inflight = false;              // its copy-out has not been seen complete yet
        long long cur_b = 0;
    };
    static constexpr int NSLOT = 4;
    Slot slot[NSLOT];
    cudaStream_t st_in = nullptr, st_comp = nullptr, st_out = nullptr;
    // uint8 rows of the host-pointer decode call: packed on the device (0) or by host threads (1); -1: not decided yet
    // or chunk by chunk by whichever side is free (2)
    HostPool *pool = nullptr;
    int host_pack = -1;
    double host_pack_rate = 0.0;            // measured pack + unpack throughput of the pool, shots / s
    double est_dev = 0.0, est_pack = 0.0, est_unpack = 0.0;   // running estimates, seconds per shot: copy-out of byte rows, host pack, host expansion
    unsigned long long host_chunks = 0, dev_chunks = 0;
    unsigned long long h2d_bytes = 0, d2h_bytes = 0;    // bytes moved by the host-pointer decode calls (cumulative)
    BPGraphDev graph() const
    {
        BPGraphDev g;
        g.m = m; g.n = n; g.E = E; g.WM = WM; g.WN = WN;
        g.uniform_row_w = uniform_row_w; g.max_col_w = max_col_w; g.two_tables = two_tables;
        g.row_ptr = d_row_ptr; g.col_idx = d_col_idx; g.var_ptr = d_var_ptr;
        g.vtab0 = d_vtab0; g.vtab1 = d_vtab1; g.colmask = d_colmask;
        return g;
    }
};

struct BPGeom {
    bool staged;
    int tiled_T;          // 0: thread-per-shot kernels; 4 / 8: lanes per shot of the tiled kernel
    bool warp_kernel;     // warp-per-shot kernel (messages in registers)
    int warp_var;         // its variant: 0 min-sum, 1 sum-product, 2 symmetric sum-product
    bool cta_kernel;      // CTA-per-shot kernel (messages in registers, several warps per shot)
    bool stage_kernel;    // CTA-per-shot kernel with the messages staged in global memory (bp_stage_kernel.cuh)
    int shots_per_cta;
    int refill_min;
    int threads, grid;
    size_t smem;
    size_t gstate_bytes;
};

// launchers (launch_*.cu)
cudaError_t launch_bp_generic(const BPParams &P, const BPGeom &G, int precision, int kv, cudaStream_t st);
cudaError_t launch_bp_tiled(const qldpc_code *c, const BPParams &P, const BPGeom &G, int precision, int kv, cudaStream_t st);
cudaError_t launch_bp_warp(const qldpc_code *c, const BPParams &P, const BPGeom &G, cudaStream_t st);
cudaError_t launch_bp_warp_sp(const qldpc_code *c, const BPParams &P, const BPGeom &G, cudaStream_t st);
cudaError_t launch_bp_warp_f64(const qldpc_code *c, const BPParams &P, const BPGeom &G, cudaStream_t st);
cudaError_t launch_bp_warp_f64_sp(const qldpc_code *c, const BPParams &P, const BPGeom &G, cudaStream_t st);
cudaError_t launch_bp_cta(const qldpc_code *c, const BPParams &P, const BPGeom &G, cudaStream_t st);
cudaError_t launch_bp_stage(const qldpc_code *c, const BPParams &P, const BPGeom &G, int precision, cudaStream_t st);
int bp_stage_occupancy(const qldpc_code *c, int precision, int threads, size_t smem);
bool osd_use_block(const qldpc_code *c);
int osd_launch(qldpc_code *c, OSDParams &P, int llr_f64, long long count_hint, cudaStream_t st, DevBuf *redo = nullptr);
