// C ABI of libqldpc_b200.so (declared in include/qldpc_b200.h): code handle, launch geometry,
// workspaces and the host/device entry points around the kernels in bp_kernel.cuh,
// osd_kernel.cuh, osdw_kernel.cuh and misc_kernels.cuh.  No torch types, no CPU fallback.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "capi_internal.h"
#include <chrono>
#include "misc_kernels.cuh"     // non-template kernels: defined in this translation unit only
#include "osdw_kernel.cuh"
#include "bp_stage_kernel.cuh"   // (layout helpers only; the kernels are instantiated in launch_bp_stage.cu)

static thread_local std::string g_err;
int qldpc_fail(int code, const std::string &msg)
{
    g_err = msg;
    return code;
}
static int fail(int code, const std::string &msg) { return qldpc_fail(code, msg); }


extern "C" const char *qldpc_last_error(void) { return g_err.c_str(); }
extern "C" int qldpc_version(void) { return 100; }
extern "C" int qldpc_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
extern "C" int qldpc_words_m(const qldpc_code *c) { return c ? c->WM : 0; }
extern "C" int qldpc_words_n(const qldpc_code *c) { return c ? c->WN : 0; }

template <typename T> static cudaError_t upload(T **dst, const std::vector<T> &src)
{
    cudaError_t e = cudaMalloc((void **)dst, std::max<size_t>(1, src.size()) * sizeof(T));
    if (e != cudaSuccess) return e;
    if (src.empty()) return cudaSuccess;
    return cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice);
}

static int code_create_impl(int32_t m, int32_t n, const int32_t *row_ptr, const int32_t *col_idx, const int32_t *var_ptr,
                            const int32_t *var_edge0, const int32_t *var_edge1, int32_t k, const uint8_t *L, qldpc_code *c);

extern "C" int qldpc_code_create(int32_t m, int32_t n, const int32_t *row_ptr, const int32_t *col_idx,
                                 const int32_t *var_ptr, const int32_t *var_edge0, const int32_t *var_edge1,
                                 int32_t k, const uint8_t *L, qldpc_code **out)
{
    if (!out || m <= 0 || n <= 0 || !row_ptr || !col_idx || !var_ptr || !var_edge0 || !var_edge1 || (k > 0 && !L))
        return fail(QLDPC_ERR_ARG, "qldpc_code_create: bad argument");
    if (qldpc_device_count() <= 0) return fail(QLDPC_ERR_CUDA, "qldpc_code_create: no CUDA device (there is no CPU fallback)");
    const int E = row_ptr[m];
    if (E <= 0 || var_ptr[n] != E) return fail(QLDPC_ERR_ARG, "qldpc_code_create: inconsistent CSR / variable pointers");
    qldpc_code *c = new qldpc_code();
    const int rc = code_create_impl(m, n, row_ptr, col_idx, var_ptr, var_edge0, var_edge1, k, L, c);
    if (rc != QLDPC_OK) {               // release whatever was uploaded before the failure
        const std::string msg = g_err;
        qldpc_code_destroy(c);
        g_err = msg;
        return rc;
    }
    *out = c;
    return QLDPC_OK;
}

// uploads the six tables of a labelling back to back; the device must be idle when *dbuf is already in use
static int layout_tables_upload(const WarpLayout &L, uint32_t **dbuf, BPWarpTables *tab, int cost[3])
{
    std::vector<uint32_t> all;
    size_t off[6];
    const std::vector<uint32_t> *parts[6] = {&L.sidx, &L.sidx0, &L.vidx, &L.cinfo, &L.vorig, &L.vpos};
    for (int i = 0; i < 6; ++i) { off[i] = all.size(); all.insert(all.end(), parts[i]->begin(), parts[i]->end()); }
    if (!*dbuf) CK(cudaMalloc(dbuf, all.size() * sizeof(uint32_t)));
    CK(cudaMemcpy(*dbuf, all.data(), all.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    *tab = BPWarpTables{*dbuf + off[0], *dbuf + off[1], *dbuf + off[2], *dbuf + off[3], *dbuf + off[4], *dbuf + off[5]};
    cost[0] = L.cost_natural; cost[1] = L.cost; cost[2] = L.floor;
    return QLDPC_OK;
}

// (re)builds the device tables of the warp-per-shot kernel from c->wlayout; the device must be idle
static int warp_tables_upload(qldpc_code *c)
{
    if (int rc = layout_tables_upload(c->wlayout->tables(), &c->d_wtab, &c->wtab, c->warp_cost)) return rc;
    return layout_tables_upload(c->wlayout64->tables(), &c->d_wtab64, &c->wtab64, c->warp64_cost);
}

static int code_create_impl(int32_t m, int32_t n, const int32_t *row_ptr, const int32_t *col_idx, const int32_t *var_ptr,
                            const int32_t *var_edge0, const int32_t *var_edge1, int32_t k, const uint8_t *L, qldpc_code *c)
{
    const int E = row_ptr[m];
    c->m = m; c->n = n; c->E = E; c->k = k;
    c->WM = (m + 31) / 32;
    c->WN = (n + 31) / 32;
    int dev = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&c->num_sms, cudaDevAttrMultiProcessorCount, dev));
    CK(cudaDeviceGetAttribute(&c->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));

    // edge -> check
    std::vector<int32_t> edge_check(E);
    int rw = row_ptr[1] - row_ptr[0];
    for (int r = 0; r < m; ++r) {
        if (row_ptr[r + 1] - row_ptr[r] != rw) rw = 0;
        for (int e = row_ptr[r]; e < row_ptr[r + 1]; ++e) {
            if (col_idx[e] < 0 || col_idx[e] >= n) return fail(QLDPC_ERR_ARG, "qldpc_code_create: column index out of range");
            edge_check[e] = r;
        }
    }
    c->uniform_row_w = (rw > 0 && rw <= 8) ? rw : 0;
    for (int r = 0; r < m; ++r) c->max_row_w = std::max(c->max_row_w, row_ptr[r + 1] - row_ptr[r]);
    std::vector<uint32_t> vt0(2 * (size_t)E), vt1(2 * (size_t)E), colmask((size_t)n * c->WM, 0u);
    for (int v = 0; v < n; ++v) {
        c->max_col_w = std::max(c->max_col_w, var_ptr[v + 1] - var_ptr[v]);
        for (int a = var_ptr[v]; a < var_ptr[v + 1]; ++a) {
            const int e0 = var_edge0[a], e1 = var_edge1[a];
            if (e0 < 0 || e0 >= E || e1 < 0 || e1 >= E || col_idx[e0] != v || col_idx[e1] != v)
                return fail(QLDPC_ERR_ARG, "qldpc_code_create: var_edge table does not match the CSR");
            vt0[2 * a] = e0; vt0[2 * a + 1] = edge_check[e0];
            vt1[2 * a] = e1; vt1[2 * a + 1] = edge_check[e1];
            if (e0 != e1) c->two_tables = 1;
            colmask[(size_t)v * c->WM + (edge_check[e1] >> 5)] |= 1u << (edge_check[e1] & 31);
        }
    }
    // ELL view for the T-lanes-per-shot kernel: uniform row weight 6 (all BB codes), column weight <= 3
    int min_col_w = n > 0 ? 1 << 30 : 0;
    for (int v = 0; v < n; ++v) min_col_w = std::min(min_col_w, var_ptr[v + 1] - var_ptr[v]);
    c->tiled_ok = (c->uniform_row_w == 6) && (c->max_col_w == 3) && (min_col_w == 3) && (m < 65536) && (c->WM == 2 || c->WM == 3 || c->WM == 5);
    std::vector<uint32_t> vell0[3], vell1[3];
    if (c->tiled_ok) {
        // entry of the t-th added edge of variable v in table `ve`
        auto entry = [&](const int32_t *ve, int v, int t) {
            const int e = ve[var_ptr[v] + t];
            return (uint32_t)edge_check[e] | ((uint32_t)(e - row_ptr[edge_check[e]]) << 16);
        };
        // shared-memory wavefronts of one warp-wide access by TL lane groups to rows r[j] with `es`-byte elements
        // (S = G mod 32 slots per row, G = 32/TL consecutive slots per group; 64-bit accesses go by half-warps)
        auto wavefronts = [&](const int *r, int TL, int es) {
            const int Gs = 32 / TL, wpe = es / 4;
            int total = 0;
            const int halves = (es == 8) ? 2 : 1;
            for (int h = 0; h < halves; ++h) {
                int cnt[32] = {0};
                int mx = 0;
                const int j0 = h * TL / halves, j1 = (h + 1) * TL / halves;
                for (int jj = j0; jj < j1; ++jj)
                    for (int gg = 0; gg < Gs; ++gg)
                        for (int wv = 0; wv < wpe; ++wv) {
                            const int bank = (((r[jj] * Gs + gg) * wpe) + wv) & 31;
                            mx = std::max(mx, ++cnt[bank]);
                        }
                total += mx;
            }
            return total;
        };
        for (int ti = 0; ti < 3; ++ti) {
            const int TL = ti == 0 ? 8 : (ti == 1 ? 4 : 8);
            std::vector<int> pos2var(n);
            std::vector<uint8_t> sw0(n, 0), sw1(n, 0);
            for (int v = 0; v < n; ++v) pos2var[v] = v;
            // cost of the steps of one 32-variable word: message rows (load + store, 4-byte) and summary rows (8-byte)
            auto step_cost = [&](const int32_t *ve, const std::vector<uint8_t> &sw, int p0, int cntp) {
                double cost = 0;
                for (int t = 0; t < 3; ++t) {
                    int rq[8], rm[8];
                    for (int jj = 0; jj < TL; ++jj) {
                        if (jj < cntp) {
                            const int v = pos2var[p0 + jj];
                            const int tt = (t < 2 && sw[v]) ? 1 - t : t;
                            const uint32_t e = entry(ve, v, tt);
                            rq[jj] = (int)((e >> 16) * m + (e & 0xffffu));
                            rm[jj] = (int)(e & 0xffffu);
                        } else { rq[jj] = -1000 - jj * 977; rm[jj] = -1000 - jj * 977; }
                    }
                    cost += 2.0 * wavefronts(rq, TL, 4) + wavefronts(rm, TL, 8);
                }
                return cost;
            };
            auto total_cost = [&]() {
                double cst = 0;
                for (int p0 = 0; p0 < n; p0 += TL) cst += step_cost(var_edge1, sw1, p0, std::min(TL, n - p0));
                return cst;
            };
            c->tiled_conflict_cost[ti][0] = total_cost();
            if (ti > 0) {
                // local search inside each 32-variable word: swap two positions / toggle the order of the two first-added
                // edges of a variable ((x + y) + z == (y + x) + z, so float results do not change)
                for (int w0 = 0; w0 < n; w0 += 32) {
                    const int w1 = std::min(n, w0 + 32);
                    auto word_cost = [&]() {
                        double cst = 0;
                        for (int p0 = w0; p0 < w1; p0 += TL) cst += step_cost(var_edge1, sw1, p0, std::min(TL, w1 - p0));
                        return cst;
                    };
                    double best = word_cost();
                    for (int pass = 0; pass < 8; ++pass) {
                        bool improved = false;
                        for (int a = w0; a < w1; ++a) {
                            const int va = pos2var[a];
                            sw1[va] ^= 1;
                            double cst = word_cost();
                            if (cst < best) { best = cst; improved = true; } else sw1[va] ^= 1;
                            for (int b = a + 1; b < w1; ++b) {
                                if (a / TL == b / TL) continue;
                                std::swap(pos2var[a], pos2var[b]);
                                cst = word_cost();
                                if (cst < best) { best = cst; improved = true; } else std::swap(pos2var[a], pos2var[b]);
                            }
                        }
                        if (!improved) break;
                    }
                }
                sw0 = sw1;      // iteration-0 table: same positions, same swaps (only used with non-uniform priors)
            }
            c->tiled_conflict_cost[ti][1] = total_cost();
            vell0[ti].assign((size_t)n * 4, 0u);
            vell1[ti].assign((size_t)n * 4, 0u);
            for (int p = 0; p < n; ++p) {
                const int v = pos2var[p];
                for (int t = 0; t < 3; ++t) {
                    vell0[ti][(size_t)p * 4 + t] = entry(var_edge0, v, (t < 2 && sw0[v]) ? 1 - t : t);
                    vell1[ti][(size_t)p * 4 + t] = entry(var_edge1, v, (t < 2 && sw1[v]) ? 1 - t : t);
                }
                vell0[ti][(size_t)p * 4 + 3] = vell1[ti][(size_t)p * 4 + 3] = (uint32_t)v;
            }
        }
    }
    // warp-per-shot kernel (bp_warp_kernel.cuh): labelling and gather tables (bp_warp_layout.h)
    c->warp_ok = c->tiled_ok && ((c->WM == 2 && (c->WN == 3 || c->WN == 4)) || (c->WM == 3 && c->WN == 5) || (c->WM == 5 && c->WN == 9));
    if (c->warp_ok) {
        c->wlayout = new WarpLayoutBuilder(m, n, row_ptr, col_idx, var_ptr, var_edge0, var_edge1, edge_check.data(), 6);
        c->wlayout64 = new WarpLayoutBuilder(m, n, row_ptr, col_idx, var_ptr, var_edge0, var_edge1, edge_check.data(), 6, 0, 0, 16);
        if (!getenv("QLDPC_WARP_NATURAL_LAYOUT")) { c->wlayout->construct(); c->wlayout64->construct(); }
        if (int rc = warp_tables_upload(c)) return rc;
    }
    // CTA-per-shot kernel: larger matrices with row weight <= 8 and column weight <= 3 (the space-time matrices): the
    // smallest number of warps NW <= 12 whose 3 NW check slots / 7 NW variable slots admit a conflict-free labelling
    if (!c->warp_ok && m > 160 && c->max_row_w <= 8 && c->max_col_w <= 3 && !getenv("QLDPC_NO_CTA_KERNEL")) {
        // shapes (check slots, variable slots per warp): (3, 7) with up to 12 warps, else (2, 5) with up to 18 (measured on
        // the 864 x 2592 matrix: 87.5 against 84.0 M shot-iterations/s at p = 0.003 -- lighter warps do not pay)
        const int shapes[2][3] = {{3, 7, 12}, {2, 5, 18}};
        const int first_shape = getenv("QLDPC_CTA_SHAPE_2_5") ? 1 : 0;
        for (int sh = first_shape; sh < 2 && !c->cta_ok; ++sh) {
            const int sc = shapes[sh][0], sv = shapes[sh][1], nwmax = shapes[sh][2];
            int tries = 0;                                  // (a failed search costs ~0.1 s: at most three per shape)
            for (int nw = std::max(2, (m + 32 * sc - 1) / (32 * sc)); nw <= nwmax && !c->cta_ok && tries < 3; ++nw) {
                if (sv * nw * 32 < n) continue;
                ++tries;
                WarpLayoutBuilder lb(m, n, row_ptr, col_idx, var_ptr, var_edge0, var_edge1, edge_check.data(), 8, sc * nw, sv * nw);
                if (!lb.construct(200000)) continue;
                if (int rc = layout_tables_upload(lb.tables(), &c->d_ctab, &c->ctab, c->cta_cost)) return rc;
                c->cta_nw = nw; c->cta_sc = sc; c->cta_sv = sv;
                c->cta_ok = bp_cta_smem(sv * nw) <= (size_t)c->smem_optin;
                if (c->cta_ok) {
                    WarpLayoutBuilder lb64(m, n, row_ptr, col_idx, var_ptr, var_edge0, var_edge1, edge_check.data(), 8, sc * nw, sv * nw, 16);
                    if (lb64.construct(400000)) {
                        if (int rc = layout_tables_upload(lb64.tables(), &c->d_ctab64, &c->ctab64, c->cta64_cost)) return rc;
                        c->cta64_ok = true;
                    }
                }
            }
        }
    }
    std::vector<uint32_t> Lrows((size_t)std::max(k, 0) * c->WN, 0u), Hrows((size_t)m * c->WN, 0u);
    for (int r = 0; r < k; ++r)
        for (int j = 0; j < n; ++j)
            if (L[(size_t)r * n + j] & 1) Lrows[(size_t)r * c->WN + (j >> 5)] |= 1u << (j & 31);
    for (int r = 0; r < m; ++r)
        for (int e = row_ptr[r]; e < row_ptr[r + 1]; ++e) Hrows[(size_t)r * c->WN + (col_idx[e] >> 5)] |= 1u << (col_idx[e] & 31);
    {   // GF(2) rank of H (bounds the pivot search of OSD)
        std::vector<uint32_t> A(Hrows);
        int r = 0;
        for (int col = 0; col < n && r < m; ++col) {
            int piv = -1;
            for (int i = r; i < m; ++i) if ((A[(size_t)i * c->WN + (col >> 5)] >> (col & 31)) & 1u) { piv = i; break; }
            if (piv < 0) continue;
            for (int w = 0; w < c->WN; ++w) std::swap(A[(size_t)r * c->WN + w], A[(size_t)piv * c->WN + w]);
            for (int i = 0; i < m; ++i)
                if (i != r && ((A[(size_t)i * c->WN + (col >> 5)] >> (col & 31)) & 1u))
                    for (int w = 0; w < c->WN; ++w) A[(size_t)i * c->WN + w] ^= A[(size_t)r * c->WN + w];
            ++r;
        }
        c->rank = r;
    }
    std::vector<int32_t> rp(row_ptr, row_ptr + m + 1), ci(col_idx, col_idx + E), vp(var_ptr, var_ptr + n + 1);
    CK(upload(&c->d_row_ptr, rp));
    CK(upload(&c->d_col_idx, ci));
    CK(upload(&c->d_var_ptr, vp));
    CK(upload(&c->d_vtab0, vt0));
    CK(upload(&c->d_vtab1, vt1));
    CK(upload(&c->d_colmask, colmask));
    if (m <= 1024 && c->max_col_w <= 3) {
        std::vector<uint32_t> colpack((size_t)n, 0u);
        for (int v = 0; v < n; ++v) {
            const int cnt = var_ptr[v + 1] - var_ptr[v];
            uint32_t e = (uint32_t)cnt << 30;
            for (int t = 0; t < cnt; ++t) e |= vt1[2 * (size_t)(var_ptr[v] + t) + 1] << (10 * t);
            colpack[v] = e;
        }
        CK(upload(&c->d_colpack, colpack));
    }
    CK(upload(&c->d_Lrows, Lrows));
    CK(upload(&c->d_Hrows, Hrows));
    for (int ti = 0; ti < 3; ++ti) {
        CK(upload(&c->d_vell0[ti], vell0[ti]));
        CK(upload(&c->d_vell1[ti], vell1[ti]));
    }
    CK(c->ctrl.reserve(sizeof(Ctrl)));
    return QLDPC_OK;
}

extern "C" void qldpc_code_destroy(qldpc_code *c)
{
    if (!c) return;
    cudaFree(c->d_row_ptr); cudaFree(c->d_col_idx); cudaFree(c->d_var_ptr);
    cudaFree(c->d_vtab0); cudaFree(c->d_vtab1); cudaFree(c->d_colmask); cudaFree(c->d_colpack); cudaFree(c->d_Lrows); cudaFree(c->d_Hrows);
    for (int ti = 0; ti < 3; ++ti) { cudaFree(c->d_vell0[ti]); cudaFree(c->d_vell1[ti]); }
    cudaFree(c->d_wtab);
    cudaFree(c->d_wtab64);
    delete c->wlayout64;
    cudaFree(c->d_ctab);
    cudaFree(c->d_ctab64);
    delete c->wlayout;
    DevBuf *bufs[] = {&c->prior32, &c->prior64, &c->ctrl, &c->gstate, &c->ws_synd, &c->ws_hard, &c->ws_err, &c->ws_conv,
                      &c->ws_iters, &c->ws_llr, &c->ws_fail, &c->ws_valid, &c->ws_u8a, &c->ws_u8b, &c->ws_flags, &c->ws_redo,
                      &c->ws_weight, &c->ws_cnt, &c->ws_llr_in, &c->ws_rec, &c->ws_inv};
    for (DevBuf *b : bufs) b->release();
    for (auto &sl : c->slot) {
        DevBuf *sb[] = {&sl.ctrl, &sl.gstate, &sl.u8in, &sl.u8out, &sl.synd, &sl.hard, &sl.conv, &sl.iters, &sl.llr, &sl.fail, &sl.redo, &sl.valid, &sl.inv};
        for (DevBuf *b : sb) b->release();
        sl.h_synd.release();
        sl.h_hard.release();
        for (cudaEvent_t e : {sl.ev_in, sl.ev_comp, sl.ev_out0, sl.ev_out})
            if (e) cudaEventDestroy(e);
    }
    for (cudaStream_t st : {c->st_in, c->st_comp, c->st_out})
        if (st) cudaStreamDestroy(st);
    delete c->pool;
    delete c;
}

// ------------------------------------------------------------------------------------------------
// BP launch geometry
// ------------------------------------------------------------------------------------------------

static int kernel_variant(int v) { return v == QLDPC_MIN_SUM ? VAR_MIN_SUM : VAR_SUM_PRODUCT; }

static int bp_geometry(const qldpc_code *c, const qldpc_bp_config *cfg, long long B, BPGeom *G)
{
    const int tsize = cfg->precision == 64 ? 8 : 4;
    const BPGraphDev g = c->graph();
    const BPSmemLayout L = bp_smem_layout(g, tsize, kernel_variant(cfg->variant));
    const bool wm_ok = (c->WM <= 5);
    G->tiled_T = 0;
    G->refill_min = 1;
    G->warp_kernel = false;
    G->cta_kernel = false;
    G->stage_kernel = false;
    // CTA-per-shot, messages staged in global memory: on request (staged = 5), and for everything the register-resident CTA
    // kernel does not serve on these matrices -- float64, and through it the reference's exact (tanh-domain) sum-product
    if (c->cta_ok && (cfg->staged == 5 || (cfg->staged == 0 && cfg->precision == 64))) {
        const int tsize = cfg->precision == 64 ? 8 : 4;
        G->staged = false;
        G->stage_kernel = true;
        G->warp_var = cfg->variant == QLDPC_MIN_SUM ? 0 : (cfg->variant == QLDPC_SUM_PRODUCT ? 1 : 2);
        G->threads = c->cta_nw * 32;
        G->shots_per_cta = 1;
        G->smem = bp_stage_smem(c->cta_sv * c->cta_nw, tsize);
        if (G->smem <= (size_t)c->smem_optin) {
            const int occ = bp_stage_occupancy(c, cfg->precision, G->threads, G->smem);
            G->grid = (int)std::max<long long>(1, std::min<long long>((long long)c->num_sms * occ, B));
            G->gstate_bytes = bp_stage_gstate(c->cta_nw, c->cta_sc, 8, tsize) * (size_t)G->grid;
            return QLDPC_OK;
        }
        G->stage_kernel = false;
    }
    // (the float32 warp / CTA kernels clip with one xorsign-min against +clip: valid for clip >= 0)
    const bool clip_ok = cfg->clip >= 0.0 || cfg->variant == QLDPC_SUM_PRODUCT;
    if ((cfg->staged == 0 || cfg->staged == 4) && c->cta_ok && cfg->precision == 32 && clip_ok) {
        G->staged = false;
        G->cta_kernel = true;
        G->warp_var = cfg->variant == QLDPC_MIN_SUM ? 0 : (cfg->variant == QLDPC_SUM_PRODUCT ? 1 : 2);
        G->threads = c->cta_nw * 32;
        G->shots_per_cta = 1;
        G->smem = bp_cta_smem(c->cta_sv * c->cta_nw);
        G->grid = 0;
        G->gstate_bytes = 0;
        return QLDPC_OK;
    }
    // (float64 warp kernel: no -0.0 canonicalisation, valid for damping > 0, clip >= 0 and max_iter <= 500 -- see its header)
    // plain sum-product (no damping, no clip) always qualifies
    const bool f64_warp_ok = cfg->variant == QLDPC_SUM_PRODUCT ||
                             (cfg->damping > 0.0 && cfg->clip >= 0.0 && (cfg->variant == QLDPC_SUM_PRODUCT_SYM || cfg->max_iter <= 500));
    if ((cfg->staged == 0 || cfg->staged == 3) && c->warp_ok && ((cfg->precision == 32 && clip_ok) || (cfg->precision == 64 && f64_warp_ok)) &&
        (cfg->staged == 3 || cfg->lanes_per_shot == 0 || cfg->lanes_per_shot == 32)) {
        G->staged = false;
        G->warp_kernel = true;
        // 0 min-sum, 1 sum-product, 2 symmetric sum-product (float32); 3 / 4 / 5 the same in float64 (3: the bit-exact parity mode)
        G->warp_var = (cfg->precision == 64 ? 3 : 0) + (cfg->variant == QLDPC_MIN_SUM ? 0 : (cfg->variant == QLDPC_SUM_PRODUCT ? 1 : 2));
        const int nw = cfg->precision == 64 ? BPW64_WARPS : BPW_WARPS;
        G->threads = nw * 32;
        G->shots_per_cta = nw;
        G->smem = cfg->precision == 64 ? bp_warp64_smem_per_warp(c->WN) * nw : bp_warp_smem_per_warp(c->WN) * nw;
        G->grid = 0;                 // filled at launch from the occupancy query
        G->gstate_bytes = 0;
        return QLDPC_OK;
    }
    if (cfg->staged == 0 && c->tiled_ok) {
        const bool f32ms = (cfg->precision == 32 && cfg->variant == QLDPC_MIN_SUM);   // both lane counts are built for it
        // T lanes per shot; NW warps with NW = 1 (mod T) keeps the check pass bank-conflict free
        int bestT = 0, bestNW = 0;
        for (int T : {4, 8}) {
            if (!f32ms && T != 8) continue;
            if (f32ms && cfg->lanes_per_shot && cfg->lanes_per_shot != T) continue;
            const BPTiledLayout TL = bp_tiled_layout(g, T, tsize, kernel_variant(cfg->variant));
            const int Gs = 32 / T;
            if (TL.tables + (size_t)Gs * TL.per_slot > (size_t)c->smem_optin) continue;
            long long smax = (long long)(((size_t)c->smem_optin - TL.tables) / TL.per_slot);
            long long nw = std::min<long long>(smax / Gs, 18);   // __launch_bounds__(576)
            nw = std::min<long long>(nw, std::max<long long>(1, (B + Gs - 1) / Gs));
            while (nw > 1 && (nw % T) != 1) --nw;
            if (nw >= 1 && (nw > bestNW || bestT == 0)) { bestT = T; bestNW = (int)nw; }
        }
        if (bestT) {
            const BPTiledLayout TL = bp_tiled_layout(g, bestT, tsize, kernel_variant(cfg->variant));
            const int Gs = 32 / bestT;
            G->staged = false;
            G->tiled_T = bestT;
            G->threads = 32 * bestNW;
            G->shots_per_cta = bestNW * Gs;
            G->smem = TL.tables + (size_t)G->shots_per_cta * TL.per_slot;
            G->grid = (int)std::max<long long>(1, std::min<long long>((B + G->shots_per_cta - 1) / G->shots_per_cta, c->num_sms));
            G->gstate_bytes = 0;
            G->refill_min = cfg->refill_min > 0 ? std::min(cfg->refill_min, Gs) : std::max(1, Gs / 4);
            return QLDPC_OK;
        }
    }
    long long nt = 0;
    if (cfg->staged != 1 && wm_ok && L.tables + 32 * L.per_shot <= (size_t)c->smem_optin)
        nt = std::min<long long>(256, (long long)(((size_t)c->smem_optin - L.tables) / L.per_shot)) / 32 * 32;
    if (nt >= 32) {
        G->staged = false;
        long long want = std::max<long long>(32, (B + 31) / 32 * 32);
        G->threads = (int)std::min<long long>(nt, want);
        G->shots_per_cta = G->threads;
        G->smem = L.tables + (size_t)G->threads * L.per_shot;
        G->grid = (int)std::max<long long>(1, std::min<long long>((B + G->threads - 1) / G->threads, c->num_sms));
        G->gstate_bytes = 0;
    } else {
        G->staged = true;
        G->threads = 128;
        G->shots_per_cta = 128;
        G->smem = 0;
        G->grid = (int)std::max<long long>(1, std::min<long long>((B + 127) / 128, (long long)c->num_sms * 8));   // __launch_bounds__(128, 8)
        const size_t per_thread = (size_t)tsize * (2 * (size_t)c->E + 2 * (size_t)c->m) + 4 * (size_t)(c->WN + c->WM);
        G->gstate_bytes = per_thread * (size_t)G->grid * G->threads;
    }
    return QLDPC_OK;
}

extern "C" int qldpc_bp_geometry(qldpc_code *c, const qldpc_bp_config *cfg, int32_t *shots_per_cta, int32_t *smem_bytes,
                                 int32_t *staged)
{
    if (!c || !cfg) return fail(QLDPC_ERR_ARG, "qldpc_bp_geometry: null argument");
    BPGeom G;
    bp_geometry(c, cfg, 1ll << 40, &G);
    if (shots_per_cta) *shots_per_cta = G.shots_per_cta;
    if (smem_bytes) *smem_bytes = (int32_t)G.smem;
    if (staged) *staged = G.staged ? 1 : (G.stage_kernel ? 134 : (G.cta_kernel ? 133 : (G.warp_kernel ? 132 : (G.tiled_T ? 100 + G.tiled_T : 0))));
    return QLDPC_OK;
}

extern "C" int qldpc_warp_layout_tune(qldpc_code *c, int64_t steps, int32_t *cost)
{
    if (!c) return fail(QLDPC_ERR_ARG, "qldpc_warp_layout_tune: null code");
    if (!c->warp_ok) return fail(QLDPC_ERR_UNSUPPORTED, "qldpc_warp_layout_tune: the warp-per-shot kernel does not apply to this code");
    CK(cudaDeviceSynchronize());
    if (steps < 0) { c->wlayout->natural(); c->wlayout64->natural(); }
    if (steps > 0) { c->wlayout->construct(steps); c->wlayout64->construct(steps); }
    if (steps != 0)
        if (int rc = warp_tables_upload(c)) return rc;
    if (cost) { cost[0] = c->warp_cost[0]; cost[1] = c->warp_cost[1]; cost[2] = c->warp_cost[2]; }
    return QLDPC_OK;
}

extern "C" int qldpc_tiled_conflict_model(qldpc_code *c, int32_t lanes_per_shot, double *before, double *after)
{
    if (!c || !c->tiled_ok) return fail(QLDPC_ERR_UNSUPPORTED, "qldpc_tiled_conflict_model: the tiled kernel does not apply to this code");
    const int ti = lanes_per_shot == 4 ? 1 : 2;
    if (before) *before = c->tiled_conflict_cost[ti][0];
    if (after) *after = c->tiled_conflict_cost[ti][1];
    return QLDPC_OK;
}

static int check_cfg(const qldpc_bp_config *cfg)
{
    if (!cfg) return fail(QLDPC_ERR_ARG, "null BP config");
    if (cfg->variant < 0 || cfg->variant > 2) return fail(QLDPC_ERR_ARG, "unknown BP variant");
    if (cfg->precision != 32 && cfg->precision != 64) return fail(QLDPC_ERR_ARG, "precision must be 32 or 64");
    if (cfg->max_iter < 1) return fail(QLDPC_ERR_ARG, "max_iter must be >= 1");
    return QLDPC_OK;
}

static int set_prior(qldpc_code *c, const double *prior_host, cudaStream_t st)
{
    if (!prior_host) return fail(QLDPC_ERR_ARG, "null prior");
    if (c->prior_cache.size() == (size_t)c->n && memcmp(c->prior_cache.data(), prior_host, sizeof(double) * c->n) == 0)
        return QLDPC_OK;
    CK(c->prior64.reserve(sizeof(double) * c->n));
    CK(c->prior32.reserve(sizeof(float) * c->n));
    std::vector<float> pf(c->n);
    for (int i = 0; i < c->n; ++i) pf[i] = (float)prior_host[i];
    // both copies are from pageable memory: staged by the driver before the call returns
    CK(cudaMemcpyAsync(c->prior64.p, prior_host, sizeof(double) * c->n, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(c->prior32.p, pf.data(), sizeof(float) * c->n, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
    c->prior_cache.assign(prior_host, prior_host + c->n);
    c->prior_max = 0.0;
    c->prior_uniform = true;
    c->prior_positive = true;
    double log_zero = 0.0;
    for (int i = 0; i < c->n; ++i) {
        c->prior_positive = c->prior_positive && prior_host[i] > 0.0 && (float)prior_host[i] > 0.f;
        log_zero -= std::log1p(std::exp(-prior_host[i]));              // log(1 - p_i)
        c->prior_max = std::max(c->prior_max, std::fabs(prior_host[i]));
        c->prior_uniform = c->prior_uniform && (memcmp(&prior_host[i], &prior_host[0], sizeof(double)) == 0);
    }
    c->zero_frac = std::exp(log_zero);
    return QLDPC_OK;
}


static int bp_decode_impl(qldpc_code *c, const qldpc_bp_config *cfg, const double *prior_host, int64_t B,
                          const uint32_t *synd, uint32_t *hard, uint8_t *conv, int32_t *iters, void *llr,
                          int32_t llr_mode, int32_t *fail_idx, uint32_t *fail_count, uint64_t *iter_total,
                          DevBuf *ctrl_buf, DevBuf *gstate_buf, void *stream)
{
    if (!c || !synd || !hard || !conv) return fail(QLDPC_ERR_ARG, "qldpc_bp_decode_dev: null argument");
    if (int rc = check_cfg(cfg)) return rc;
    if (B <= 0) return QLDPC_OK;
    if (B > 0x7fffffffll) return fail(QLDPC_ERR_ARG, "qldpc_bp_decode_dev: B must be < 2^31 per call (chunk the batch)");
    cudaStream_t st = (cudaStream_t)stream;
    if (int rc = set_prior(c, prior_host, st)) return rc;
    BPGeom G;
    bp_geometry(c, cfg, B, &G);
    if (G.staged || G.stage_kernel) CK(gstate_buf->reserve(G.gstate_bytes));
    CK(ctrl_buf->reserve(sizeof(Ctrl)));
    Ctrl *ctrl = ctrl_buf->as<Ctrl>();
    CK(cudaMemsetAsync(ctrl, 0, sizeof(Ctrl), st));
    if (fail_count) CK(cudaMemsetAsync(fail_count, 0, sizeof(uint32_t), st));

    BPParams P;
    P.g = c->graph();
    P.B = B;
    P.synd = synd;
    P.prior = cfg->precision == 64 ? c->prior64.p : c->prior32.p;
    P.max_iter = cfg->max_iter;
    P.sym = (cfg->variant == QLDPC_SUM_PRODUCT_SYM);
    P.alpha = cfg->alpha;
    P.damping = cfg->damping;
    P.one_minus_damping = 1.0 - cfg->damping;     // `(1 - damping)` evaluated in float64 (decoding.py:65)
    P.clip = cfg->clip;
    P.qpad = std::max(cfg->clip, c->prior_max);
    P.prior_uniform = c->prior_uniform ? 1 : 0;
    // 2: launch the instantiation with the zero-syndrome shortcut -- worth its registers from ~30 % error-free shots on for min-sum
    // (measured: [[144,12,12]] at p = 0.01, 23 %: -2 %; [[108,8,10]], 34 %: +7 %), from 10 % on for the costlier sum-product iteration
    const double zero_min = (cfg->variant == QLDPC_MIN_SUM) ? 0.3 : 0.1;
    P.zero_ok = (c->prior_positive && cfg->alpha >= 0.0 && !getenv("QLDPC_NO_ZERO_SHORTCUT")) ? (c->zero_frac >= zero_min || getenv("QLDPC_FORCE_ZERO_SHORTCUT") ? 2 : 1) : 0;
    P.hard = hard;
    P.conv = conv;
    P.iters = iters;
    P.llr = llr;
    P.llr_mode = llr ? llr_mode : LLR_NONE;
    P.cursor = &ctrl->cursor;
    P.fail_idx = fail_idx;
    P.fail_count = fail_count ? fail_count : &ctrl->fail_count;
    P.iter_total = (unsigned long long *)iter_total;
    P.gstate = gstate_buf->p;
    P.r_dump = nullptr;
    P.dump_iter = -1;
    cudaError_t e;
    const int kv = kernel_variant(cfg->variant);
    if (G.stage_kernel)
        e = launch_bp_stage(c, P, G, cfg->precision, st);
    else if (G.cta_kernel)
        e = launch_bp_cta(c, P, G, st);
    else if (G.warp_kernel)
        e = launch_bp_warp(c, P, G, st);
    else if (G.tiled_T)
        e = launch_bp_tiled(c, P, G, cfg->precision, kv, st);
    else
        e = launch_bp_generic(P, G, cfg->precision, kv, st);
    if (e != cudaSuccess) return fail(QLDPC_ERR_CUDA, std::string("bp_decode_kernel launch: ") + cudaGetErrorString(e));
    return QLDPC_OK;
}

extern "C" int qldpc_bp_decode_dev(qldpc_code *c, const qldpc_bp_config *cfg, const double *prior_host, int64_t B,
                                   const uint32_t *synd, uint32_t *hard, uint8_t *conv, int32_t *iters, void *llr,
                                   int32_t llr_mode, int32_t *fail_idx, uint32_t *fail_count, uint64_t *iter_total,
                                   void *stream)
{
    if (!c) return fail(QLDPC_ERR_ARG, "qldpc_bp_decode_dev: null code");
    return bp_decode_impl(c, cfg, prior_host, B, synd, hard, conv, iters, llr, llr_mode, fail_idx, fail_count, iter_total,
                          &c->ctrl, &c->gstate, stream);
}

// ------------------------------------------------------------------------------------------------
// OSD
// ------------------------------------------------------------------------------------------------

extern "C" int qldpc_osd_decode_dev(qldpc_code *c, const int32_t *idx, const uint32_t *count_dev, int64_t count_host,
                                    const uint32_t *synd, const void *llr, int32_t llr_f64, const uint32_t *hard,
                                    uint32_t *out, uint8_t *valid, void *stream)
{
    if (!c || !synd || !llr || !hard || !out) return fail(QLDPC_ERR_ARG, "qldpc_osd_decode_dev: null argument");
    if (!count_dev && count_host <= 0) return QLDPC_OK;
    OSDParams P;
    memset(&P, 0, sizeof(P));
    P.idx = idx;
    P.count_dev = count_dev;
    P.count_host = count_host;
    P.synd = synd; P.llr = llr; P.hard = hard; P.out = out; P.valid = valid;
    P.cap = count_dev ? count_host : 0;          // with a device-side count, count_host (if > 0) bounds it
    return osd_launch(c, P, llr_f64, count_dev ? -1 : count_host, (cudaStream_t)stream, &c->ws_redo);
}

// ------------------------------------------------------------------------------------------------
// small kernels
// ------------------------------------------------------------------------------------------------
static int grid_for(long long work, int threads, int num_sms)
{
    long long g = (work + threads - 1) / threads;
    return (int)std::max<long long>(1, std::min<long long>(g, (long long)num_sms * 16));
}

extern "C" int qldpc_pack_bits_dev(const uint8_t *in, uint32_t *out, int64_t B, int32_t nbits, void *stream)
{
    if (B <= 0) return QLDPC_OK;
    const int W = (nbits + 31) / 32;
    pack_bits_kernel<<<grid_for(B * W, 256, 148), 256, 0, (cudaStream_t)stream>>>(in, out, B, nbits, W);
    CK(cudaGetLastError());
    return QLDPC_OK;
}
extern "C" int qldpc_unpack_bits_dev(const uint32_t *in, uint8_t *out, int64_t B, int32_t nbits, void *stream)
{
    if (B <= 0) return QLDPC_OK;
    const int W = (nbits + 31) / 32;
    if (nbits % 16 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0)
        unpack_bits16_kernel<<<grid_for(B * (long long)(nbits / 16), 256, 148), 256, 0, (cudaStream_t)stream>>>(
            in, reinterpret_cast<uint4 *>(out), B, nbits, W);
    else
        unpack_bits_kernel<<<grid_for(B * (long long)nbits, 256, 148), 256, 0, (cudaStream_t)stream>>>(in, out, B, nbits, W);
    CK(cudaGetLastError());
    return QLDPC_OK;
}

template <int WM> static void launch_sample(const SampleParams &P, int grid, cudaStream_t st) { sample_kernel<WM><<<grid, 128, 0, st>>>(P); }
template <int WM> static void launch_check(const CheckParams &P, int grid, cudaStream_t st) { check_kernel<WM><<<grid, 256, 0, st>>>(P); }

extern "C" int qldpc_sample_dev(qldpc_code *c, double p, uint64_t seed, uint64_t first_shot, int32_t draws, int64_t B,
                                uint32_t *err, uint32_t *synd, void *stream)
{
    if (!c || !err || !synd) return fail(QLDPC_ERR_ARG, "qldpc_sample_dev: null argument");
    if (!(p >= 0.0 && p < 1.0) || draws < 1 || draws > 2) return fail(QLDPC_ERR_ARG, "qldpc_sample_dev: need 0 <= p < 1 and draws in {1,2}");
    if (B <= 0) return QLDPC_OK;
    SampleParams P;
    P.m = c->m; P.n = c->n; P.WM = c->WM; P.WN = c->WN;
    P.colmask = c->d_colmask;
    P.B = B; P.first_shot = first_shot; P.seed = seed;
    P.threshold = (uint32_t)std::min<double>(4294967295.0, p * 4294967296.0);
    P.draws = draws;
    P.err = err; P.synd = synd;
    const int grid = (int)((B + 127) / 128);
    cudaStream_t st = (cudaStream_t)stream;
    switch (c->WM) {
    case 1: launch_sample<1>(P, grid, st); break;
    case 2: launch_sample<2>(P, grid, st); break;
    case 3: launch_sample<3>(P, grid, st); break;
    case 4: launch_sample<4>(P, grid, st); break;
    case 5: launch_sample<5>(P, grid, st); break;
    default:
        launch_sample<0>(P, grid, st);                       // errors only ...
        CK(cudaGetLastError());
        return qldpc_syndrome_dev(c, B, err, synd, stream);  // ... syndromes from the CSR rows
    }
    CK(cudaGetLastError());
    return QLDPC_OK;
}

extern "C" int qldpc_measurement_noise_dev(qldpc_code *c, double q, uint64_t seed, uint64_t first_shot, int64_t B, uint32_t *synd,
                                           void *stream)
{
    if (!c || !synd) return fail(QLDPC_ERR_ARG, "qldpc_measurement_noise_dev: null argument");
    if (!(q >= 0.0 && q < 1.0)) return fail(QLDPC_ERR_ARG, "qldpc_measurement_noise_dev: need 0 <= q < 1");
    if (B <= 0 || q == 0.0) return QLDPC_OK;
    const uint32_t thr = (uint32_t)std::min<double>(4294967295.0, q * 4294967296.0);
    meas_noise_kernel<<<grid_for(B * c->WM, 256, c->num_sms), 256, 0, (cudaStream_t)stream>>>(synd, B, c->m, c->WM, thr, seed, first_shot);
    CK(cudaGetLastError());
    return QLDPC_OK;
}

extern "C" int qldpc_check_dev(qldpc_code *c, int64_t B, const uint32_t *err, const uint32_t *corr, const uint32_t *synd,
                               const uint8_t *conv, const int32_t *iters, int32_t distance, uint8_t *flags,
                               int32_t *weight, uint64_t *counters_dev, void *stream)
{
    if (!c || !err || !corr || !synd) return fail(QLDPC_ERR_ARG, "qldpc_check_dev: null argument");
    if (B <= 0) return QLDPC_OK;
    CheckParams P;
    P.m = c->m; P.n = c->n; P.k = c->k; P.WM = c->WM; P.WN = c->WN;
    P.colmask = c->d_colmask; P.Lrows = c->d_Lrows;
    P.B = B; P.err = err; P.corr = corr; P.synd = synd; P.conv = conv; P.iters = iters;
    P.half_distance = distance / 2;
    P.counters = (unsigned long long *)counters_dev;
    P.flags = flags; P.weight = weight;
    P.corr_synd = nullptr;
    const int grid = grid_for(B, 256, c->num_sms);
    cudaStream_t st = (cudaStream_t)stream;
    if (c->WM > 5) {                                         // large H: syndrome of the correction from the CSR rows first
        CK(c->ws_valid.reserve(4 * (size_t)B * c->WM));
        if (int rc = qldpc_syndrome_dev(c, B, corr, c->ws_valid.as<uint32_t>(), stream)) return rc;
        P.corr_synd = c->ws_valid.as<uint32_t>();
        launch_check<0>(P, grid, st);
        CK(cudaGetLastError());
        return QLDPC_OK;
    }
    switch (c->WM) {
    case 1: launch_check<1>(P, grid, st); break;
    case 2: launch_check<2>(P, grid, st); break;
    case 3: launch_check<3>(P, grid, st); break;
    case 4: launch_check<4>(P, grid, st); break;
    default: launch_check<5>(P, grid, st); break;
    }
    CK(cudaGetLastError());
    return QLDPC_OK;
}

// ------------------------------------------------------------------------------------------------
// OSD-w sweep stage (performOSD_enhanced with order > 0, OSD_enhanced.py:58-131) behind an OSD-0 launch that wrote `valid`:
// the shots of the list whose OSD-0 solution misses the syndrome are compacted on the device and swept by osdw_kernel,
// with the count read from device memory (no host synchronisation).  hard == null: BP hard decision = llr < 0.
// ------------------------------------------------------------------------------------------------
static int osdw_supported(const qldpc_code *c, int32_t order)
{
    if (order <= 0) return QLDPC_OK;
    if (order > OSDW_MAX_ORDER) return fail(QLDPC_ERR_ARG, "OSD order must be <= 16");
    // block-per-shot OSD (m > 160): no sweep kernel.  A full-rank H makes every syndrome consistent, so the sweep is dead
    // code there (OSD_enhanced.py:58-60) and OSD-0 is the exact answer; a rank-deficient one is refused.
    if (osd_use_block(c) && c->rank < c->m)
        return fail(QLDPC_ERR_UNSUPPORTED, "OSD-w sweep (order > 0) is not available for rank-deficient check matrices with more than 160 rows");
    return QLDPC_OK;
}

static int osdw_stage(qldpc_code *c, const int32_t *idx, const uint32_t *count_dev, long long count_host, long long cap,
                      const uint32_t *synd, const void *llr, int llr_f64, const uint32_t *hard, uint32_t *out, const uint8_t *valid,
                      int32_t order, int64_t max_combinations, DevBuf *inv_buf, cudaStream_t st)
{
    if (order <= 0 || osd_use_block(c) || cap <= 0) return QLDPC_OK;
    CK(inv_buf->reserve(sizeof(int32_t) * (size_t)cap + 16));
    unsigned int *inv_count = inv_buf->as<unsigned int>();
    int32_t *inv_idx = inv_buf->as<int32_t>() + 4;
    CK(cudaMemsetAsync(inv_count, 0, sizeof(unsigned int), st));
    compact_invalid_kernel<<<(int)std::max<long long>(1, std::min<long long>((cap + 255) / 256, (long long)c->num_sms * 8)), 256, 0, st>>>(
        idx, count_dev, count_host, valid, inv_idx, inv_count);
    CK(cudaGetLastError());
    OSDWParams W;
    memset(&W, 0, sizeof(W));
    W.m = c->m; W.n = c->n; W.WM = c->WM; W.WN = c->WN; W.rank = c->rank;
    W.Hrows = c->d_Hrows; W.colmask = c->d_colmask;
    W.idx = inv_idx; W.count_dev = inv_count; W.count = 0;
    W.synd = synd; W.llr = llr; W.llr_f32 = llr_f64 ? 0 : 1; W.hard = hard;
    W.sol = out; W.valid = valid;
    W.order = order;
    W.max_combinations = max_combinations > 0 ? max_combinations : 0;
    cudaError_t e = launch_osdw(W, c->num_sms, st);
    if (e != cudaSuccess) return fail(QLDPC_ERR_CUDA, std::string("osdw_kernel launch: ") + cudaGetErrorString(e));
    return QLDPC_OK;
}

extern "C" int qldpc_osdw_decode_dev(qldpc_code *c, const int32_t *idx, const uint32_t *count_dev, int64_t count_host,
                                     const uint32_t *synd, const void *llr, int32_t llr_f64, const uint32_t *hard, uint32_t *out,
                                     const uint8_t *valid, int32_t order, int64_t max_combinations, void *stream)
{
    if (!c || !synd || !llr || !out || !valid) return fail(QLDPC_ERR_ARG, "qldpc_osdw_decode_dev: null argument");
    if (!count_dev && count_host <= 0) return QLDPC_OK;
    if (count_dev && count_host <= 0) return fail(QLDPC_ERR_ARG, "qldpc_osdw_decode_dev: count_host must bound a device-side count");
    if (int rc = osdw_supported(c, order)) return rc;
    return osdw_stage(c, idx, count_dev, count_dev ? 0 : count_host, count_host, synd, llr, llr_f64, hard, out, valid, order,
                      max_combinations, &c->ws_inv, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------
// fused BP -> OSD
// ------------------------------------------------------------------------------------------------
static const long long CHUNK = 1ll << 24;   // shots per internal launch (bounds the LLR hand-off workspace: 4n bytes per shot)

// BP, then OSD-0 on the compacted BP failures -- and, for osd_order > 0, the OSD-w sweep on those whose OSD-0 solution misses
// the syndrome (never the case for syndromes of the form e * H^T) -- for at most CHUNK shots, with explicit workspaces
static int bposd_chunk(qldpc_code *c, const qldpc_bp_config *cfg, const double *prior_host, long long b, const uint32_t *synd,
                       int32_t osd_order, uint32_t *corr, uint8_t *conv, int32_t *iters, uint64_t *iter_total,
                       DevBuf *ctrl, DevBuf *gstate, DevBuf *llr_buf, DevBuf *fail_buf, DevBuf *redo_buf, DevBuf *valid_buf, DevBuf *inv_buf,
                       cudaStream_t st)
{
    const int tsize = cfg->precision == 64 ? 8 : 4;
    void *llr = nullptr;
    int32_t *fidx = nullptr;
    uint32_t *fcnt = nullptr;
    if (int rc = osdw_supported(c, osd_order)) return rc;
    if (osd_order >= 0) {
        CK(llr_buf->reserve((size_t)b * c->n * tsize));
        CK(fail_buf->reserve(sizeof(int32_t) * (size_t)b + 16));
        llr = llr_buf->p;
        fcnt = fail_buf->as<uint32_t>();
        fidx = fail_buf->as<int32_t>() + 4;
    }
    int rc = bp_decode_impl(c, cfg, prior_host, b, synd, corr, conv, iters, llr, QLDPC_LLR_FAILED, fidx, fcnt, iter_total,
                            ctrl, gstate, st);
    if (rc) return rc;
    if (osd_order >= 0) {
        const bool sweep = osd_order > 0 && !osd_use_block(c);
        OSDParams P;
        memset(&P, 0, sizeof(P));
        P.idx = fidx; P.count_dev = fcnt; P.count_host = 0;
        P.synd = synd;
        P.llr = llr;
        P.hard = corr;
        P.out = corr;
        if (sweep) {                       // (the flag is written for the listed shots only; the sweep reads no others)
            CK(valid_buf->reserve((size_t)b));
            P.valid = valid_buf->as<uint8_t>();
        }
        P.cap = b;
        rc = osd_launch(c, P, tsize == 8, -1, st, redo_buf);
        if (rc) return rc;
        if (sweep)
            if ((rc = osdw_stage(c, fidx, fcnt, 0, b, synd, llr, tsize == 8, nullptr, corr, P.valid, osd_order, 0, inv_buf, st))) return rc;
    }
    return QLDPC_OK;
}

extern "C" int qldpc_bposd_decode_dev(qldpc_code *c, const qldpc_bp_config *cfg, const double *prior_host, int64_t B,
                                      const uint32_t *synd, int32_t osd_order, uint32_t *corr, uint8_t *conv,
                                      int32_t *iters, uint64_t *iter_total, void *stream)
{
    if (!c || !synd || !corr || !conv) return fail(QLDPC_ERR_ARG, "qldpc_bposd_decode_dev: null argument");
    if (int rc = check_cfg(cfg)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    for (long long o = 0; o < B; o += CHUNK) {
        const long long b = std::min<long long>(CHUNK, B - o);
        if (int rc = bposd_chunk(c, cfg, prior_host, b, synd + (size_t)o * c->WM, osd_order, corr + (size_t)o * c->WN, conv + o,
                                 iters ? iters + o : nullptr, iter_total, &c->ctrl, &c->gstate, &c->ws_llr, &c->ws_fail, &c->ws_redo, &c->ws_valid, &c->ws_inv, st))
            return rc;
    }
    return QLDPC_OK;
}

// ------------------------------------------------------------------------------------------------
// host-pointer API
// ------------------------------------------------------------------------------------------------
extern "C" int qldpc_bp_decode_host(qldpc_code *c, const qldpc_bp_config *cfg, const double *prior, int64_t B,
                                    const uint8_t *synd, int8_t *hard, uint8_t *conv, int32_t *iters, double *llr)
{
    if (!c || !synd || !hard || !conv) return fail(QLDPC_ERR_ARG, "qldpc_bp_decode_host: null argument");
    if (int rc = check_cfg(cfg)) return rc;
    cudaStream_t st = 0;
    const int tsize = cfg->precision == 64 ? 8 : 4;
    const long long chunk = std::min<long long>(CHUNK, 1ll << 20);
    for (long long o = 0; o < B; o += chunk) {
        const long long b = std::min<long long>(chunk, B - o);
        CK(c->ws_u8a.reserve((size_t)b * std::max(c->m, c->n)));
        CK(c->ws_synd.reserve(4 * (size_t)b * c->WM));
        CK(c->ws_hard.reserve(4 * (size_t)b * c->WN));
        CK(c->ws_conv.reserve((size_t)b));
        CK(c->ws_iters.reserve(4 * (size_t)b));
        if (llr) CK(c->ws_llr.reserve((size_t)b * c->n * 8));   // room for the float64 copy
        CK(cudaMemcpyAsync(c->ws_u8a.p, synd + (size_t)o * c->m, (size_t)b * c->m, cudaMemcpyHostToDevice, st));
        if (int rc = qldpc_pack_bits_dev(c->ws_u8a.as<uint8_t>(), c->ws_synd.as<uint32_t>(), b, c->m, st)) return rc;
        void *dllr = nullptr;
        if (llr) {
            // float32 results are widened on the device into the same buffer, back to front is unsafe: use a second one
            if (tsize == 4) { CK(c->ws_llr_in.reserve((size_t)b * c->n * 4)); dllr = c->ws_llr_in.p; }
            else dllr = c->ws_llr.p;
        }
        if (int rc = qldpc_bp_decode_dev(c, cfg, prior, b, c->ws_synd.as<uint32_t>(), c->ws_hard.as<uint32_t>(),
                                         c->ws_conv.as<uint8_t>(), c->ws_iters.as<int32_t>(), dllr, QLDPC_LLR_ALL, nullptr,
                                         nullptr, nullptr, st))
            return rc;
        if (int rc = qldpc_unpack_bits_dev(c->ws_hard.as<uint32_t>(), c->ws_u8a.as<uint8_t>(), b, c->n, st)) return rc;
        CK(cudaMemcpyAsync(hard + (size_t)o * c->n, c->ws_u8a.p, (size_t)b * c->n, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(conv + o, c->ws_conv.p, (size_t)b, cudaMemcpyDeviceToHost, st));
        if (iters) CK(cudaMemcpyAsync(iters + o, c->ws_iters.p, 4 * (size_t)b, cudaMemcpyDeviceToHost, st));
        if (llr) {
            if (tsize == 4) {
                cast_kernel<float, double><<<grid_for(b * c->n, 256, c->num_sms), 256, 0, st>>>(
                    c->ws_llr_in.as<float>(), c->ws_llr.as<double>(), b * (long long)c->n);
                CK(cudaGetLastError());
            }
            CK(cudaMemcpyAsync(llr + (size_t)o * c->n, c->ws_llr.p, (size_t)b * c->n * 8, cudaMemcpyDeviceToHost, st));
        }
        CK(cudaStreamSynchronize(st));
    }
    return QLDPC_OK;
}

// alpha_estimation=True return paths of the reference: check-to-variable messages per edge
extern "C" int qldpc_bp_messages_host(qldpc_code *c, const qldpc_bp_config *cfg, const double *prior, int64_t B,
                                      const uint8_t *synd, int32_t dump_iter, double *r_edges)
{
    if (!c || !synd || !r_edges) return fail(QLDPC_ERR_ARG, "qldpc_bp_messages_host: null argument");
    if (int rc = check_cfg(cfg)) return rc;
    if (cfg->precision != 64) return fail(QLDPC_ERR_ARG, "qldpc_bp_messages_host: precision must be 64");
    if (dump_iter < 0 || dump_iter >= cfg->max_iter) return fail(QLDPC_ERR_ARG, "qldpc_bp_messages_host: dump_iter must be in [0, max_iter)");
    cudaStream_t st = 0;
    if (int rc = set_prior(c, prior, st)) return rc;
    const long long chunk = 1ll << 16;
    for (long long o = 0; o < B; o += chunk) {
        const long long b = std::min<long long>(chunk, B - o);
        CK(c->ws_u8a.reserve((size_t)b * std::max(c->m, c->n)));
        CK(c->ws_synd.reserve(4 * (size_t)b * c->WM));
        CK(c->ws_hard.reserve(4 * (size_t)b * c->WN));
        CK(c->ws_conv.reserve((size_t)b));
        CK(c->ws_rec.reserve(8 * (size_t)b * c->E));
        CK(cudaMemcpyAsync(c->ws_u8a.p, synd + (size_t)o * c->m, (size_t)b * c->m, cudaMemcpyHostToDevice, st));
        if (int rc = qldpc_pack_bits_dev(c->ws_u8a.as<uint8_t>(), c->ws_synd.as<uint32_t>(), b, c->m, st)) return rc;
        CK(cudaMemsetAsync(c->ws_rec.p, 0, 8 * (size_t)b * c->E, st));
        // same launch path as qldpc_bp_decode_dev, thread-per-shot kernel, with the dump enabled
        BPGeom G;
        qldpc_bp_config cf = *cfg;
        if (cf.staged == 0 || cf.staged == 3 || cf.staged == 4) cf.staged = 2;
        bp_geometry(c, &cf, b, &G);
        if (G.staged) CK(c->gstate.reserve(G.gstate_bytes));
        CK(c->ctrl.reserve(sizeof(Ctrl)));
        Ctrl *ctrl = c->ctrl.as<Ctrl>();
        CK(cudaMemsetAsync(ctrl, 0, sizeof(Ctrl), st));
        BPParams P;
        P.g = c->graph();
        P.B = b;
        P.synd = c->ws_synd.as<uint32_t>();
        P.prior = c->prior64.p;
        P.max_iter = cf.max_iter;
        P.sym = (cf.variant == QLDPC_SUM_PRODUCT_SYM);
        P.alpha = cf.alpha; P.damping = cf.damping; P.one_minus_damping = 1.0 - cf.damping; P.clip = cf.clip; P.qpad = std::max(cf.clip, c->prior_max); P.prior_uniform = 0; P.zero_ok = 0;
        P.hard = c->ws_hard.as<uint32_t>(); P.conv = c->ws_conv.as<uint8_t>(); P.iters = nullptr;
        P.llr = nullptr; P.llr_mode = LLR_NONE;
        P.cursor = &ctrl->cursor; P.fail_idx = nullptr; P.fail_count = &ctrl->fail_count; P.iter_total = nullptr;
        P.gstate = c->gstate.p;
        P.r_dump = c->ws_rec.p;
        P.dump_iter = dump_iter;
        cudaError_t e = launch_bp_generic(P, G, 64, kernel_variant(cf.variant), st);
        if (e != cudaSuccess) return fail(QLDPC_ERR_CUDA, std::string("bp_decode_kernel (messages) launch: ") + cudaGetErrorString(e));
        CK(cudaMemcpyAsync(r_edges + (size_t)o * c->E, c->ws_rec.p, 8 * (size_t)b * c->E, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
    }
    return QLDPC_OK;
}

extern "C" int qldpc_osd_decode_host(qldpc_code *c, int64_t B, const uint8_t *synd, const double *llr, const uint8_t *hard,
                                     int32_t order, int64_t max_combinations, uint8_t *out)
{
    if (!c || !synd || !llr || !hard || !out) return fail(QLDPC_ERR_ARG, "qldpc_osd_decode_host: null argument");
    if (int rc = osdw_supported(c, order)) return rc;
    cudaStream_t st = 0;
    const long long chunk = 1ll << 18;
    for (long long o = 0; o < B; o += chunk) {
        const long long b = std::min<long long>(chunk, B - o);
        CK(c->ws_u8a.reserve((size_t)b * std::max(c->m, c->n)));
        CK(c->ws_u8b.reserve((size_t)b * c->n));
        CK(c->ws_synd.reserve(4 * (size_t)b * c->WM));
        CK(c->ws_hard.reserve(4 * (size_t)b * c->WN));
        CK(c->ws_err.reserve(4 * (size_t)b * c->WN));
        CK(c->ws_valid.reserve((size_t)b));
        CK(c->ws_llr.reserve((size_t)b * c->n * 8));
        CK(cudaMemcpyAsync(c->ws_u8a.p, synd + (size_t)o * c->m, (size_t)b * c->m, cudaMemcpyHostToDevice, st));
        if (int rc = qldpc_pack_bits_dev(c->ws_u8a.as<uint8_t>(), c->ws_synd.as<uint32_t>(), b, c->m, st)) return rc;
        CK(cudaMemcpyAsync(c->ws_u8b.p, hard + (size_t)o * c->n, (size_t)b * c->n, cudaMemcpyHostToDevice, st));
        if (int rc = qldpc_pack_bits_dev(c->ws_u8b.as<uint8_t>(), c->ws_hard.as<uint32_t>(), b, c->n, st)) return rc;
        CK(cudaMemcpyAsync(c->ws_llr.p, llr + (size_t)o * c->n, (size_t)b * c->n * 8, cudaMemcpyHostToDevice, st));
        OSDParams P;
        memset(&P, 0, sizeof(P));
        P.idx = nullptr; P.count_dev = nullptr; P.count_host = b;
        P.synd = c->ws_synd.as<uint32_t>();
        P.llr = c->ws_llr.p;
        P.hard = c->ws_hard.as<uint32_t>();
        P.out = c->ws_err.as<uint32_t>();
        P.valid = c->ws_valid.as<uint8_t>();
        if (int rc = osd_launch(c, P, 1, b, st, &c->ws_redo)) return rc;
        // OSD-w sweep on the shots whose OSD-0 solution misses the syndrome (OSD_enhanced.py:58-131)
        if (int rc = osdw_stage(c, nullptr, nullptr, b, b, P.synd, P.llr, 1, P.hard, P.out, P.valid, order, max_combinations, &c->ws_inv, st))
            return rc;
        if (int rc = qldpc_unpack_bits_dev(c->ws_err.as<uint32_t>(), c->ws_u8b.as<uint8_t>(), b, c->n, st)) return rc;
        CK(cudaMemcpyAsync(out + (size_t)o * c->n, c->ws_u8b.p, (size_t)b * c->n, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
    }
    return QLDPC_OK;
}

// Who packs the uint8 rows of qldpc_bposd_decode_host: the device (byte rows cross PCIe: n + m + 5 bytes per shot) or host
// threads (bit-packed rows cross it: 4 (WM + WN) + 5 bytes).  QLDPC_HOST_PACK = 0 / 1 forces a side; otherwise the pool's pack +
// unpack throughput is measured once per code handle on a 2^16-shot sample and the host side is taken when it clearly
// outruns what the bus would carry (D2H ~50 GB/s, H2D ~25 GB/s when both directions are busy: measured, tools/pcie_bw.py).
// QLDPC_HOST_THREADS sets the pool size (default: hardware threads / visible GPUs, at most 16).
static int host_pack_decide(qldpc_code *c, int want)        // want: -1 measure (the environment may force a side), 0 device, 1 host
{
    if (want < 0)
        if (const char *e = getenv("QLDPC_HOST_PACK")) want = (e[0] == '0') ? 0 : (e[0] == '1') ? 1 : (e[0] == '2') ? 2 : -1;
    if (want == 0) return c->host_pack = 0;
    if (!c->pool) {
        int threads = 0;
        if (const char *t = getenv("QLDPC_HOST_THREADS")) threads = atoi(t);
        if (threads <= 0) {
            int ndev = 1;
            if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) ndev = 1;
            threads = (int)std::min<unsigned>(16u, std::max(1u, std::thread::hardware_concurrency() / (unsigned)ndev));
        }
        if (want < 0 && threads < 4) return c->host_pack = 0;
        c->pool = new HostPool(threads);
        const long long Bs = 1ll << 16;
        std::vector<uint8_t> a((size_t)Bs * c->m, 1), o((size_t)Bs * c->n + 16);
        std::vector<uint32_t> pa((size_t)Bs * c->WM), po((size_t)Bs * c->WN, 0x55555555u);
        uint8_t *oal = o.data() + ((16 - (reinterpret_cast<uintptr_t>(o.data()) & 15u)) & 15u);
        double best = 1e30;
        for (int rep = 0; rep < 3; ++rep) {
            const auto t0 = std::chrono::steady_clock::now();
            host_pack_rows(*c->pool, a.data(), pa.data(), Bs, c->m, c->WM);
            host_unpack_rows(*c->pool, po.data(), oal, Bs, c->n, c->WN);
            best = std::min(best, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
        }
        c->host_pack_rate = (double)Bs / best;
    }
    c->est_pack = (1.0 / 3.0) / c->host_pack_rate;
    c->est_unpack = (2.0 / 3.0) / c->host_pack_rate;
    c->est_dev = (c->n + 5.0) / 50e9;
    if (want >= 1) return c->host_pack = want;
    // Host threads when the pool comes within reach of what the bus would carry at its single-GPU best: measured on the 1-, 2-
    // and 8-GPU boxes of this pool (profiles/r2t_*), the host side wins or ties wherever its pool has >= 4 threads -- on the
    // 8-GPU box the byte rows of the 8 ranks share 93 GB/s of device-to-host bandwidth (57 GB/s for one GPU alone), and the
    // bus assumed here is far from what a rank gets.  (Mode 2, chunk by chunk on whichever side is free, is kept selectable.)
    const double bus_rate = 1.0 / std::max((c->n + 5.0) / 50e9, c->m / 25e9);
    return c->host_pack = (c->host_pack_rate >= 0.6 * bus_rate) ? 1 : 0;
}
static int host_pack_mode(qldpc_code *c) { return c->host_pack >= 0 ? c->host_pack : host_pack_decide(c, -1); }

// packed = false: synd [B][m] / corr [B][n] uint8 (the reference's dtypes);  packed = true: bit-packed uint32 rows
static int bposd_decode_host_impl(qldpc_code *c, const qldpc_bp_config *cfg, const double *prior, int64_t B,
                                  const void *synd_v, int32_t osd_order, void *corr_v, uint8_t *conv, int32_t *iters, bool packed)
{
    const uint8_t *synd = reinterpret_cast<const uint8_t *>(synd_v);
    uint8_t *corr = reinterpret_cast<uint8_t *>(corr_v);
    if (!c || !synd || !corr || !conv) return fail(QLDPC_ERR_ARG, "qldpc_bposd_decode_host: null argument");
    const size_t in_row = packed ? 4 * (size_t)c->WM : (size_t)c->m, out_row = packed ? 4 * (size_t)c->WN : (size_t)c->n;
    if (int rc = check_cfg(cfg)) return rc;
    if (B <= 0) return QLDPC_OK;
    // Copy-in / (pack, BP, OSD, unpack) / copy-out of consecutive chunks run on three streams chained by events (true
    // overlap needs pinned host buffers; pageable ones still work).
    const int hp_mode = packed ? 0 : host_pack_mode(c);       // 0: device, 1: host threads, 2: chunk by chunk
    // (measured, 10^7 [[144,12,12]] shots: host-packed rows 2.98e8 shots/s with 2^20-shot chunks, 3.09e8 with 2^21 -- fewer kernel
    //  tails; byte rows over PCIe 2.77e8 / 2.65e8 -- their first copy-in and last copy-out are not hidden)
    long long chunk = (hp_mode == 1) ? (1ll << 21) : (1ll << 20);
    if (const char *e = getenv("QLDPC_HOST_CHUNK")) chunk = std::max<long long>(1024, atoll(e));
    chunk = std::min<long long>(chunk, CHUNK);
    if (int rc = set_prior(c, prior, 0)) return rc;
    // enqueue every chunk; on any failure stop enqueuing, but always drain the streams before returning, so that no copy
    // into the caller's buffers is still in flight
    // QLDPC_TRACE=1: per-chunk timeline (ms since the first enqueue) on stderr -- diagnostic for the overlap of copies and kernels
    const bool trace = getenv("QLDPC_TRACE") != nullptr;
    std::vector<cudaEvent_t> tev;
    auto mark = [&](cudaStream_t st) {
        if (!trace) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, st);
        tev.push_back(e);
    };
    if (!c->st_in) {
        CK(cudaStreamCreateWithFlags(&c->st_in, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&c->st_comp, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&c->st_out, cudaStreamNonBlocking));
    }
    for (auto &sl : c->slot) { sl.used = false; sl.pend = false; }
    // host-side packing (uint8 rows only): host threads pack the syndromes of a chunk into the slot's pinned buffer before its
    // copy-in and expand its corrections after its copy-out, while the GPU works on the neighbouring chunks
    for (auto &sl : c->slot) { sl.inflight = false; sl.host_mode = false; }
    if (hp_mode == 2)                                         // either kind of chunk may land in any slot: no allocation on the way
        for (auto &sl : c->slot) {
            const long long bmax = std::min<long long>(chunk, B);
            CK(sl.h_synd.reserve(4 * (size_t)bmax * c->WM));
            CK(sl.h_hard.reserve(4 * (size_t)bmax * c->WN));
            CK(sl.u8in.reserve((size_t)bmax * c->m));
            CK(sl.u8out.reserve((size_t)bmax * c->n));
        }
    auto seconds = [](std::chrono::steady_clock::time_point t0) { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); };
    auto finish = [&](qldpc_code::Slot &sl) {                // corrections of the slot's pending chunk -> the caller's rows
        const auto t0 = std::chrono::steady_clock::now();
        host_unpack_rows(*c->pool, sl.h_hard.as<uint32_t>(), corr + (size_t)sl.pend_o * c->n, sl.pend_b, c->n, c->WN);
        if (sl.pend_b >= 4096) c->est_unpack = 0.75 * c->est_unpack + 0.25 * seconds(t0) / (double)sl.pend_b;
        sl.pend = false;
    };
    // byte-row copy-outs still queued, in seconds of bus time; completed ones update the running estimate of the bus
    auto bus_backlog = [&]() {
        double t = 0.0;
        for (auto &s2 : c->slot) {
            if (!s2.inflight || s2.host_mode) continue;
            if (cudaEventQuery(s2.ev_out) == cudaSuccess) {
                float ms = 0.f;
                if (s2.cur_b >= 4096 && cudaEventElapsedTime(&ms, s2.ev_out0, s2.ev_out) == cudaSuccess && ms > 0.f)
                    c->est_dev = 0.75 * c->est_dev + 0.25 * (1e-3 * ms / (double)s2.cur_b);
                s2.inflight = false;
            } else {
                t += (double)s2.cur_b * c->est_dev;
            }
        }
        return t;
    };
    long long i = 0;
    auto finish_ready = [&]() {                               // oldest first, as far as their copy-out has completed
        for (int a = 0; a < qldpc_code::NSLOT; ++a) {
            qldpc_code::Slot &sl = c->slot[(i + a) % qldpc_code::NSLOT];
            if (sl.pend && cudaEventQuery(sl.ev_out) == cudaSuccess) finish(sl);
        }
    };
    auto enqueue = [&](long long o, long long b, qldpc_code::Slot &sl) -> int {
        if (!sl.ev_in) {
            CK(cudaEventCreateWithFlags(&sl.ev_in, cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&sl.ev_comp, cudaEventDisableTiming));
            CK(cudaEventCreate(&sl.ev_out0));
            CK(cudaEventCreate(&sl.ev_out));
        }
        CK(sl.synd.reserve(4 * (size_t)b * c->WM));
        CK(sl.hard.reserve(4 * (size_t)b * c->WN));
        CK(sl.conv.reserve((size_t)b));
        CK(sl.iters.reserve(4 * (size_t)b));
        bool hostpack = (hp_mode == 1);
        if (hp_mode == 2) hostpack = bus_backlog() >= 0.9 * (double)b * (c->est_pack + c->est_unpack);
        const size_t in_bus = hostpack ? 4 * (size_t)c->WM : in_row, out_bus = hostpack ? 4 * (size_t)c->WN : out_row;
        if (hp_mode != 0) finish_ready();
        if (sl.pend) {                                        // the slot's previous chunk was packed by the host: expand it first
            CK(cudaEventSynchronize(sl.ev_out));
            finish(sl);
        }
        if (hostpack) {
            CK(sl.h_synd.reserve(4 * (size_t)b * c->WM));
            CK(sl.h_hard.reserve(4 * (size_t)b * c->WN));
            const auto t0 = std::chrono::steady_clock::now();
            host_pack_rows(*c->pool, synd + (size_t)o * c->m, sl.h_synd.as<uint32_t>(), b, c->m, c->WM);
            if (b >= 4096) c->est_pack = 0.75 * c->est_pack + 0.25 * seconds(t0) / (double)b;
            ++c->host_chunks;
        } else if (!packed) {
            ++c->dev_chunks;
            CK(sl.u8in.reserve((size_t)b * c->m));
            CK(sl.u8out.reserve((size_t)b * c->n));
        }
        // ---- copy in (the slot's input buffer is free once the previous chunk that used it has been computed; waiting
        //      for its copy-out as well keeps one rule for all buffers of the slot)
        if (sl.used) CK(cudaStreamWaitEvent(c->st_in, sl.ev_out, 0));
        mark(c->st_in);
        CK(cudaMemcpyAsync((packed || hostpack) ? sl.synd.p : sl.u8in.p, hostpack ? sl.h_synd.p : (const void *)(synd + (size_t)o * in_row),
                           (size_t)b * in_bus, cudaMemcpyHostToDevice, c->st_in));
        c->h2d_bytes += (size_t)b * in_bus;
        CK(cudaEventRecord(sl.ev_in, c->st_in));
        mark(c->st_in);
        // ---- compute
        CK(cudaStreamWaitEvent(c->st_comp, sl.ev_in, 0));
        if (!packed && !hostpack)
            if (int rc = qldpc_pack_bits_dev(sl.u8in.as<uint8_t>(), sl.synd.as<uint32_t>(), b, c->m, c->st_comp)) return rc;
        if (int rc = bposd_chunk(c, cfg, prior, b, sl.synd.as<uint32_t>(), osd_order, sl.hard.as<uint32_t>(), sl.conv.as<uint8_t>(),
                                 sl.iters.as<int32_t>(), nullptr, &sl.ctrl, &sl.gstate, &sl.llr, &sl.fail, &sl.redo, &sl.valid, &sl.inv, c->st_comp))
            return rc;
        if (!packed && !hostpack)
            if (int rc = qldpc_unpack_bits_dev(sl.hard.as<uint32_t>(), sl.u8out.as<uint8_t>(), b, c->n, c->st_comp)) return rc;
        CK(cudaEventRecord(sl.ev_comp, c->st_comp));
        mark(c->st_comp);
        // ---- copy out
        CK(cudaStreamWaitEvent(c->st_out, sl.ev_comp, 0));
        CK(cudaEventRecord(sl.ev_out0, c->st_out));
        CK(cudaMemcpyAsync(hostpack ? sl.h_hard.p : (void *)(corr + (size_t)o * out_row), (packed || hostpack) ? sl.hard.p : sl.u8out.p,
                           (size_t)b * out_bus, cudaMemcpyDeviceToHost, c->st_out));
        CK(cudaMemcpyAsync(conv + o, sl.conv.p, (size_t)b, cudaMemcpyDeviceToHost, c->st_out));
        if (iters) CK(cudaMemcpyAsync(iters + o, sl.iters.p, 4 * (size_t)b, cudaMemcpyDeviceToHost, c->st_out));
        c->d2h_bytes += (size_t)b * (out_bus + 1 + (iters ? 4 : 0));
        CK(cudaEventRecord(sl.ev_out, c->st_out));
        mark(c->st_out);
        sl.used = true;
        sl.host_mode = hostpack;
        sl.inflight = true;
        sl.cur_b = b;
        if (hostpack) { sl.pend = true; sl.pend_o = o; sl.pend_b = b; }
        return QLDPC_OK;
    };
    int rc_all = QLDPC_OK;
    // Tapered schedule for large batches: the pipeline fills with a quarter- and a half-sized chunk and drains with a half-
    // and a quarter-sized one, so that the first copy-in and the last copy-out (not hidden under compute) are short.
    const bool taper = !getenv("QLDPC_HOST_NO_TAPER") && B >= 4 * chunk;
    for (long long o = 0; o < B && rc_all == QLDPC_OK; ++i) {
        long long b = chunk;
        if (taper) {
            const long long left = B - o;
            if (o == 0) b = chunk / 4;
            else if (o == chunk / 4) b = chunk / 2;
            else if (left <= chunk / 4) b = left;
            else if (left <= 3 * chunk / 4) b = left - chunk / 4;
            else if (left < 7 * chunk / 4) b = left - 3 * chunk / 4;
        }
        b = std::min<long long>(b, B - o);
        rc_all = enqueue(o, b, c->slot[i % qldpc_code::NSLOT]);
        o += b;
    }
    const std::string msg = g_err;
    if (hp_mode != 0)                                         // expand the chunks still in flight, oldest first
        for (int a = 0; a < qldpc_code::NSLOT; ++a) {
            qldpc_code::Slot &sl = c->slot[(i + a) % qldpc_code::NSLOT];
            if (!sl.pend) continue;
            if (cudaEventSynchronize(sl.ev_out) == cudaSuccess && rc_all == QLDPC_OK) finish(sl);
            sl.pend = false;
        }
    for (cudaStream_t st : {c->st_in, c->st_comp, c->st_out})
        if (st && cudaStreamSynchronize(st) != cudaSuccess && rc_all == QLDPC_OK)
            rc_all = fail(QLDPC_ERR_CUDA, "qldpc_bposd_decode_host: stream synchronisation failed");
    if (trace) {
        for (size_t q = 0; q + 3 < tev.size(); q += 4) {
            float t[4];
            for (int x = 0; x < 4; ++x) cudaEventElapsedTime(&t[x], tev[0], tev[q + x]);
            fprintf(stderr, "[qldpc trace] chunk %zu: h2d %.2f - %.2f  compute done %.2f  d2h done %.2f ms\n", q / 4, t[0], t[1], t[2], t[3]);
        }
        for (cudaEvent_t e : tev) cudaEventDestroy(e);
    }
    if (rc_all != QLDPC_OK && !msg.empty()) g_err = msg;
    return rc_all;
}

extern "C" int qldpc_bposd_decode_host(qldpc_code *c, const qldpc_bp_config *cfg, const double *prior, int64_t B,
                                       const uint8_t *synd, int32_t osd_order, uint8_t *corr, uint8_t *conv, int32_t *iters)
{
    if (!c) return fail(QLDPC_ERR_ARG, "qldpc_bposd_decode_host: null code");
    return bposd_decode_host_impl(c, cfg, prior, B, synd, osd_order, corr, conv, iters, false);
}

extern "C" int qldpc_bposd_decode_host_packed(qldpc_code *c, const qldpc_bp_config *cfg, const double *prior, int64_t B,
                                              const uint32_t *synd, int32_t osd_order, uint32_t *corr, uint8_t *conv, int32_t *iters)
{
    if (!c) return fail(QLDPC_ERR_ARG, "qldpc_bposd_decode_host_packed: null code");
    return bposd_decode_host_impl(c, cfg, prior, B, synd, osd_order, corr, conv, iters, true);
}

extern "C" int qldpc_set_host_pack(qldpc_code *c, int32_t mode)
{
    if (!c) return fail(QLDPC_ERR_ARG, "qldpc_set_host_pack: null code");
    if (mode < -1 || mode > 2) return fail(QLDPC_ERR_ARG, "qldpc_set_host_pack: mode must be -1 (measure), 0 (device), 1 (host) or 2 (chunk by chunk)");
    if (mode < 0) c->host_pack = -1;             // decided again at the next call
    else host_pack_decide(c, mode);
    return QLDPC_OK;
}

extern "C" int qldpc_host_transfer_stats(qldpc_code *c, uint64_t *h2d_bytes, uint64_t *d2h_bytes, int32_t *host_pack,
                                         double *host_pack_rate, uint64_t *chunks_host, uint64_t *chunks_device)
{
    if (!c) return fail(QLDPC_ERR_ARG, "qldpc_host_transfer_stats: null code");
    if (h2d_bytes) *h2d_bytes = c->h2d_bytes;
    if (d2h_bytes) *d2h_bytes = c->d2h_bytes;
    if (host_pack) *host_pack = c->host_pack;
    if (host_pack_rate) *host_pack_rate = c->host_pack_rate;
    if (chunks_host) *chunks_host = c->host_chunks;
    if (chunks_device) *chunks_device = c->dev_chunks;
    return QLDPC_OK;
}

extern "C" int qldpc_check_host(qldpc_code *c, int64_t B, const uint8_t *err, const uint8_t *corr, const uint8_t *synd,
                                const uint8_t *conv, const int32_t *iters, int32_t distance, uint8_t *flags,
                                int32_t *weight, uint64_t *counters)
{
    if (!c || !err || !corr || !synd) return fail(QLDPC_ERR_ARG, "qldpc_check_host: null argument");
    cudaStream_t st = 0;
    CK(c->ws_cnt.reserve(8 * QLDPC_NUM_COUNTERS));
    CK(cudaMemsetAsync(c->ws_cnt.p, 0, 8 * QLDPC_NUM_COUNTERS, st));
    const long long chunk = 1ll << 20;
    for (long long o = 0; o < B; o += chunk) {
        const long long b = std::min<long long>(chunk, B - o);
        CK(c->ws_u8a.reserve((size_t)b * std::max(c->m, c->n)));
        CK(c->ws_synd.reserve(4 * (size_t)b * c->WM));
        CK(c->ws_hard.reserve(4 * (size_t)b * c->WN));
        CK(c->ws_err.reserve(4 * (size_t)b * c->WN));
        CK(c->ws_conv.reserve((size_t)b));
        CK(c->ws_iters.reserve(4 * (size_t)b));
        CK(c->ws_flags.reserve((size_t)b));
        CK(c->ws_weight.reserve(4 * (size_t)b));
        CK(cudaMemcpyAsync(c->ws_u8a.p, synd + (size_t)o * c->m, (size_t)b * c->m, cudaMemcpyHostToDevice, st));
        if (int rc = qldpc_pack_bits_dev(c->ws_u8a.as<uint8_t>(), c->ws_synd.as<uint32_t>(), b, c->m, st)) return rc;
        CK(cudaMemcpyAsync(c->ws_u8a.p, err + (size_t)o * c->n, (size_t)b * c->n, cudaMemcpyHostToDevice, st));
        if (int rc = qldpc_pack_bits_dev(c->ws_u8a.as<uint8_t>(), c->ws_err.as<uint32_t>(), b, c->n, st)) return rc;
        CK(cudaMemcpyAsync(c->ws_u8a.p, corr + (size_t)o * c->n, (size_t)b * c->n, cudaMemcpyHostToDevice, st));
        if (int rc = qldpc_pack_bits_dev(c->ws_u8a.as<uint8_t>(), c->ws_hard.as<uint32_t>(), b, c->n, st)) return rc;
        if (conv) CK(cudaMemcpyAsync(c->ws_conv.p, conv + o, (size_t)b, cudaMemcpyHostToDevice, st));
        if (iters) CK(cudaMemcpyAsync(c->ws_iters.p, iters + o, 4 * (size_t)b, cudaMemcpyHostToDevice, st));
        if (int rc = qldpc_check_dev(c, b, c->ws_err.as<uint32_t>(), c->ws_hard.as<uint32_t>(), c->ws_synd.as<uint32_t>(),
                                     conv ? c->ws_conv.as<uint8_t>() : nullptr, iters ? c->ws_iters.as<int32_t>() : nullptr,
                                     distance, c->ws_flags.as<uint8_t>(), c->ws_weight.as<int32_t>(), c->ws_cnt.as<uint64_t>(), st))
            return rc;
        if (flags) CK(cudaMemcpyAsync(flags + o, c->ws_flags.p, (size_t)b, cudaMemcpyDeviceToHost, st));
        if (weight) CK(cudaMemcpyAsync(weight + o, c->ws_weight.p, 4 * (size_t)b, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
    }
    if (counters) CK(cudaMemcpy(counters, c->ws_cnt.p, 8 * QLDPC_NUM_COUNTERS, cudaMemcpyDeviceToHost));
    return QLDPC_OK;
}

extern "C" int qldpc_syndrome_dev(qldpc_code *c, int64_t B, const uint32_t *err, uint32_t *synd, void *stream)
{
    if (!c || !err || !synd) return fail(QLDPC_ERR_ARG, "qldpc_syndrome_dev: null argument");
    if (B <= 0) return QLDPC_OK;
    syndrome_kernel<<<grid_for(B * c->WM, 256, c->num_sms), 256, 0, (cudaStream_t)stream>>>(c->d_row_ptr, c->d_col_idx, c->m, c->WM, c->WN,
                                                                                           B, err, synd);
    CK(cudaGetLastError());
    return QLDPC_OK;
}

extern "C" int qldpc_syndrome_host(qldpc_code *c, int64_t B, const uint8_t *err, uint8_t *synd)
{
    if (!c || !err || !synd) return fail(QLDPC_ERR_ARG, "qldpc_syndrome_host: null argument");
    cudaStream_t st = 0;
    const long long chunk = 1ll << 20;
    for (long long o = 0; o < B; o += chunk) {
        const long long b = std::min<long long>(chunk, B - o);
        CK(c->ws_u8a.reserve((size_t)b * std::max(c->m, c->n)));
        CK(c->ws_synd.reserve(4 * (size_t)b * c->WM));
        CK(c->ws_err.reserve(4 * (size_t)b * c->WN));
        CK(cudaMemcpyAsync(c->ws_u8a.p, err + (size_t)o * c->n, (size_t)b * c->n, cudaMemcpyHostToDevice, st));
        if (int rc = qldpc_pack_bits_dev(c->ws_u8a.as<uint8_t>(), c->ws_err.as<uint32_t>(), b, c->n, st)) return rc;
        if (int rc = qldpc_syndrome_dev(c, b, c->ws_err.as<uint32_t>(), c->ws_synd.as<uint32_t>(), st)) return rc;
        if (int rc = qldpc_unpack_bits_dev(c->ws_synd.as<uint32_t>(), c->ws_u8a.as<uint8_t>(), b, c->m, st)) return rc;
        CK(cudaMemcpyAsync(synd + (size_t)o * c->m, c->ws_u8a.p, (size_t)b * c->m, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
    }
    return QLDPC_OK;
}

extern "C" int qldpc_sample_host(qldpc_code *c, double p, uint64_t seed, uint64_t first_shot, int32_t draws, int64_t B,
                                 uint8_t *err, uint8_t *synd)
{
    return qldpc_sample_noisy_host(c, p, 0.0, seed, first_shot, draws, B, err, synd);
}

extern "C" int qldpc_sample_noisy_host(qldpc_code *c, double p, double q_meas, uint64_t seed, uint64_t first_shot, int32_t draws,
                                       int64_t B, uint8_t *err, uint8_t *synd)
{
    if (!c || !err || !synd) return fail(QLDPC_ERR_ARG, "qldpc_sample_host: null argument");
    cudaStream_t st = 0;
    const long long chunk = 1ll << 20;
    for (long long o = 0; o < B; o += chunk) {
        const long long b = std::min<long long>(chunk, B - o);
        CK(c->ws_u8a.reserve((size_t)b * std::max(c->m, c->n)));
        CK(c->ws_synd.reserve(4 * (size_t)b * c->WM));
        CK(c->ws_err.reserve(4 * (size_t)b * c->WN));
        if (int rc = qldpc_sample_dev(c, p, seed, first_shot + (uint64_t)o, draws, b, c->ws_err.as<uint32_t>(),
                                      c->ws_synd.as<uint32_t>(), st))
            return rc;
        if (int rc = qldpc_measurement_noise_dev(c, q_meas, seed, first_shot + (uint64_t)o, b, c->ws_synd.as<uint32_t>(), st)) return rc;
        if (int rc = qldpc_unpack_bits_dev(c->ws_err.as<uint32_t>(), c->ws_u8a.as<uint8_t>(), b, c->n, st)) return rc;
        CK(cudaMemcpyAsync(err + (size_t)o * c->n, c->ws_u8a.p, (size_t)b * c->n, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (int rc = qldpc_unpack_bits_dev(c->ws_synd.as<uint32_t>(), c->ws_u8a.as<uint8_t>(), b, c->m, st)) return rc;
        CK(cudaMemcpyAsync(synd + (size_t)o * c->m, c->ws_u8a.p, (size_t)b * c->m, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
    }
    return QLDPC_OK;
}

// BP on device-sampled shots with the posterior LLRs histogrammed on the device instead of returned:
// hist [3][nbins] uint64 (true bit 0 / true bit 1 / BP-failed shots), ADDED into the host array; counters as qldpc_mc_sweep.
extern "C" int qldpc_bp_llr_histogram(qldpc_code *c, const qldpc_bp_config *cfg, const double *prior, double p, uint64_t seed,
                                      uint64_t first_shot, int64_t nshots, int32_t draws, double lo, double hi, int32_t nbins,
                                      uint64_t *hist, uint64_t *bp_failed)
{
    if (!c || !hist || nbins <= 0 || nbins > 4096 || !(hi > lo)) return fail(QLDPC_ERR_ARG, "qldpc_bp_llr_histogram: bad argument");
    if (int rc = check_cfg(cfg)) return rc;
    cudaStream_t st = 0;
    const int tsize = cfg->precision == 64 ? 8 : 4;
    CK(c->ws_cnt.reserve(8 * (size_t)3 * nbins + 16));
    CK(cudaMemsetAsync(c->ws_cnt.p, 0, 8 * (size_t)3 * nbins + 16, st));
    unsigned long long *d_hist = c->ws_cnt.as<unsigned long long>();
    unsigned long long *d_fail = d_hist + 3 * nbins;
    const long long chunk = 1ll << 20;
    for (long long o = 0; o < nshots; o += chunk) {
        const long long b = std::min<long long>(chunk, nshots - o);
        CK(c->ws_synd.reserve(4 * (size_t)b * c->WM));
        CK(c->ws_hard.reserve(4 * (size_t)b * c->WN));
        CK(c->ws_err.reserve(4 * (size_t)b * c->WN));
        CK(c->ws_conv.reserve((size_t)b));
        CK(c->ws_iters.reserve(4 * (size_t)b));
        CK(c->ws_llr.reserve((size_t)b * c->n * tsize));
        if (int rc = qldpc_sample_dev(c, p, seed, first_shot + (uint64_t)o, draws, b, c->ws_err.as<uint32_t>(), c->ws_synd.as<uint32_t>(), st))
            return rc;
        if (int rc = qldpc_bp_decode_dev(c, cfg, prior, b, c->ws_synd.as<uint32_t>(), c->ws_hard.as<uint32_t>(), c->ws_conv.as<uint8_t>(),
                                         c->ws_iters.as<int32_t>(), c->ws_llr.p, QLDPC_LLR_ALL, nullptr, nullptr, nullptr, st))
            return rc;
        const int grid = grid_for(b * (long long)c->n, 256, c->num_sms);
        const size_t smem = 4 * (size_t)3 * nbins;
        if (tsize == 8)
            llr_hist_kernel<double><<<grid, 256, smem, st>>>(c->ws_llr.as<double>(), c->ws_err.as<uint32_t>(), c->ws_conv.as<uint8_t>(), b,
                                                             c->n, c->WN, lo, hi, nbins, d_hist, d_fail);
        else
            llr_hist_kernel<float><<<grid, 256, smem, st>>>(c->ws_llr.as<float>(), c->ws_err.as<uint32_t>(), c->ws_conv.as<uint8_t>(), b,
                                                            c->n, c->WN, lo, hi, nbins, d_hist, d_fail);
        CK(cudaGetLastError());
    }
    std::vector<uint64_t> h((size_t)3 * nbins + 1);
    CK(cudaMemcpy(h.data(), d_hist, 8 * h.size(), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < (size_t)3 * nbins; ++i) hist[i] += h[i];
    if (bp_failed) *bp_failed += h[(size_t)3 * nbins];
    return QLDPC_OK;
}

// rework/Alvarado.py:10-66 on the device: the four edge counts behind the alpha estimator (see alpha_counts_kernel)
extern "C" int qldpc_alpha_counts(qldpc_code *c, int64_t B, const uint8_t *err_host, double p, uint64_t seed, uint64_t first_shot,
                                  uint64_t *counts)
{
    if (!c || !counts) return fail(QLDPC_ERR_ARG, "qldpc_alpha_counts: null argument");
    if (!err_host && !(p >= 0.0 && p < 1.0)) return fail(QLDPC_ERR_ARG, "qldpc_alpha_counts: need 0 <= p < 1");
    cudaStream_t st = 0;
    CK(c->ws_cnt.reserve(8 * QLDPC_NUM_COUNTERS));
    CK(cudaMemsetAsync(c->ws_cnt.p, 0, 8 * 4, st));
    const long long chunk = 1ll << 20;
    for (long long o = 0; o < B; o += chunk) {
        const long long b = std::min<long long>(chunk, B - o);
        CK(c->ws_err.reserve(4 * (size_t)b * c->WN));
        if (err_host) {
            CK(c->ws_u8a.reserve((size_t)b * std::max(c->m, c->n)));
            CK(cudaMemcpyAsync(c->ws_u8a.p, err_host + (size_t)o * c->n, (size_t)b * c->n, cudaMemcpyHostToDevice, st));
            if (int rc = qldpc_pack_bits_dev(c->ws_u8a.as<uint8_t>(), c->ws_err.as<uint32_t>(), b, c->n, st)) return rc;
        } else {
            CK(c->ws_synd.reserve(4 * (size_t)b * c->WM));
            if (int rc = qldpc_sample_dev(c, p, seed, first_shot + (uint64_t)o, 1, b, c->ws_err.as<uint32_t>(), c->ws_synd.as<uint32_t>(), st))
                return rc;
        }
        alpha_counts_kernel<<<grid_for(b, 128, c->num_sms), 128, 0, st>>>(c->d_Hrows, c->d_row_ptr, c->m, c->WN, b, c->ws_err.as<uint32_t>(),
                                                                         c->ws_cnt.as<unsigned long long>());
        CK(cudaGetLastError());
        if (err_host) CK(cudaStreamSynchronize(st));          // (the staging buffer is reused by the next chunk)
    }
    uint64_t h[4];
    CK(cudaMemcpyAsync(h, c->ws_cnt.p, sizeof(h), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    for (int i = 0; i < 4; ++i) counts[i] += h[i];
    return QLDPC_OK;
}

extern "C" int qldpc_mc_sweep(qldpc_code *c, const qldpc_bp_config *cfg, const double *prior, double p, uint64_t seed,
                              uint64_t first_shot, int64_t nshots, int32_t draws, int32_t osd_order, int32_t distance,
                              uint64_t *counters)
{
    return qldpc_mc_sweep_noisy(c, cfg, prior, p, 0.0, seed, first_shot, nshots, draws, osd_order, distance, counters);
}

extern "C" int qldpc_mc_sweep_noisy(qldpc_code *c, const qldpc_bp_config *cfg, const double *prior, double p, double q_meas,
                                    uint64_t seed, uint64_t first_shot, int64_t nshots, int32_t draws, int32_t osd_order,
                                    int32_t distance, uint64_t *counters)
{
    if (!c || !counters) return fail(QLDPC_ERR_ARG, "qldpc_mc_sweep: null argument");
    if (int rc = check_cfg(cfg)) return rc;
    cudaStream_t st = 0;
    CK(c->ws_cnt.reserve(8 * QLDPC_NUM_COUNTERS));
    CK(cudaMemsetAsync(c->ws_cnt.p, 0, 8 * QLDPC_NUM_COUNTERS, st));
    const long long chunk = CHUNK;
    for (long long o = 0; o < nshots; o += chunk) {
        const long long b = std::min<long long>(chunk, nshots - o);
        CK(c->ws_synd.reserve(4 * (size_t)b * c->WM));
        CK(c->ws_hard.reserve(4 * (size_t)b * c->WN));
        CK(c->ws_err.reserve(4 * (size_t)b * c->WN));
        CK(c->ws_conv.reserve((size_t)b));
        CK(c->ws_iters.reserve(4 * (size_t)b));
        if (int rc = qldpc_sample_dev(c, p, seed, first_shot + (uint64_t)o, draws, b, c->ws_err.as<uint32_t>(),
                                      c->ws_synd.as<uint32_t>(), st))
            return rc;
        if (int rc = qldpc_measurement_noise_dev(c, q_meas, seed, first_shot + (uint64_t)o, b, c->ws_synd.as<uint32_t>(), st)) return rc;
        if (int rc = qldpc_bposd_decode_dev(c, cfg, prior, b, c->ws_synd.as<uint32_t>(), osd_order, c->ws_hard.as<uint32_t>(),
                                            c->ws_conv.as<uint8_t>(), c->ws_iters.as<int32_t>(), nullptr, st))
            return rc;
        if (int rc = qldpc_check_dev(c, b, c->ws_err.as<uint32_t>(), c->ws_hard.as<uint32_t>(), c->ws_synd.as<uint32_t>(),
                                     c->ws_conv.as<uint8_t>(), c->ws_iters.as<int32_t>(), distance, nullptr, nullptr,
                                     c->ws_cnt.as<uint64_t>(), st))
            return rc;
    }
    uint64_t h[QLDPC_NUM_COUNTERS];
    CK(cudaMemcpyAsync(h, c->ws_cnt.p, sizeof(h), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    for (int i = 0; i < QLDPC_NUM_COUNTERS; ++i) counters[i] += h[i];
    return QLDPC_OK;
}
