// Batched belief propagation for quantum LDPC codes on sm_100a -- one THREAD per shot.
//
// Replaces (reference paths relative to michelebanfi/qLDPC):
//   rework/decoding.py:5-75      performMinSum_Symmetric            (VAR_MIN_SUM)
//   decoding/beliefPropagation.py:88-144 performBeliefPropagationFast (VAR_SUM_PRODUCT, sym = 0)
//   decoding/beliefPropagation.py:6-85   performBeliefPropagation    (same maths, sequential sums)
//   rework/decoding.py:131-191   performBeliefPropagation_Symmetric  (VAR_SUM_PRODUCT, sym = 1)
//   decoding/beliefPropagationGPU.py:81-178 performBeliefPropagationBatch (the batch axis)
//
// Design.  The Tanner graph is the same for every shot, so with one
// thread per shot every graph index is warp-uniform (broadcast table reads) and the per-shot
// message state is laid out [edge][thread]: lane i touches word i of a 128-byte row, i.e.
// conflict-free in shared memory and perfectly coalesced in HBM.  No shuffles, no barriers in
// the decode loop.  The kernel is persistent: a thread whose shot matched its syndrome (or hit
// max_iter) writes its result and pulls the next shot id from a global cursor, so slow shots
// never hold up a tile.
//
//   STATE_SMEM = true : state + graph tables in shared memory (code-capacity H, <= ~7 KB/shot)
//   STATE_SMEM = false: state staged in HBM, tables through L1 (space-time / DEM sized H)
//
// Per-edge state is ONE word: the variable-to-check message Q (or tanh(Q/2) for plain
// sum-product).  The check pass leaves a 2-word summary per check (signed min1, min2 -- or the
// signed row product); the variable pass rebuilds each check-to-variable message from the
// summary and the still-unmodified Q, which is also the Q_old the reference damps against.
//
// Exactness: the float64 instantiation performs the reference's float64 operations in the
// reference's order (no FMA contraction: __dmul_rn/__dadd_rn; posterior additions follow the
// per-variable order table computed on the host from NumPy's summation scheme),
// so min-sum is bit-identical to the reference, including posterior LLRs and exit iteration.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>
#include <type_traits>

namespace qldpc {

enum { VAR_MIN_SUM = 0, VAR_SUM_PRODUCT = 1 };
enum { LLR_NONE = 0, LLR_FAILED = 1, LLR_ALL = 2 };

struct BPGraphDev {
    int m, n, E;
    int WM, WN;                 // 32-bit words per packed syndrome / packed error
    int uniform_row_w;          // row weight if all rows have the same weight (<= 8), else 0
    int max_col_w;
    int two_tables;             // vtab1 differs from vtab0
    const int32_t *row_ptr;     // [m+1]
    const int32_t *col_idx;     // [E]    ascending column inside a row
    const int32_t *var_ptr;     // [n+1]
    const uint32_t *vtab0;      // [2E]   (edge, check) pairs of each variable in ADD ORDER, iteration 0
    const uint32_t *vtab1;      // [2E]   same for iterations >= 1
    const uint32_t *colmask;    // [n][WM] packed columns of H
};

struct BPParams {
    BPGraphDev g;
    long long B;                // shots in this launch
    const uint32_t *synd;       // [B][WM] packed syndromes
    const void *prior;          // [n] T
    int max_iter;
    int sym;                    // sum-product: apply alpha / damping / clip (decoding.py:131)
    double alpha, damping, one_minus_damping, clip;
    int prior_uniform;          // all priors equal (warp kernel: the message state starts as one constant)
    int zero_ok;                // all priors > 0 (also as float32) and alpha >= 0: a shot with an all-zero syndrome converges at iteration 0
                                //   with the all-zero hard decision (every message and posterior is positive) -- the warp kernel retires it unseen
    double qpad;                // max(clip, largest prior): start value of padding edge slots (bp_warp / bp_cta kernels)
    uint32_t *hard;             // [B][WN] packed hard decisions (out)
    uint8_t *conv;              // [B] (out)
    int32_t *iters;             // [B] 0-based exit iteration (out, may be null)
    void *llr;                  // [B][n] T posterior (out, may be null)
    int llr_mode;
    unsigned long long *cursor; // global shot cursor (zeroed before launch)
    int32_t *fail_idx;          // compacted list of BP-failed shots (may be null)
    unsigned int *fail_count;
    unsigned long long *iter_total; // sum over shots of executed iterations (may be null)
    void *gstate;               // STATE_SMEM = false: [(2E + 2m) T + WN + WM words][total threads]
    void *r_dump;               // [B][E] T or null: check-to-variable messages of iteration `dump_iter` (CSR edge order),
    int dump_iter;              //   the alpha_estimation=True return of the reference (decoding.py:58-59,168-169)
};

template <typename T> struct Num;
template <> struct Num<float> {
    typedef uint32_t bits_t;
    static constexpr bits_t SIGN = 0x80000000u;
    __device__ static __forceinline__ bits_t bits(float x) { return __float_as_uint(x); }
    __device__ static __forceinline__ float from_bits(bits_t b) { return __uint_as_float(b); }
    __device__ static __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    __device__ static __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    __device__ static __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    __device__ static __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
    __device__ static __forceinline__ float inf() { return CUDART_INF_F; }
    __device__ static __forceinline__ float tanh_(float x) { return tanhf(x); }
    __device__ static __forceinline__ float atanh_(float x) { return atanhf(x); }
};
template <> struct Num<double> {
    typedef unsigned long long bits_t;
    static constexpr bits_t SIGN = 0x8000000000000000ull;
    __device__ static __forceinline__ bits_t bits(double x) { return (bits_t)__double_as_longlong(x); }
    __device__ static __forceinline__ double from_bits(bits_t b) { return __longlong_as_double((long long)b); }
    __device__ static __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    __device__ static __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    __device__ static __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    __device__ static __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
    __device__ static __forceinline__ double inf() { return CUDART_INF; }
    __device__ static __forceinline__ double tanh_(double x) { return tanh(x); }
    __device__ static __forceinline__ double atanh_(double x) { return atanh(x); }
};

// Damping  Q = d * Q_new + (1 - d) * Q_old  (rework/decoding.py:65).  float64 keeps the reference's three
// roundings (NumPy evaluates the two products and the sum as separate ufuncs); the float32 production
// kernels fuse the second product into the add (one rounding fewer, one instruction fewer).
__device__ __forceinline__ double bp_damp(double d, double qn, double omd, double qo) { return __dadd_rn(__dmul_rn(d, qn), __dmul_rn(omd, qo)); }
__device__ __forceinline__ float bp_damp(float d, float qn, float omd, float qo) { return __fmaf_rn(d, qn, __fmul_rn(omd, qo)); }
// sign(0) is + in the reference (decoding.py:30): the float64 kernel turns a -0.0 message into +0.0 so that the
// sign-bit XOR of the check pass is exact.  -0.0 can only arise from an underflow in float32; not canonicalised there.
__device__ __forceinline__ double bp_canon(double q) { return __dadd_rn(q, 0.0); }
__device__ __forceinline__ float bp_canon(float q) { return q; }

// Check-to-variable message of sum-product from the leave-one-out product x:  2 * atanh(clip(x, +-0.9999999))
// (beliefPropagation.py:125-126).  0.9999999 is not a float32 number (it rounds to 1 - 2^-23, which would move the
// saturation value from 16.81 to 16.64), so the float32 kernels clip x to the largest float below 1 and clamp the
// result to the reference's saturation value 2 * atanh(0.9999999) instead.
__device__ __forceinline__ double bp_sp_r(double x)
{
    x = fmin(fmax(x, -0.9999999), 0.9999999);
    return __dmul_rn(2.0, atanh(x));
}
__device__ __forceinline__ float bp_sp_r(float x)
{
    x = fminf(fmaxf(x, -0x1.fffffep-1f), 0x1.fffffep-1f);
    const float r = __fmul_rn(2.f, atanhf(x));
    return fminf(fmaxf(r, -16.811242831518264f), 16.811242831518264f);
}

// Message-state accesses.  (Measured on B200: evict-first hints -- ld/st.global.cs -- on the HBM-staged state make the
// kernel 10 % slower, because the per-check summaries written by the check pass are re-read from L2 by the variable
// pass; default caching is kept.)
template <bool SMEM, typename U> __device__ __forceinline__ U ld_state(const U *p) { return *p; }
template <bool SMEM, typename U> __device__ __forceinline__ void st_state(U *p, U v) { *p = v; }

// Shared-memory footprint, shared with the host (capi.cu) so both agree on the carve-up.
struct BPSmemLayout {
    size_t off_rowptr, off_varptr, off_vtab0, off_vtab1, off_colmask, off_prior, off_state;
    size_t per_shot;            // bytes of state per shot slot
    size_t tables;              // bytes before the state
};
__host__ __device__ inline BPSmemLayout bp_smem_layout(const BPGraphDev &g, int tsize, int variant)
{
    BPSmemLayout L;
    size_t o = 0;
    L.off_rowptr = o;  o += 4 * (size_t)(g.m + 1);
    L.off_varptr = o;  o += 4 * (size_t)(g.n + 1);
    o = (o + 7) & ~(size_t)7;
    L.off_vtab0 = o;   o += 8 * (size_t)g.E;
    L.off_vtab1 = g.two_tables ? o : L.off_vtab0;
    if (g.two_tables) o += 8 * (size_t)g.E;
    L.off_colmask = o; o += 4 * (size_t)g.n * g.WM;
    o = (o + 7) & ~(size_t)7;
    L.off_prior = o;   o += (size_t)tsize * g.n;
    o = (o + 15) & ~(size_t)15;
    L.off_state = o;
    L.tables = o;
    L.per_shot = (size_t)tsize * (g.E + (size_t)g.m * (variant == VAR_MIN_SUM ? 2 : 1)) + 4 * (size_t)g.WN;
    return L;
}

// ------------------------------------------------------------------------------------------
// WMS > 0: syndrome words (and the running syndrome of the hard decision) live in registers,
//          the hard-decision syndrome is accumulated from packed columns of H in the variable pass.
// WMS == 0: (HBM-staged) syndrome words live in the state, the hard-decision syndrome is
//          evaluated check by check after the variable pass.
// ------------------------------------------------------------------------------------------
template <typename T, int VAR, int WMS, bool STATE_SMEM>
__global__ void __launch_bounds__(STATE_SMEM ? 256 : 128, STATE_SMEM ? 1 : 8)
bp_decode_kernel(const BPParams P)
{
    typedef Num<T> N;
    typedef typename N::bits_t bits_t;
    typedef typename std::conditional<STATE_SMEM, int, size_t>::type idx_t;   // word offsets inside the state
    const BPGraphDev &g = P.g;
    const int m = g.m, n = g.n, E = g.E, WN = g.WN;
    const int WM = (WMS > 0) ? WMS : g.WM;
    const int lane = threadIdx.x & 31;

    extern __shared__ __align__(16) unsigned char smem[];

    // ---- tables -------------------------------------------------------------------------
    const int32_t *row_ptr;
    const int32_t *var_ptr;
    const uint2 *vtab0, *vtab1;
    const uint32_t *colmask;
    const T *prior;
    // ---- state (stride S words between consecutive edges of one shot) ---------------------
    // (__restrict__: the arrays are disjoint, so loads of the next check may be hoisted above the summary stores)
    // HBM-staged mode ping-pongs between two message arrays (Q is only read, QW only written during an iteration), so that
    // the loads of the next variable can be issued before the stores of the current one retire: memory-level parallelism
    T *Q, *QW;
    T *__restrict__ M1, *__restrict__ M2;
    uint32_t *__restrict__ HW, *__restrict__ SY = nullptr;
    int S;

    if (STATE_SMEM) {
        const BPSmemLayout L = bp_smem_layout(g, (int)sizeof(T), VAR);
        int32_t *s_rowptr = reinterpret_cast<int32_t *>(smem + L.off_rowptr);
        int32_t *s_varptr = reinterpret_cast<int32_t *>(smem + L.off_varptr);
        uint2 *s_vtab0 = reinterpret_cast<uint2 *>(smem + L.off_vtab0);
        uint2 *s_vtab1 = reinterpret_cast<uint2 *>(smem + L.off_vtab1);
        uint32_t *s_colmask = reinterpret_cast<uint32_t *>(smem + L.off_colmask);
        T *s_prior = reinterpret_cast<T *>(smem + L.off_prior);
        for (int i = threadIdx.x; i <= m; i += blockDim.x) s_rowptr[i] = g.row_ptr[i];
        for (int i = threadIdx.x; i <= n; i += blockDim.x) s_varptr[i] = g.var_ptr[i];
        for (int i = threadIdx.x; i < E; i += blockDim.x) {
            s_vtab0[i] = reinterpret_cast<const uint2 *>(g.vtab0)[i];
            if (g.two_tables) s_vtab1[i] = reinterpret_cast<const uint2 *>(g.vtab1)[i];
        }
        for (int i = threadIdx.x; i < n * WM; i += blockDim.x) s_colmask[i] = g.colmask[i];
        for (int i = threadIdx.x; i < n; i += blockDim.x) s_prior[i] = reinterpret_cast<const T *>(P.prior)[i];
        __syncthreads();
        row_ptr = s_rowptr; var_ptr = s_varptr; vtab0 = s_vtab0; vtab1 = s_vtab1;
        colmask = s_colmask; prior = s_prior;
        S = blockDim.x;
        T *st = reinterpret_cast<T *>(smem + L.off_state);
        Q = st + threadIdx.x;
        QW = Q;
        M1 = Q + (idx_t)E * S;
        M2 = M1 + (idx_t)m * S;
        HW = reinterpret_cast<uint32_t *>(st + (idx_t)(E + (VAR == VAR_MIN_SUM ? 2 : 1) * m) * S) + threadIdx.x;
    } else {
        row_ptr = g.row_ptr; var_ptr = g.var_ptr;
        vtab0 = reinterpret_cast<const uint2 *>(g.vtab0);
        vtab1 = reinterpret_cast<const uint2 *>(g.vtab1);
        colmask = g.colmask;
        prior = reinterpret_cast<const T *>(P.prior);
        S = gridDim.x * blockDim.x;
        const size_t gt = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
        T *st = reinterpret_cast<T *>(P.gstate);
        Q = st + gt;
        QW = Q + (idx_t)E * S;
        M1 = QW + (idx_t)E * S;
        M2 = M1 + (idx_t)m * S;
        HW = reinterpret_cast<uint32_t *>(st + (idx_t)(2 * E + 2 * m) * S) + gt;
        SY = HW + (idx_t)WN * S;
    }

    const T alpha = (T)P.alpha, damp = (T)P.damping, omd = (T)P.one_minus_damping, clipv = (T)P.clip;
    const int max_iter = P.max_iter;
    const bool slot_is_tanh = (VAR == VAR_SUM_PRODUCT) && !P.sym;
    const int rw = g.uniform_row_w;

    constexpr int WREG = (WMS > 0) ? WMS : 1;
    uint32_t synd[WREG], acc[WREG];
    // WMS == 0: running syndrome of the hard decision as a thread-private bit array (dynamically indexed, i.e. in local
    // memory), toggled only for the rare variables decided as 1 -- instead of re-reading the hard-decision words check by check
    constexpr int PARW = (WMS > 0) ? 1 : 64;
    uint32_t par[PARW];
    const bool use_par = (WMS == 0) && (WM <= PARW);
    long long shot = -1;
    int iter = 0;
    bool active = false, exhausted = false;
    unsigned long long iter_sum = 0;

    while (true) {
        // ---- refill: finished lanes claim the next shot ids (one atomic per warp) ----------
        const unsigned need = __ballot_sync(0xffffffffu, !active && !exhausted);
        if (need) {                                  // warp-uniform
            const int leader = __ffs(need) - 1;
            unsigned long long base = 0;
            if (lane == leader) base = atomicAdd(P.cursor, (unsigned long long)__popc(need));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (!active && !exhausted) {
                shot = (long long)(base + __popc(need & ((1u << lane) - 1)));
                if (shot < P.B) {
                    const uint32_t *sp = P.synd + (size_t)shot * WM;
                    if (WMS > 0) {
#pragma unroll
                        for (int w = 0; w < WREG; ++w) synd[w] = sp[w];
                    } else {
                        for (int w = 0; w < WM; ++w) SY[(idx_t)w * S] = sp[w];
                    }
                    // Q = where(mask, prior, 0)   (beliefPropagation.py:107 / decoding.py:21)
                    for (int v = 0; v < n; ++v) {
                        T pv = N::add(prior[v], (T)0);          // -0.0 -> +0.0 (sign(0) is + in the reference)
                        if (slot_is_tanh) pv = N::tanh_(N::mul(pv, (T)0.5));
                        for (int a = var_ptr[v]; a < var_ptr[v + 1]; ++a) Q[(idx_t)vtab1[a].x * S] = pv;
                    }
                    iter = 0;
                    active = true;
                } else {
                    exhausted = true;
                }
            }
        }
        if (!__any_sync(0xffffffffu, active)) break;
        if (active) {

        // ================= horizontal step: per-check summaries ============================
        for (int w = 0; w * 32 < m; ++w) {
            uint32_t sw;
            if (WMS > 0) {
                sw = 0;
#pragma unroll
                for (int k = 0; k < WREG; ++k) if (k == w) sw = synd[k];
            } else {
                sw = SY[(idx_t)w * S];
            }
            const int cend = min(32, m - 32 * w);
#pragma unroll 2
            for (int b = 0; b < cend; ++b) {
                const int c = 32 * w + b;
                const bits_t sbit = ((sw >> b) & 1u) ? N::SIGN : (bits_t)0;
                if (VAR == VAR_MIN_SUM) {
                    // decoding.py:28-53: sign product, min1, min2 (|Q| == min1 -> min2, compares values)
                    T min1 = N::inf(), min2 = N::inf();
                    bits_t sg = sbit;
                    if (rw == 6) {
                        const T *q = Q + (idx_t)(6 * c) * S;
                        T x[6];
#pragma unroll
                        for (int k = 0; k < 6; ++k) x[k] = ld_state<STATE_SMEM>(q + (idx_t)k * S);
#pragma unroll
                        for (int k = 0; k < 6; ++k) {
                            sg ^= N::bits(x[k]);
                            const T a = fabs(x[k]);
                            const T t = fmax(min1, a);
                            min1 = fmin(min1, a);
                            min2 = fmin(min2, t);
                        }
                    } else {
                        // up to 8 messages are loaded up front (memory-level parallelism matters when the state is
                        // staged in HBM); +inf is neutral for the minima and for the sign parity
                        const int e0 = row_ptr[c], deg = row_ptr[c + 1] - e0;
                        T x[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k) x[k] = (k < deg) ? ld_state<STATE_SMEM>(Q + (idx_t)(e0 + k) * S) : N::inf();
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            sg ^= N::bits(x[k]);
                            const T a = fabs(x[k]);
                            const T t = fmax(min1, a);
                            min1 = fmin(min1, a);
                            min2 = fmin(min2, t);
                        }
                        for (int e = e0 + 8; e < e0 + deg; ++e) {
                            const T xx = Q[(idx_t)e * S];
                            sg ^= N::bits(xx);
                            const T a = fabs(xx);
                            const T t = fmax(min1, a);
                            min1 = fmin(min1, a);
                            min2 = fmin(min2, t);
                        }
                    }
                    st_state<STATE_SMEM>(M1 + (idx_t)c * S, N::from_bits(N::bits(min1) | (sg & N::SIGN))); // signed min1
                    st_state<STATE_SMEM>(M2 + (idx_t)c * S, min2);
                } else {
                    // beliefPropagation.py:114-118: row product of tanh(Q/2), ascending column order
                    T prod = (T)1;
                    const int e1 = row_ptr[c + 1];
                    for (int e = row_ptr[c]; e < e1; ++e) {
                        T t = Q[(idx_t)e * S];
                        if (!slot_is_tanh) t = N::tanh_(N::mul(t, (T)0.5));
                        prod = N::mul(prod, t);
                    }
                    M1[(idx_t)c * S] = N::from_bits(N::bits(prod) ^ sbit);           // * (1 - 2 s)
                }
            }
        }

        // ================= vertical step ===================================================
        const uint2 *vt = (iter == 0) ? vtab0 : vtab1;
        const bool last = (iter == max_iter - 1);
        // read side / write side of this iteration: distinct arrays (hence __restrict__) when staged in HBM; the same array
        // in shared memory, where each edge is read before it is written, by the same thread
        typedef typename std::conditional<STATE_SMEM, const T *, const T *__restrict__>::type qr_t;
        typedef typename std::conditional<STATE_SMEM, T *, T *__restrict__>::type qw_t;
        qr_t qr = Q;
        qw_t qw = QW;
        const bool wr_llr = (P.llr != nullptr) && (P.llr_mode == LLR_ALL || (P.llr_mode == LLR_FAILED && last));
        T *llr_out = wr_llr ? reinterpret_cast<T *>(P.llr) + (size_t)shot * n : nullptr;
        const bool dumping = (P.r_dump != nullptr) && (iter == P.dump_iter);
        T *rdump = dumping ? reinterpret_cast<T *>(P.r_dump) + (size_t)shot * E : nullptr;
        if (WMS > 0) {
#pragma unroll
            for (int k = 0; k < WREG; ++k) acc[k] = 0;
        } else if (use_par) {
            for (int w = 0; w < WM; ++w) par[w] = 0;
        }
        for (int wv = 0; wv * 32 < n; ++wv) {
            uint32_t hw = 0;
            const int vend = min(32, n - 32 * wv);
#pragma unroll 2
            for (int b = 0; b < vend; ++b) {
                const int v = 32 * wv + b;
                const int a0 = var_ptr[v];
                const int deg = var_ptr[v + 1] - a0;
                // The first KB = 3 edges are handled branch-free: table entries and messages are loaded for a clamped edge
                // index (a valid address even when deg < 3) and masked afterwards, so that all loads of a variable are in
                // flight together -- essential when the state is staged in HBM.  Further edges go through the tail loop.
                constexpr int KB = 3;
                T r[KB], qo[KB];
                uint32_t eo[KB];
                T sum = (T)0;
                {
                    uint2 ecs[KB];
                    T qs[KB], s1s[KB], s2s[KB];
#pragma unroll
                    for (int k = 0; k < KB; ++k) ecs[k] = vt[min(a0 + min(k, max(deg - 1, 0)), E - 1)];
#pragma unroll
                    for (int k = 0; k < KB; ++k) {
                        qs[k] = ld_state<STATE_SMEM>(qr + (idx_t)ecs[k].x * S);
                        s1s[k] = ld_state<STATE_SMEM>(M1 + (idx_t)ecs[k].y * S);
                        s2s[k] = (VAR == VAR_MIN_SUM) ? ld_state<STATE_SMEM>(M2 + (idx_t)ecs[k].y * S) : (T)0;
                    }
#pragma unroll
                    for (int k = 0; k < KB; ++k) {
                        const T q = qs[k], s1 = s1s[k];
                        T rr;
                        if (VAR == VAR_MIN_SUM) {
                            const T a1 = fabs(s1);
                            const T mag = (fabs(q) == a1) ? s2s[k] : a1;                 // decoding.py:51-53
                            // R = alpha * syndrome_sign * r_signs * mag                   (decoding.py:55)
                            rr = N::from_bits(N::bits(N::mul(alpha, mag)) ^ ((N::bits(s1) ^ N::bits(q)) & N::SIGN));
                            if (dumping && k < deg) rdump[ecs[k].x] = N::div(rr, alpha);  // R_new / alpha (decoding.py:59)
                        } else {
                            T t = slot_is_tanh ? q : N::tanh_(N::mul(q, (T)0.5));
                            const T ts = (fabs(t) < (T)1e-15) ? (T)1e-15 : t;             // beliefPropagation.py:122
                            rr = bp_sp_r(N::div(s1, ts));                                // :125-126
                            if (dumping && k < deg) rdump[ecs[k].x] = rr;                 // R before scaling (decoding.py:169)
                            if (P.sym) rr = N::mul(rr, alpha);                           // decoding.py:171
                        }
                        if (k == 0) sum = (deg > 0) ? rr : (T)0;
                        else if (k < deg) sum = N::add(sum, rr);
                        r[k] = rr; qo[k] = q; eo[k] = ecs[k].x;
                    }
                }
                for (int k = KB; k < deg; ++k) {                                       // column weight > 3: remaining messages
                    const uint2 ec = vt[a0 + k];
                    const T q = qr[(idx_t)ec.x * S];
                    const T s1 = M1[(idx_t)ec.y * S];
                    T rr;
                    if (VAR == VAR_MIN_SUM) {
                        const T s2 = M2[(idx_t)ec.y * S];
                        const T a1 = fabs(s1);
                        const T mag = (fabs(q) == a1) ? s2 : a1;
                        rr = N::from_bits(N::bits(N::mul(alpha, mag)) ^ ((N::bits(s1) ^ N::bits(q)) & N::SIGN));
                    } else {
                        T t = slot_is_tanh ? q : N::tanh_(N::mul(q, (T)0.5));
                        const T ts = (fabs(t) < (T)1e-15) ? (T)1e-15 : t;
                        rr = bp_sp_r(N::div(s1, ts));
                        if (P.sym) rr = N::mul(rr, alpha);
                    }
                    sum = N::add(sum, rr);
                }
                const T val = N::add(sum, prior[v]);                                 // values = R_sum + prior
                const bool hd = val < (T)0;
                hw |= (uint32_t)hd << b;
                if (wr_llr) llr_out[v] = val;
                if (WMS > 0) {
                    if (hd) {
#pragma unroll
                        for (int k = 0; k < WREG; ++k) acc[k] ^= colmask[v * WREG + k];
                    }
                } else if (use_par && hd) {
                    for (int a = a0; a < a0 + deg; ++a) {
                        const uint32_t c = vt[a].y;
                        par[c >> 5] ^= 1u << (c & 31);
                    }
                }
                // Q update: Q_new = values - R; damping against Q_old; clip  (decoding.py:63-66)
#pragma unroll
                for (int k = 0; k < KB; ++k) {
                    if (k < deg) {
                        T qn = N::sub(val, r[k]);
                        if (VAR == VAR_MIN_SUM || P.sym) {
                            qn = bp_damp(damp, qn, omd, qo[k]);
                            qn = fmin(fmax(qn, -clipv), clipv);
                            qn = bp_canon(qn);
                        } else if (slot_is_tanh) {
                            qn = N::tanh_(N::mul(qn, (T)0.5));
                        }
                        st_state<STATE_SMEM>(qw + (idx_t)eo[k] * S, qn);
                    }
                }
                if (deg > KB) {
                    // generic tail (column weight > 3): recompute the message of each remaining edge
                    for (int k = KB; k < deg; ++k) {
                        const uint2 ec = vt[a0 + k];
                        const T q = qr[(idx_t)ec.x * S];
                        const T s1 = M1[(idx_t)ec.y * S];
                        T rr;
                        if (VAR == VAR_MIN_SUM) {
                            const T s2 = M2[(idx_t)ec.y * S];
                            const T a1 = fabs(s1);
                            const T mag = (fabs(q) == a1) ? s2 : a1;
                            rr = N::from_bits(N::bits(N::mul(alpha, mag)) ^ ((N::bits(s1) ^ N::bits(q)) & N::SIGN));
                            if (dumping) rdump[ec.x] = N::div(rr, alpha);
                        } else {
                            T t = slot_is_tanh ? q : N::tanh_(N::mul(q, (T)0.5));
                            const T ts = (fabs(t) < (T)1e-15) ? (T)1e-15 : t;
                            rr = bp_sp_r(N::div(s1, ts));
                            if (dumping) rdump[ec.x] = rr;
                            if (P.sym) rr = N::mul(rr, alpha);
                        }
                        T qn = N::sub(val, rr);
                        if (VAR == VAR_MIN_SUM || P.sym) {
                            qn = bp_damp(damp, qn, omd, q);
                            qn = fmin(fmax(qn, -clipv), clipv);
                            qn = bp_canon(qn);
                        } else if (slot_is_tanh) {
                            qn = N::tanh_(N::mul(qn, (T)0.5));
                        }
                        qw[(idx_t)ec.x * S] = qn;
                    }
                }
            }
            HW[(idx_t)wv * S] = hw;
        }

        // ================= syndrome of the hard decision ===================================
        bool conv = true;
        if (WMS > 0) {
#pragma unroll
            for (int k = 0; k < WREG; ++k) conv = conv && (acc[k] == synd[k]);
        } else if (use_par) {
            for (int w = 0; w < WM; ++w) conv = conv && (par[w] == SY[(idx_t)w * S]);
        } else {
            for (int w = 0; w * 32 < m && conv; ++w) {
                uint32_t par = 0;
                const int cend = min(32, m - 32 * w);
                for (int b = 0; b < cend; ++b) {
                    const int c = 32 * w + b;
                    uint32_t p1 = 0;
                    const int e1 = row_ptr[c + 1];
                    for (int e = row_ptr[c]; e < e1; ++e) {
                        const int v = g.col_idx[e];
                        p1 ^= HW[(idx_t)(v >> 5) * S] >> (v & 31);
                    }
                    par |= (p1 & 1u) << b;
                }
                conv = (par == SY[(idx_t)w * S]);
            }
        }

        if (P.r_dump) conv = false;          // `and not alpha_estimation`: no early exit while dumping (decoding.py:72,188)
        if (conv || last || dumping) {
            // ---- retire the shot -----------------------------------------------------------
            uint32_t *ho = P.hard + (size_t)shot * WN;
            for (int w = 0; w < WN; ++w) ho[w] = HW[(idx_t)w * S];
            P.conv[shot] = conv ? 1 : 0;
            if (P.iters) P.iters[shot] = iter;
            if (!conv && P.fail_idx) P.fail_idx[atomicAdd(P.fail_count, 1u)] = (int32_t)shot;
            iter_sum += (unsigned long long)(iter + 1);
            active = false;
        } else {
            ++iter;
            if (!STATE_SMEM) { T *t_ = Q; Q = QW; QW = t_; }      // ping-pong
        }
        }  // if (active)
    }

    if (P.iter_total) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) iter_sum += __shfl_xor_sync(0xffffffffu, iter_sum, o);
        if (lane == 0 && iter_sum) atomicAdd(P.iter_total, iter_sum);
    }
}

}  // namespace qldpc
