// Bit-packed ordered-statistics decoding (OSD-0) on sm_100a -- one WARP per BP-failed shot.
//
// Replaces decoding/OSD.py:3-72 (performOSD + gf2_elimination) and the OSD-0 front half of
// decoding/OSD_enhanced.py:5-64 (and its verbatim copy rework/decoding.py:193-347).
//
// Formulation.  The reference permutes the columns of H by ascending |LLR| and runs Gauss-Jordan
// on the (m, n) matrix with the residual syndrome as an extra column.  Row operations act from the
// left, so it is enough to track the transform T (m x m bits, A = T * H[:, ordering]) and the
// syndrome column b: the entry (r, j) of the reduced matrix is parity(T[r] & h_j), where h_j is the
// packed column ordering[j] of H (the same `colmask` table the BP kernel uses).  T starts as the
// identity, so nothing but `colmask` is read from memory, a row operation costs ceil(m/32) words
// instead of ceil(n/32), and the whole state of a shot lives in registers:
//     lane l owns rows r = l + 32*i (i < WM); row r of T is WM words; T[r] initially has word i =
//     1 << l.
// Per column: every lane tests its rows (AND/XOR/POPC), the pivot is the candidate row with the
// smallest POSITION (the reference swaps the pivot row up, OSD.py:56-58; positions are tracked
// instead of moving data) found with one REDUX.MIN, the pivot row is broadcast with shuffles and
// XORed into every other row that has the bit (OSD.py:64-68).
//
// Ordering contract: stable ascending sort of |LLR| (ties -> lower column index),
// computed as a rank by counting, keys compared as unsigned integers (|x| bit patterns are
// monotone; NaN sorts last like np.argsort).
//
// Outputs per shot: the solution hard ^ correction (OSD.py:23-26), whether it satisfies the
// syndrome (<=> every non-pivot row ends with b == 0; OSD_enhanced.py:58-60), and, on request,
// the elimination record the OSD-w sweep needs (ordering, pivot columns by position, reduced
// syndrome by position).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace qldpc {

struct OSDParams {
    int m, n, WM, WN;
    int rank;                     // GF(2) rank of H: no pivot can be found once `rank` rows are pivoted
    const uint32_t *colmask;      // [n][WM]
    const int32_t *idx;           // [count] shot ids to process (null: identity)
    const unsigned int *count_dev;// number of entries in idx (device) ...
    long long count_host;         // ... or given by the host when count_dev == null
    long long cap;                // upper bound of the count when it lives on the device (0: unknown)
    const uint32_t *synd;         // [B][WM]
    const void *llr;              // [B][n] float or double
    const uint32_t *hard;         // [B][WN]  BP hard decision
    uint32_t *out;                // [B][WN]  solution (may alias hard)
    uint8_t *valid;               // [B] solution satisfies the syndrome (may be null)
};

// cnt -= (a < b), as a subtract-with-borrow pair (2 instructions; the compiler's own lowering of `cnt += (a < b)` is
// add / compare / predicated move)
__device__ __forceinline__ void osd_count_lt(int &cnt, uint32_t a, uint32_t b)
{
    asm("{\n\t.reg .u32 t;\n\tsub.cc.u32 t, %1, %2;\n\tsubc.u32 %0, %0, 0;\n\t}" : "+r"(cnt) : "r"(a), "r"(b));
}
__device__ __forceinline__ void osd_count_lt(int &cnt, unsigned long long a, unsigned long long b)
{
    asm("{\n\t.reg .u32 t;\n\tsub.cc.u32 t, %1, %3;\n\tsubc.cc.u32 t, %2, %4;\n\tsubc.u32 %0, %0, 0;\n\t}"
        : "+r"(cnt)
        : "r"((uint32_t)a), "r"((uint32_t)(a >> 32)), "r"((uint32_t)b), "r"((uint32_t)(b >> 32)));
}

// The same count on the FMA pipe (the OSD kernels saturate the ALU pipe): x = a - b as a multiply-add, then the sign bit of
// x -- hi32(x * 2) -- accumulated by a second one.  cnt counts UP here; valid for a, b <= 2^31 (float keys).
__device__ __forceinline__ void osd_count_lt_fma(int &cnt, uint32_t a, uint32_t b)
{
    asm("{\n\t.reg .u32 t;\n\tmad.lo.u32 t, %2, 0xffffffff, %1;\n\tmad.hi.u32 %0, t, 2, %0;\n\t}" : "+r"(cnt) : "r"(a), "r"(b));
}

template <typename K> struct KeyBits;
template <> struct KeyBits<float> {
    typedef uint32_t type;
    __device__ static __forceinline__ type get(float x) { return __float_as_uint(x) & 0x7fffffffu; }
};
template <> struct KeyBits<double> {
    typedef unsigned long long type;
    __device__ static __forceinline__ type get(double x) { return (unsigned long long)__double_as_longlong(x) & 0x7fffffffffffffffull; }
};

// Rank of the stable argsort by counting, for a group of LPS lanes that owns one shot: lane `lid` of the group ranks its keys
// ki[t] (index LPS * t + lid) against all n keys, fetched block by block.  Keys of blocks before the lane's own compare with
// <=, keys of later blocks with <, only the diagonal block needs the tie rule per lane (2 instructions per 32-bit compare:
// the subtraction as an IMAD on the FMA pipe, the sign accumulated by a LEA.HI; 3 on the ALU pipe for 64-bit keys).
// cnt[t] = #{j : key_j < key_i, or key_j == key_i and j < i}.
template <typename KB, int NSLOT, int LPS, typename Fetch>
__device__ __forceinline__ void osd_rank_count(int n, int lid, const KB (&ki)[NSLOT], int (&cnt)[NSLOT], Fetch fetch)
{
#pragma unroll
    for (int t = 0; t < NSLOT; ++t) cnt[t] = 0;
#pragma unroll
    for (int jb = 0; jb < NSLOT; ++jb) {                                 // keys LPS jb .. LPS jb + LPS - 1
        KB thr[NSLOT];
#pragma unroll
        for (int t = 0; t < NSLOT; ++t) thr[t] = (jb < t) ? ki[t] + (KB)1 : ki[t];         // j < i: count key_j <= key_i
        const int jend = min(LPS, n - LPS * jb);
#pragma unroll 4
        for (int jl = 0; jl < jend; ++jl) {
            const KB kj = fetch(LPS * jb + jl);
#pragma unroll
            for (int t = 0; t < NSLOT; ++t) {
                const KB bound = (t == jb) ? ki[t] + (KB)(jl < lid) : thr[t];               // diagonal block: tie rule per lane
                if constexpr (sizeof(KB) == 4) {
                    osd_count_lt_fma(cnt[t], kj, bound);                   // subtraction on the FMA pipe, counts up
                } else {
                    osd_count_lt(cnt[t], kj, bound);                       // borrow chain, counts down
                }
            }
        }
    }
    if constexpr (sizeof(KB) == 8) {
#pragma unroll
        for (int t = 0; t < NSLOT; ++t) cnt[t] = -cnt[t];
    }
}

// The same count for float64 keys, compared as DOUBLES: the key pattern of |x| is the double |x|, DSETP runs on the FP64 pipe
// and the counter is bumped by a predicated IMAD on the FMA pipe -- two issue slots per compare and not one cycle of the ALU
// pipe, which bounds the OSD kernels (the 64-bit integer borrow chain is three ALU instructions, two cycles each).  The
// tie rule is carried by the comparison: <= for keys of earlier blocks, < for later ones, (<) or (== and j < i) on the
// diagonal block.  NaN keys compare false: the caller falls back to the integer ranking when a shot holds one.
template <int NSLOT, int LPS>
__device__ __forceinline__ void osd_rank_count_f64(int n, int lid, const double (&ki)[NSLOT], int (&cnt)[NSLOT], const double *keys, int one)
{
#pragma unroll
    for (int t = 0; t < NSLOT; ++t) cnt[t] = 0;
#pragma unroll
    for (int jb = 0; jb < NSLOT; ++jb) {
        const int jend = min(LPS, n - LPS * jb);
#pragma unroll 4
        for (int jl = 0; jl < jend; ++jl) {
            const double kj = keys[LPS * jb + jl];
            const int before = (jl < lid) ? 1 : 0;
#pragma unroll
            for (int t = 0; t < NSLOT; ++t) {
                if (t > jb)
                    asm("{\n\t.reg .pred p;\n\tsetp.le.f64 p, %1, %2;\n\t@p mad.lo.s32 %0, %3, %3, %0;\n\t}" : "+r"(cnt[t]) : "d"(kj), "d"(ki[t]), "r"(one));
                else if (t < jb)
                    asm("{\n\t.reg .pred p;\n\tsetp.lt.f64 p, %1, %2;\n\t@p mad.lo.s32 %0, %3, %3, %0;\n\t}" : "+r"(cnt[t]) : "d"(kj), "d"(ki[t]), "r"(one));
                else
                    asm("{\n\t.reg .pred p, q, r;\n\tsetp.ne.s32 r, %4, 0;\n\tsetp.eq.and.f64 q, %1, %2, r;\n\tsetp.lt.or.f64 p, %1, %2, q;\n\t"
                        "@p mad.lo.s32 %0, %3, %3, %0;\n\t}" : "+r"(cnt[t]) : "d"(kj), "d"(ki[t]), "r"(one), "r"(before));
            }
        }
    }
}

// The ordering of one shot into ord[] (shared memory; keys[] already holds the n key patterns, ki[] the lane's own).
template <typename KB, int NSLOT, int LPS>
__device__ __forceinline__ void osd_order_shot(int n, int lid, const KB *keys, const KB (&ki)[NSLOT], uint16_t *ord)
{
    int cnt[NSLOT];
    bool as_doubles = false;
    if constexpr (sizeof(KB) == 8) {
        bool nan = false;
#pragma unroll
        for (int t = 0; t < NSLOT; ++t) nan = nan || (LPS * t + lid < n && ki[t] > (KB)0x7ff0000000000000ull);
        as_doubles = !__any_sync(0xffffffffu, nan);                        // (warp-uniform: both halves of a two-shot warp agree)
        if (as_doubles) {
            double kd[NSLOT];
#pragma unroll
            for (int t = 0; t < NSLOT; ++t) kd[t] = __longlong_as_double((long long)ki[t]);      // (padding: a NaN pattern, never stored)
            osd_rank_count_f64<NSLOT, LPS>(n, lid, kd, cnt, reinterpret_cast<const double *>(keys), n > 0 ? 1 : 0);
        }
    }
    if constexpr (sizeof(KB) == 4) {
        osd_rank_count<KB, NSLOT, LPS>(n, lid, ki, cnt, [&](int j) { return keys[j]; });
    } else if (!as_doubles) {
        // a NaN among the keys (sorts last, like np.argsort): plain integer ranking, kept small -- it practically never runs
#pragma unroll
        for (int t = 0; t < NSLOT; ++t) {                                  // (unrolled over the register arrays, not over j)
            const int i = LPS * t + lid;
            int c = 0;
#pragma unroll 1
            for (int j = 0; j < n; ++j) c += (keys[j] < ki[t]) || (keys[j] == ki[t] && j < i);
            cnt[t] = c;
        }
    }
#pragma unroll
    for (int t = 0; t < NSLOT; ++t)
        if (LPS * t + lid < n) ord[cnt[t]] = (uint16_t)(LPS * t + lid);
    __syncwarp();
}

constexpr int OSD_WARPS = 4;

// dynamic shared memory: one copy of colmask per CTA, then per warp: n keys + n ordering entries + solution words
__host__ __device__ inline size_t osd_smem_colmask(int n, int WM) { return (4 * (size_t)n * WM + 15) & ~(size_t)15; }
template <typename K>
__host__ __device__ inline size_t osd_smem_per_warp(int n)
{
    size_t o = sizeof(typename KeyBits<K>::type) * (size_t)n;
    o += sizeof(uint16_t) * (size_t)n;
    o = (o + 3) & ~(size_t)3;
    o += sizeof(uint32_t) * (size_t)((n + 31) / 32);   // solution words
    return (o + 15) & ~(size_t)15;
}

// Row-major elimination of one shot by one warp, given its ordering `ord` (shared memory): residual syndrome,
// gf2_elimination with the reference's pivot-row rule, solution, validity flag, optional elimination record.
// (inputs / outputs of ONE shot; the record pointers may be null, or point into shared memory -- osdw_kernel.cuh)
struct OSDShotIO {
    const uint32_t *hard;         // [WN]
    const uint32_t *synd;         // [WM]
    uint32_t *out;                // [WN]
    uint8_t *valid;               // [1] or null
    int32_t *rec_ordering;        // [n] or null: with it rec_pivcol [m], rec_sred [m], rec_npiv [1]
    int32_t *rec_pivcol;
    uint8_t *rec_sred;
    int32_t *rec_npiv;
};

template <int WM>
__device__ __forceinline__ void osd0_rowmajor_shot(int m, int n, int WN, int rank_of_H, const OSDShotIO &io, const uint32_t *cmask,
                                                   const uint16_t *ord, uint32_t *solw, int lane)
{
    const unsigned FULL = 0xffffffffu;
    const uint32_t *hard = io.hard;
            // ---- residual syndrome s ^ H*hard (OSD.py:7-8) as packed words, warp-uniform --------
            uint32_t rs[WM];
    #pragma unroll
            for (int w = 0; w < WM; ++w) rs[w] = 0;
            for (int v = lane; v < n; v += 32) {
                if ((hard[v >> 5] >> (v & 31)) & 1u) {
    #pragma unroll
                    for (int w = 0; w < WM; ++w) rs[w] ^= cmask[v * WM + w];
                }
            }
    #pragma unroll
            for (int w = 0; w < WM; ++w) rs[w] = __reduce_xor_sync(FULL, rs[w]) ^ io.synd[w];

            // ---- T = I, b = residual, positions = row index ------------------------------------
            uint32_t T[WM][WM];
            uint32_t bbit[WM];
            int pos[WM], pcol[WM];
    #pragma unroll
            for (int i = 0; i < WM; ++i) {
    #pragma unroll
                for (int w = 0; w < WM; ++w) T[i][w] = (w == i) ? (1u << lane) : 0u;
                const int r = lane + 32 * i;
                bbit[i] = (rs[i] >> lane) & 1u;
                pos[i] = (r < m) ? r : 0x7fffffff;
                pcol[i] = -1;
            }

            // ---- gf2_elimination (OSD.py:31-72) -------------------------------------------------
            int row = 0;
            const int rank = rank_of_H;
            for (int j = 0; j < n && row < rank; ++j) {     // `row >= m` (:43); past `rank` pivots no column can pivot
                const int col = ord[j];
                uint32_t cm[WM];
    #pragma unroll
                for (int w = 0; w < WM; ++w) cm[w] = cmask[col * WM + w];
                uint32_t has[WM];
                int best = 0x7fffffff;
    #pragma unroll
                for (int i = 0; i < WM; ++i) {
                    uint32_t x = 0;
    #pragma unroll
                    for (int w = 0; w < WM; ++w) x ^= T[i][w] & cm[w];
                    has[i] = __popc(x) & 1u;
                    if (has[i] && pos[i] >= row && pos[i] < best) best = pos[i];   // first row >= current (:47-50)
                }
                const int pmin = __reduce_min_sync(FULL, best);
                if (pmin == 0x7fffffff) continue;                                    // no pivot in this column (:52-53)
                // the owner of the pivot row broadcasts it
                uint32_t prow[WM];
                uint32_t pb = 0;
                bool mine = false;
    #pragma unroll
                for (int w = 0; w < WM; ++w) prow[w] = 0;
    #pragma unroll
                for (int i = 0; i < WM; ++i) {
                    if (pos[i] == pmin) {
                        mine = true;
                        pb = bbit[i];
    #pragma unroll
                        for (int w = 0; w < WM; ++w) prow[w] = T[i][w];
                    }
                }
                const int owner = __ffs(__ballot_sync(FULL, mine)) - 1;
                pb = __shfl_sync(FULL, pb, owner);
    #pragma unroll
                for (int w = 0; w < WM; ++w) prow[w] = __shfl_sync(FULL, prow[w], owner);
                // swap the pivot row up (:56-58) == exchange positions; then eliminate all other rows (:64-68)
    #pragma unroll
                for (int i = 0; i < WM; ++i) {
                    if (pos[i] == pmin) {
                        pos[i] = row;
                        pcol[i] = j;
                    } else {
                        if (pos[i] == row) pos[i] = pmin;
                        if (has[i]) {
    #pragma unroll
                            for (int w = 0; w < WM; ++w) T[i][w] ^= prow[w];
                            bbit[i] ^= pb;
                        }
                    }
                }
                ++row;
            }

            // ---- e_permuted[pivot col] = s_reduced[pivot row]; unpermute; xor hard (OSD.py:16-26) ----
            // validity: every non-pivot row must end with b == 0
            bool bad = false;
    #pragma unroll
            for (int i = 0; i < WM; ++i) bad = bad || (pcol[i] < 0 && bbit[i] && (lane + 32 * i) < m);
            const bool any_bad = __any_sync(FULL, bad);
            for (int w = lane; w < WN; w += 32) solw[w] = hard[w];
            __syncwarp();
    #pragma unroll
            for (int i = 0; i < WM; ++i) {
                if (pcol[i] >= 0 && bbit[i]) {
                    const int v = ord[pcol[i]];
                    atomicXor(&solw[v >> 5], 1u << (v & 31));
                }
            }
            __syncwarp();
            for (int w = lane; w < WN; w += 32) io.out[w] = solw[w];
            if (lane == 0 && io.valid) *io.valid = any_bad ? 0 : 1;

            // ---- elimination record for the OSD-w sweep -----------------------------------------
            if (io.rec_ordering) {
                for (int j = lane; j < n; j += 32) io.rec_ordering[j] = ord[j];
                for (int r = lane; r < m; r += 32) io.rec_pivcol[r] = -1;
                __syncwarp();
    #pragma unroll
                for (int i = 0; i < WM; ++i) {
                    if (lane + 32 * i < m) {
                        io.rec_pivcol[pos[i]] = pcol[i];
                        io.rec_sred[pos[i]] = (uint8_t)bbit[i];
                    }
                }
                if (lane == 0) *io.rec_npiv = row;
            }
}

// the same on the batch arrays of OSDParams (no record)
template <typename K, int WM>
__device__ __forceinline__ void osd0_rowmajor_shot(const OSDParams &P, long long it, long long shot, const uint32_t *cmask,
                                                   const uint16_t *ord, uint32_t *solw, int lane)
{
    OSDShotIO io;
    io.hard = P.hard + (size_t)shot * P.WN;
    io.synd = P.synd + (size_t)shot * P.WM;
    io.out = P.out + (size_t)shot * P.WN;
    io.valid = P.valid ? P.valid + shot : nullptr;
    io.rec_ordering = nullptr; io.rec_pivcol = nullptr; io.rec_sred = nullptr; io.rec_npiv = nullptr;
    (void)it;
    osd0_rowmajor_shot<WM>(P.m, P.n, P.WN, P.rank, io, cmask, ord, solw, lane);
}

template <typename K, int WM>
__global__ void __launch_bounds__(OSD_WARPS * 32)
osd0_kernel(const OSDParams P)
{
    typedef typename KeyBits<K>::type kbits;
    const int n = P.n;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    extern __shared__ __align__(16) unsigned char smem[];
    uint32_t *cmask = reinterpret_cast<uint32_t *>(smem);
    for (int i = threadIdx.x; i < n * WM; i += blockDim.x) cmask[i] = P.colmask[i];
    __syncthreads();
    unsigned char *base = smem + osd_smem_colmask(n, WM) + osd_smem_per_warp<K>(n) * warp;
    kbits *keys = reinterpret_cast<kbits *>(base);
    uint16_t *ord = reinterpret_cast<uint16_t *>(base + sizeof(kbits) * (size_t)n);
    uint32_t *solw = reinterpret_cast<uint32_t *>(base + ((sizeof(kbits) * (size_t)n + sizeof(uint16_t) * (size_t)n + 3) & ~(size_t)3));

    const long long count = P.count_dev ? (long long)*P.count_dev : P.count_host;
    const long long nwarps = (long long)gridDim.x * OSD_WARPS;

    for (long long it = (long long)blockIdx.x * OSD_WARPS + warp; it < count; it += nwarps) {
        const long long shot = P.idx ? (long long)P.idx[it] : it;
        const K *llr = reinterpret_cast<const K *>(P.llr) + (size_t)shot * n;

        // ---- ordering = argsort(|llr|), stable (OSD.py:10-11) ------------------------------
        for (int j = lane; j < n; j += 32) keys[j] = KeyBits<K>::get(llr[j]);
        __syncwarp();
        // rank of key i = #{j : key_j < key_i, or key_j == key_i and j < i}; each lane ranks up to KPL of its keys
        // per pass over all n keys (one pass for n <= 32 * KPL)
        constexpr int KPL = (WM <= 3) ? 5 : 9;                  // covers n <= 160 / 288 in a single pass
        for (int i0 = 0; i0 < n; i0 += 32 * KPL) {
            kbits ki[KPL];
            int ii[KPL], cnt[KPL];
#pragma unroll
            for (int t = 0; t < KPL; ++t) {
                ii[t] = i0 + 32 * t + lane;
                ki[t] = (ii[t] < n) ? keys[ii[t]] : (kbits)0;
                cnt[t] = 0;
            }
#pragma unroll 2
            for (int j = 0; j < n; ++j) {
                const kbits kj = keys[j];
#pragma unroll
                for (int t = 0; t < KPL; ++t) cnt[t] += (kj < ki[t] + (kbits)(j < ii[t]));   // kj <= ki for j < i (keys < 2^(bits-1))
            }
#pragma unroll
            for (int t = 0; t < KPL; ++t) if (ii[t] < n) ord[cnt[t]] = (uint16_t)ii[t];
        }
        __syncwarp();

        osd0_rowmajor_shot<K, WM>(P, it, shot, cmask, ord, solw, lane);
        __syncwarp();
    }
}


// ------------------------------------------------------------------------------------------------
// OSD-0, column-major: the production kernel when no elimination record is requested (BP+OSD pipelines).
//
// The OSD-0 solution does not depend on WHICH row pivots a column: the pivot columns are the greedy set of linearly
// independent columns in |LLR| order (independent of the row choices), and the error supported on them that reproduces
// the syndrome is unique.  The reference's "first row at or below the current one" rule (OSD.py:47-58) therefore only
// matters for the elimination record of the OSD-w sweep, which osd0_kernel above still produces.
//
// That freedom allows the cheap formulation: lane l keeps the REDUCED columns at sorted positions 32 s + l (s < NS) in
// registers as WM-word bit vectors, the syndrome column and the set of used pivot rows are warp-uniform registers.
// Per column (broadcast from its owner with WM shuffles): free rows = column & ~used; none -> dependent column, next;
// else the lowest free row p pivots, and every later column c with bit p set becomes c ^ (column - bit p) -- one
// predicated 3-word XOR per register slot, on all 32 lanes at once.  About 45 warp-instructions per pivot and 10 per
// dependent column instead of ~55 + 70 per column in the row-major kernel, and no REDUX / pivot-row search.
// The rank of the stable argsort is counted block by block: keys before a lane's own 32-block compare with <=, keys
// after it with <, only the diagonal block needs the tie rule per lane (2 instructions per compare instead of 6).
// ------------------------------------------------------------------------------------------------
template <typename K, int WM, int NS>
__global__ void __launch_bounds__(OSD_WARPS * 32)
osd0_fast_kernel(const OSDParams P)
{
    typedef unsigned long long kbits;       // keys are always |double(llr)|: float32 LLRs are widened (order and ties are preserved)
    const int m = P.m, n = P.n, WN = P.WN;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    extern __shared__ __align__(16) unsigned char smem[];
    uint32_t *cmask = reinterpret_cast<uint32_t *>(smem);
    for (int i = threadIdx.x; i < n * WM; i += blockDim.x) cmask[i] = P.colmask[i];
    __syncthreads();
    unsigned char *base = smem + osd_smem_colmask(n, WM) + osd_smem_per_warp<double>(n) * warp;
    kbits *keys = reinterpret_cast<kbits *>(base);
    uint16_t *ord = reinterpret_cast<uint16_t *>(base + sizeof(kbits) * (size_t)n);
    uint32_t *solw = reinterpret_cast<uint32_t *>(base + ((sizeof(kbits) * (size_t)n + sizeof(uint16_t) * (size_t)n + 3) & ~(size_t)3));

    const long long count = P.count_dev ? (long long)*P.count_dev : P.count_host;
    const long long nwarps = (long long)gridDim.x * OSD_WARPS;
    const unsigned FULL = 0xffffffffu;
    const int rank = P.rank;

    for (long long it = (long long)blockIdx.x * OSD_WARPS + warp; it < count; it += nwarps) {
        const long long shot = P.idx ? (long long)P.idx[it] : it;
        const K *llr = reinterpret_cast<const K *>(P.llr) + (size_t)shot * n;
        const uint32_t *hard = P.hard + (size_t)shot * WN;

        // ---- ordering = argsort(|llr|), stable (OSD.py:10-11): rank by counting --------------
        kbits ki[NS];
#pragma unroll
        for (int t = 0; t < NS; ++t) {
            const int i = 32 * t + lane;
            ki[t] = (i < n) ? KeyBits<double>::get((double)llr[i]) : ~(kbits)0;        // padding sorts last
            if (i < n) keys[i] = ki[t];
        }
        __syncwarp();
        osd_order_shot<kbits, NS, 32>(n, lane, keys, ki, ord);

        // ---- residual syndrome s ^ H*hard (OSD.py:7-8) as packed words, warp-uniform --------
        uint32_t b[WM];
#pragma unroll
        for (int w = 0; w < WM; ++w) b[w] = 0;
        for (int v = lane; v < n; v += 32) {
            if ((hard[v >> 5] >> (v & 31)) & 1u) {
#pragma unroll
                for (int w = 0; w < WM; ++w) b[w] ^= cmask[v * WM + w];
            }
        }
#pragma unroll
        for (int w = 0; w < WM; ++w) b[w] = __reduce_xor_sync(FULL, b[w]) ^ P.synd[(size_t)shot * WM + w];

        // ---- the columns in sorted order, WM words each -------------------------------------
        uint32_t c[NS][WM];
        int prow[NS];
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            const int j = 32 * s + lane;
            const int col = (j < n) ? ord[j] : 0;
#pragma unroll
            for (int w = 0; w < WM; ++w) c[s][w] = (j < n) ? cmask[col * WM + w] : 0u;
            prow[s] = -1;
        }

        // ---- elimination ----------------------------------------------------------------------
        uint32_t used[WM];
#pragma unroll
        for (int w = 0; w < WM; ++w) used[w] = 0;
        int npiv = 0;
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            // A column without a free row can never get one back (it is only updated through a free row it holds), so
            // the next column to pivot is the first lane of the slot, past the last pivot, that still has a free row:
            // one ballot per PIVOT; dependent columns cost nothing.
            unsigned todo = FULL;
            while (npiv < rank) {
                uint32_t freebits = 0;
#pragma unroll
                for (int w = 0; w < WM; ++w) freebits |= c[s][w] & ~used[w];
                const unsigned live = __ballot_sync(FULL, freebits != 0) & todo;
                if (live == 0) break;
                const int l = __ffs(live) - 1;
                todo &= ~((2u << l) - 1u);
                uint32_t col[WM], fr[WM];
#pragma unroll
                for (int w = 0; w < WM; ++w) {
                    col[w] = __shfl_sync(FULL, c[s][w], l);
                    fr[w] = col[w] & ~used[w];
                }
                int pw = WM - 1;
                uint32_t fw = fr[WM - 1];
#pragma unroll
                for (int w = WM - 2; w >= 0; --w)
                    if (fr[w] != 0) { pw = w; fw = fr[w]; }                  // first word with a free row
                const uint32_t pbit = fw & (0u - fw);                        // lowest free row of the column
                if (lane == l) prow[s] = 32 * pw + __ffs(pbit) - 1;
                ++npiv;
                // pw is warp-uniform: one specialised copy of the update per word
#pragma unroll
                for (int w0 = 0; w0 < WM; ++w0) {
                    if (pw == w0) {
                        used[w0] |= pbit;
                        col[w0] &= ~pbit;                                    // the pivot row itself is not touched
                        if (b[w0] & pbit) {
#pragma unroll
                            for (int w = 0; w < WM; ++w) b[w] ^= col[w];
                        }
#pragma unroll
                        for (int s2 = s; s2 < NS; ++s2) {                    // (earlier positions are finished)
                            if (c[s2][w0] & pbit) {
#pragma unroll
                                for (int w = 0; w < WM; ++w) c[s2][w] ^= col[w];
                            }
                        }
                    }
                }
            }
        }

        // ---- e[pivot column] = reduced syndrome at its pivot row; xor hard (OSD.py:16-26) -----
        // validity: every row that pivots no column must end with b == 0
        bool bad = false;
#pragma unroll
        for (int w = 0; w < WM; ++w) {
            const int rows = m - 32 * w;                                     // rows of H in this word
            const uint32_t live = rows >= 32 ? 0xffffffffu : (rows > 0 ? ((1u << rows) - 1u) : 0u);
            bad = bad || ((b[w] & ~used[w] & live) != 0);
        }
        if (bad) {
            // Inconsistent syndrome (never the case for s = e H^T): the reference's output then depends on its pivot-row
            // choices, so this shot is redone with the row-major rule.
            osd0_rowmajor_shot<K, WM>(P, it, shot, cmask, ord, solw, lane);
            __syncwarp();
            continue;
        }
        for (int w = lane; w < WN; w += 32) solw[w] = hard[w];
        __syncwarp();
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            if (prow[s] >= 0) {
                uint32_t bw = b[0];
#pragma unroll
                for (int w = 1; w < WM; ++w) if ((prow[s] >> 5) == w) bw = b[w];
                if ((bw >> (prow[s] & 31)) & 1u) {
                    const int v = ord[32 * s + lane];
                    atomicXor(&solw[v >> 5], 1u << (v & 31));
                }
            }
        }
        __syncwarp();
        for (int w = lane; w < WN; w += 32) P.out[(size_t)shot * WN + w] = solw[w];
        if (lane == 0 && P.valid) P.valid[shot] = 1;
        __syncwarp();
    }
}


// ------------------------------------------------------------------------------------------------
// osd0_fast_kernel with TWO shots per warp (16 lanes each): the ~50 warp-level instructions that find and broadcast a
// pivot, and every instruction of the column update, now serve two shots, and a lane's rank counting serves twice as
// many of its own keys per broadcast key.  Nothing is warp-uniform any more (the halves pivot different rows), so the
// pivot word is picked with selects and a half without a pivot left runs the round with an empty pivot mask (no-ops).
// The two halves walk the register slots together (slot s holds the columns at sorted positions 16 s + lane-in-half).
// ------------------------------------------------------------------------------------------------
template <typename K, int WM, int NS2>
__global__ void __launch_bounds__(OSD_WARPS * 32)
osd0_fast2_kernel(const OSDParams P)
{
    typedef unsigned long long kbits;       // (see osd0_fast_kernel)
    const int m = P.m, n = P.n, WN = P.WN;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, half = lane >> 4, hl = lane & 15;
    extern __shared__ __align__(16) unsigned char smem[];
    uint32_t *cmask = reinterpret_cast<uint32_t *>(smem);
    for (int i = threadIdx.x; i < n * WM; i += blockDim.x) cmask[i] = P.colmask[i];
    __syncthreads();
    unsigned char *base = smem + osd_smem_colmask(n, WM) + osd_smem_per_warp<double>(n) * (2 * warp + half);
    kbits *keys = reinterpret_cast<kbits *>(base);
    uint16_t *ord = reinterpret_cast<uint16_t *>(base + sizeof(kbits) * (size_t)n);
    uint32_t *solw = reinterpret_cast<uint32_t *>(base + ((sizeof(kbits) * (size_t)n + sizeof(uint16_t) * (size_t)n + 3) & ~(size_t)3));

    const long long count = P.count_dev ? (long long)*P.count_dev : P.count_host;
    const long long npairs = (long long)gridDim.x * OSD_WARPS;
    const unsigned FULL = 0xffffffffu;
    const int rank = P.rank;

    for (long long pair = (long long)blockIdx.x * OSD_WARPS + warp; 2 * pair < count; pair += npairs) {
        const long long it = 2 * pair + half;
        const bool live_half = it < count;                                  // (an odd count leaves the last half empty)
        const long long shot = live_half ? (P.idx ? (long long)P.idx[it] : it) : 0;
        const K *llr = reinterpret_cast<const K *>(P.llr) + (size_t)shot * n;
        const uint32_t *hard = P.hard + (size_t)shot * WN;

        // ---- ordering = argsort(|llr|), stable (OSD.py:10-11): rank by counting within the half ----------
        kbits ki[NS2];
#pragma unroll
        for (int t = 0; t < NS2; ++t) {
            const int i = 16 * t + hl;
            ki[t] = (i < n) ? KeyBits<double>::get((double)llr[i]) : ~(kbits)0;        // padding sorts last
            if (i < n) keys[i] = ki[t];
        }
        __syncwarp();
        osd_order_shot<kbits, NS2, 16>(n, hl, keys, ki, ord);

        // ---- residual syndrome s ^ H*hard (OSD.py:7-8), uniform within the half --------------------------
        uint32_t b[WM];
#pragma unroll
        for (int w = 0; w < WM; ++w) b[w] = 0;
        for (int v = hl; v < n; v += 16) {
            if ((hard[v >> 5] >> (v & 31)) & 1u) {
#pragma unroll
                for (int w = 0; w < WM; ++w) b[w] ^= cmask[v * WM + w];
            }
        }
#pragma unroll
        for (int w = 0; w < WM; ++w) {
#pragma unroll
            for (int o = 8; o >= 1; o >>= 1) b[w] ^= __shfl_xor_sync(FULL, b[w], o);
            b[w] ^= P.synd[(size_t)shot * WM + w];
        }

        // ---- the columns in sorted order, WM words each ---------------------------------------------------
        uint32_t c[NS2][WM];
        int prow[NS2];
#pragma unroll
        for (int s = 0; s < NS2; ++s) {
            const int j = 16 * s + hl;
            const int col = (j < n) ? ord[j] : 0;
#pragma unroll
            for (int w = 0; w < WM; ++w) c[s][w] = (j < n && live_half) ? cmask[col * WM + w] : 0u;
            prow[s] = -1;
        }

        // ---- elimination -------------------------------------------------------------------------------------
        uint32_t used[WM];
#pragma unroll
        for (int w = 0; w < WM; ++w) used[w] = 0;
        int npiv = 0;
#pragma unroll
        for (int s = 0; s < NS2; ++s) {
            unsigned todo = 0xffffu;
            while (true) {
                uint32_t freebits = 0;
#pragma unroll
                for (int w = 0; w < WM; ++w) freebits |= c[s][w] & ~used[w];
                const unsigned mine = (__ballot_sync(FULL, freebits != 0) >> (16 * half)) & todo;
                const bool act = (npiv < rank) && (mine != 0);                 // this half pivots in this round
                if (!__any_sync(FULL, act)) break;
                const int l = act ? (__ffs(mine) - 1) : 0;
                if (act) todo &= ~((2u << l) - 1u);
                uint32_t col[WM], fr[WM];
#pragma unroll
                for (int w = 0; w < WM; ++w) {
                    col[w] = __shfl_sync(FULL, c[s][w], 16 * half + l);
                    fr[w] = col[w] & ~used[w];
                }
                int pw = WM - 1;
                uint32_t fw = fr[WM - 1];
#pragma unroll
                for (int w = WM - 2; w >= 0; --w)
                    if (fr[w] != 0) { pw = w; fw = fr[w]; }                  // first word with a free row
                const uint32_t pbit = act ? (fw & (0u - fw)) : 0u;           // empty mask: the round is a no-op for this half
                if (act && hl == l) prow[s] = 32 * pw + __ffs(pbit) - 1;
                npiv += act ? 1 : 0;
                uint32_t bsel = b[0];
#pragma unroll
                for (int w = 0; w < WM; ++w) {
                    const uint32_t pm = (w == pw) ? pbit : 0u;
                    used[w] |= pm;
                    col[w] &= ~pm;                                           // the pivot row itself is not touched
                    if (w == pw) bsel = b[w];
                }
                const uint32_t bm = (bsel & pbit) ? 0xffffffffu : 0u;
#pragma unroll
                for (int w = 0; w < WM; ++w) b[w] ^= col[w] & bm;
#pragma unroll
                for (int s2 = s; s2 < NS2; ++s2) {                            // (earlier positions are finished)
                    uint32_t cw = c[s2][0];
#pragma unroll
                    for (int w = 1; w < WM; ++w) if (w == pw) cw = c[s2][w];
                    const uint32_t cm = (cw & pbit) ? 0xffffffffu : 0u;
#pragma unroll
                    for (int w = 0; w < WM; ++w) c[s2][w] ^= col[w] & cm;
                }
            }
        }

        // ---- validity; e[pivot column] = reduced syndrome at its pivot row; xor hard (OSD.py:16-26) -----
        bool bad = false;
#pragma unroll
        for (int w = 0; w < WM; ++w) {
            const int rows = m - 32 * w;
            const uint32_t live = rows >= 32 ? 0xffffffffu : (rows > 0 ? ((1u << rows) - 1u) : 0u);
            bad = bad || ((b[w] & ~used[w] & live) != 0);
        }
        bad = bad && live_half;
        if (!bad && live_half) {
            for (int w = hl; w < WN; w += 16) solw[w] = hard[w];
        }
        __syncwarp();
        if (!bad && live_half) {
#pragma unroll
            for (int s = 0; s < NS2; ++s) {
                if (prow[s] >= 0) {
                    uint32_t bw = b[0];
#pragma unroll
                    for (int w = 1; w < WM; ++w) if ((prow[s] >> 5) == w) bw = b[w];
                    if ((bw >> (prow[s] & 31)) & 1u) {
                        const int v = ord[16 * s + hl];
                        atomicXor(&solw[v >> 5], 1u << (v & 31));
                    }
                }
            }
        }
        __syncwarp();
        if (!bad && live_half) {
            for (int w = hl; w < WN; w += 16) P.out[(size_t)shot * WN + w] = solw[w];
            if (hl == 0 && P.valid) P.valid[shot] = 1;
        }
        // Inconsistent syndrome (never the case for s = e H^T): redo that shot with the row-major rule, the whole warp on it.
        const unsigned badmask = __ballot_sync(FULL, bad);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if ((badmask >> (16 * h)) & 1u) {
                const long long it_h = 2 * pair + h;
                const long long shot_h = P.idx ? (long long)P.idx[it_h] : it_h;
                unsigned char *base_h = smem + osd_smem_colmask(n, WM) + osd_smem_per_warp<double>(n) * (2 * warp + h);
                uint16_t *ord_h = reinterpret_cast<uint16_t *>(base_h + sizeof(kbits) * (size_t)n);
                uint32_t *solw_h = reinterpret_cast<uint32_t *>(base_h + ((sizeof(kbits) * (size_t)n + sizeof(uint16_t) * (size_t)n + 3) & ~(size_t)3));
                osd0_rowmajor_shot<K, WM>(P, it_h, shot_h, cmask, ord_h, solw_h, lane);
                __syncwarp();
            }
        }
        __syncwarp();
    }
}


// ------------------------------------------------------------------------------------------------
// OSD-0 for large check matrices (space-time / detector-error-model H: m > 160), one CTA per shot.
//
// Same algorithm and the same transform-matrix formulation as osd0_kernel, but T (m x m bits), the
// syndrome column, the row positions and the sort keys live in shared memory and the columns of H
// are read in sparse form (their <= few check indices) instead of packed words: entry (r, j) of the
// reduced matrix is the XOR of the bits T[r][c] over the checks c of column ordering[j].
// Row r is owned by thread r % blockDim.x.  Per column: every thread evaluates its rows, the pivot
// (smallest position >= current row among rows with the bit set) is found with a warp REDUX.MIN and a
// shared atomicMin, the pivot row is XORed into the other rows that have the bit.
// ------------------------------------------------------------------------------------------------
struct OSDBlockParams {
    int m, n, WM, WN, rank;
    int max_col_w;                // largest column weight of H
    const int32_t *var_ptr;       // [n+1]  CSC of H
    const uint32_t *vtab;         // [2E]   (edge, check) pairs per variable (any order)
    const uint32_t *colpack;      // [n]    checks of a column, 3 x 10 bits + count << 30 (null unless m <= 1024 and column weight <= 3)
    const int32_t *idx;
    const unsigned int *count_dev;
    long long count_host;
    const uint32_t *synd;         // [B][WM]
    const void *llr;              // [B][n]
    const uint32_t *hard;         // [B][WN]
    uint32_t *out;                // [B][WN]
    uint8_t *valid;               // [B]
    // osd0_block_fast_kernel only: shots whose syndrome turns out inconsistent are appended here and redone by
    // osd0_block_kernel (the reference's output then depends on its pivot-row rule)
    int32_t *redo_idx;            // [capacity >= count]
    unsigned int *redo_count;
};

constexpr int OSDB_THREADS = 256;

template <typename K>
__host__ __device__ inline size_t osdb_smem_bytes(int m, int n)
{
    const int WM = (m + 31) / 32, WN = (n + 31) / 32;
    size_t o = 4 * (size_t)m * WM;                              // T
    o += 2 * (size_t)m * 2;                                     // pos (int16), pcol (uint16)
    o = (o + 3) & ~(size_t)3;
    o += 4 * (size_t)WM * 2;                                    // residual syndrome words, b words
    o += 4 * (size_t)WN;                                        // solution words
    o = (o + 7) & ~(size_t)7;
    o += sizeof(typename KeyBits<K>::type) * (size_t)n;         // keys
    o += 2 * (size_t)n;                                         // ordering (uint16)
    return o + 64;
}

template <typename K>
__global__ void __launch_bounds__(OSDB_THREADS) osd0_block_kernel(const OSDBlockParams P)
{
    typedef typename KeyBits<K>::type kbits;
    const int m = P.m, n = P.n, WM = P.WM, WN = P.WN;
    const int tid = threadIdx.x, NT = OSDB_THREADS, lane = tid & 31;
    extern __shared__ __align__(16) unsigned char smem[];
    uint32_t *T = reinterpret_cast<uint32_t *>(smem);                       // [m][WM]
    int16_t *pos = reinterpret_cast<int16_t *>(T + (size_t)m * WM);       // [m]  (16-bit: two CTAs per SM fit for float keys)
    uint16_t *pcol = reinterpret_cast<uint16_t *>(pos + m);                // [m]  0xFFFF: not a pivot row
    uint32_t *bw = reinterpret_cast<uint32_t *>((reinterpret_cast<uintptr_t>(pcol + m) + 3) & ~(uintptr_t)3);   // [WM] syndrome column
    uint32_t *rsw = bw + WM;                                                // [WM] scratch
    uint32_t *solw = rsw + WM;                                              // [WN]
    kbits *keys = reinterpret_cast<kbits *>((reinterpret_cast<uintptr_t>(solw + WN) + 7) & ~(uintptr_t)7);
    uint16_t *ord = reinterpret_cast<uint16_t *>(keys + n);
    __shared__ int s_pmin[2], s_prow;

    const long long count = P.count_dev ? (long long)*P.count_dev : P.count_host;
    for (long long it = blockIdx.x; it < count; it += gridDim.x) {
        const long long shot = P.idx ? (long long)P.idx[it] : it;
        const K *llr = reinterpret_cast<const K *>(P.llr) + (size_t)shot * n;
        const uint32_t *hard = P.hard + (size_t)shot * WN;
        __syncthreads();
        // ---- stable ascending order of |llr| by rank counting ---------------------------------
        for (int j = tid; j < n; j += NT) keys[j] = KeyBits<K>::get(llr[j]);
        for (int w = tid; w < WM; w += NT) rsw[w] = P.synd[(size_t)shot * WM + w];
        for (int w = tid; w < WN; w += NT) solw[w] = hard[w];
        __syncthreads();
        for (int i = tid; i < n; i += NT) {
            const kbits ki = keys[i];
            int cnt = 0;
#pragma unroll 4
            for (int j = 0; j < n; ++j) cnt += (keys[j] < ki + (kbits)(j < i));
            ord[cnt] = (uint16_t)i;
        }
        // ---- residual syndrome s ^ H*hard; T = I; positions ----------------------------------
        for (int v = tid; v < n; v += NT)
            if ((hard[v >> 5] >> (v & 31)) & 1u)
                for (int a = P.var_ptr[v]; a < P.var_ptr[v + 1]; ++a) {
                    const int c = (int)P.vtab[2 * a + 1];
                    atomicXor(&rsw[c >> 5], 1u << (c & 31));
                }
        for (int t = tid; t < m * WM; t += NT) {
            const int r = t / WM, w = t - r * WM;
            T[t] = (w == (r >> 5)) ? (1u << (r & 31)) : 0u;
        }
        for (int r = tid; r < m; r += NT) { pos[r] = (int16_t)r; pcol[r] = 0xFFFFu; }
        __syncthreads();
        for (int w = tid; w < WM; w += NT) bw[w] = rsw[w];
        __syncthreads();

        // ---- gf2_elimination (OSD.py:31-72) -----------------------------------------------------
        int row = 0;
        const int rank = P.rank;
        for (int j = 0; j < n && row < rank; ++j) {
            const int col = ord[j];
            const int a0 = P.var_ptr[col], a1 = P.var_ptr[col + 1];
            if (tid == 0) s_pmin[j & 1] = 0x7fffffff;     // double-buffered: readers of column j-1 may still be behind
            __syncthreads();
            int best = 0x7fffffff;
            unsigned hasmask = 0;                                   // bit s: my s-th row has the bit
            int sidx = 0;
            int cw[4], cs[4];                                      // word / shift of the first 4 checks of this column
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int c = (a0 + k < a1) ? (int)P.vtab[2 * (a0 + k) + 1] : -1;
                cw[k] = c >> 5;
                cs[k] = (c < 0) ? -1 : (c & 31);
            }
            for (int r = tid; r < m; r += NT, ++sidx) {
                uint32_t x = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (cs[k] >= 0) x ^= T[(size_t)r * WM + cw[k]] >> cs[k];
                for (int a = a0 + 4; a < a1; ++a) {
                    const int c = (int)P.vtab[2 * a + 1];
                    x ^= T[(size_t)r * WM + (c >> 5)] >> (c & 31);
                }
                if (x & 1u) {
                    hasmask |= 1u << sidx;
                    const int p = pos[r];
                    if (p >= row && p < best) best = p;
                }
            }
            best = __reduce_min_sync(0xffffffffu, best);
            if (lane == 0 && best != 0x7fffffff) atomicMin(&s_pmin[j & 1], best);
            __syncthreads();
            const int pmin = s_pmin[j & 1];
            if (pmin == 0x7fffffff) continue;                       // uniform: no pivot in this column
            // locate the pivot row (position pmin) and the row currently at position `row`
            for (int r = tid; r < m; r += NT) if (pos[r] == pmin) s_prow = r;
            __syncthreads();
            const int prow = s_prow;
            const uint32_t pb = (bw[prow >> 5] >> (prow & 31)) & 1u;
            sidx = 0;
            for (int r = tid; r < m; r += NT, ++sidx) {
                if (r == prow) continue;
                if (pos[r] == row) pos[r] = (int16_t)pmin;          // swap positions (OSD.py:56-58)
                if ((hasmask >> sidx) & 1u) {                       // eliminate (OSD.py:64-68)
                    for (int w = 0; w < WM; ++w) T[(size_t)r * WM + w] ^= T[(size_t)prow * WM + w];
                    if (pb) atomicXor(&bw[r >> 5], 1u << (r & 31));
                }
            }
            __syncthreads();
            if (tid == 0) { pos[prow] = (int16_t)row; pcol[prow] = (uint16_t)j; }
            ++row;
        }
        __syncthreads();

        // ---- solution and validity -----------------------------------------------------------------
        int bad = 0;
        for (int r = tid; r < m; r += NT) {
            const uint32_t b = (bw[r >> 5] >> (r & 31)) & 1u;
            if (pcol[r] != 0xFFFFu) {
                if (b) { const int v = ord[pcol[r]]; atomicXor(&solw[v >> 5], 1u << (v & 31)); }
            } else if (b) {
                bad = 1;
            }
        }
        bad = __syncthreads_or(bad);
        for (int w = tid; w < WN; w += NT) P.out[(size_t)shot * WN + w] = solw[w];
        if (tid == 0 && P.valid) P.valid[shot] = bad ? 0 : 1;
    }
}

// ------------------------------------------------------------------------------------------------
// OSD-0 for large check matrices, column-major transform, forward elimination in batches: the production block kernel.
//
// Same observation as osd0_fast_kernel: the OSD-0 solution does not depend on which row pivots a column, and a column
// with no free row never gets one back.  The transform is stored by COLUMNS, an m-bit vector of WM words each: the reduced
// column j = T h_j is the XOR of the <= 3 columns of T that belong to the checks of variable ordering[j]; a row operation
// "rows S ^= row p" becomes "every column of T with bit p set ^= S".  Only FREE rows are eliminated (S = free rows of the
// pivot column): pivot rows are frozen once chosen, T fills in like L^-1 instead of B^-1, and the solution follows from a
// back-substitution over the pivots in reverse order.  Row c of T is added to other rows only once c is a pivot row, so the
// column of a free row is the unit vector and is never stored: slot a holds the column of the a-th pivot row (rowpiv[c] =
// its slot, or the zero slot m+1 while c is free), slot m the syndrome column.
//
// One CTA of 512 threads per shot; a round takes OSDB_BATCH = 32 consecutive candidate columns:
//   evaluate  lane = candidate; warp w computes the words w, w + 16 of all 32 reduced columns (three loads per word serve 32
//             candidates) and, per word, the bits that exactly one candidate has (prefix-OR scan over the lanes + two warp
//             reductions); every candidate's lowest exclusive bit / lowest bit is left in shared memory (atomicMin);
//   resolve   warp 0, lane k = candidate k.  A candidate with an exclusive bit pivots on it: nobody else has that row, so it is
//             neither reduced by anybody nor changed when the others pivot.  For the rest the 32 x 32 interaction matrix
//             I[k'][k] = "candidate k' has the pivot row of candidate k" is gathered once; reducing k' by k is
//             row[k'] ^= row[k] on it, the vectors are touched only to carry the XOR out, and a candidate whose own bit
//             disappears is dependent or takes a new pivot row.  Entries above the diagonal are folded into the vectors
//             afterwards, in descending order, S'_b = S_b ^ sum_{a > b, S_b[p_a]} S'_a, which turns the SEQUENCE of row
//             operations of the round into ONE linear map  c -> c ^ sum_a c[p_a] S'_a  whose coefficients are bits of the
//             column as it stood BEFORE the round;
//   apply     warp 0 publishes the pivot ROWS as soon as the forward reduction is done (bar.arrive on a named barrier) and folds
//             while warps 1..15 gather: a lane collects the <= 32 coefficient bits of its stored column (consecutive slots:
//             conflict-free; independent loads, no chain through the pivots); the columns that have any (~65 per round) go to a
//             shared list.  After a block barrier the XORs of the selected S' vectors into those columns (a word per lane) are
//             dealt out evenly over all 16 warps -- the hits cluster -- and the columns of the new pivot rows (unit vector ^ S')
//             are appended with lane = pivot.
// Four block barriers and one named barrier per 32 candidates (the round-1 kernel spent three per 8 and walked every column
// through the pivots of the batch one after the other: 1.2 M of its 2.0 M cycles per shot).  2.4e5 -> 1.17e6 failed shots/s on
// 864 x 2592 (with the radix sort of the keys: 1.21e6).
// Inconsistent syndromes are handed to osd0_block_kernel (redo list).  The checks of a column come from a per-code table
// packed 3 x 10 bits (m <= 1024, column weight <= 3: the space-time matrices; otherwise from the CSC in global memory).
// ------------------------------------------------------------------------------------------------
constexpr int OSDB_BATCH = 32;
#ifndef OSDBF_THREADS_N
#define OSDBF_THREADS_N 512
#endif
constexpr int OSDBF_THREADS = OSDBF_THREADS_N;

// the sort runs inside the (still idle) transform area: a radix sort when its buffers fit there (the space-time matrices), else a
// bitonic sort of the padded length if THAT fits; only the rank-counting fallback needs the keys outside of it.
template <typename K>
__host__ __device__ inline bool osdbf_bitonic_fits(int m, int n)
{
    size_t N2 = 1;
    while (N2 < (size_t)n) N2 <<= 1;
    return (sizeof(typename KeyBits<K>::type) + 2) * N2 <= 4 * (size_t)m * ((m + 31) / 32);
}
template <typename K>
__host__ __device__ inline size_t osdbf_key_area(int m, int n)
{
    return osdbf_bitonic_fits<K>(m, n) ? 0 : sizeof(typename KeyBits<K>::type) * (size_t)n;
}

template <typename K>
__host__ __device__ inline size_t osdbf_smem_bytes(int m, int n)
{
    const int WM = (m + 31) / 32, WN = (n + 31) / 32;
    size_t o = 4 * (size_t)(m + 2) * WM;                        // TC, syndrome column, zero slot
    o += 2 * (size_t)m * 3;                                     // pivot rows / sorted positions of the pivot columns / pivot index of a row (uint16)
    o = (o + 3) & ~(size_t)3;
    o += 4 * (size_t)WM;                                        // used
    o += 4 * (size_t)WM * OSDB_BATCH;                           // candidate columns of the round
    o += 4 * (size_t)WN;                                        // solution words
    o = (o + 7) & ~(size_t)7;
    o += osdbf_key_area<K>(m, n);                               // keys (rank-counting path only)
    o += 2 * (size_t)n;                                         // ordering (uint16)
    o = (o + 3) & ~(size_t)3;
    o += 6 * (size_t)(m + 1);                                   // list of the columns that take part in a round: coefficient bits, slot
    return o + 64;
}


// (key, index) pairs of the register-resident bitonic sort that double keys still use (eight radix passes are slower than it)
template <typename KBt> struct OsdSortElem;
template <> struct OsdSortElem<unsigned long long> {
    unsigned long long k;
    uint32_t i;
    static constexpr int BYTES = 10;
    __device__ __forceinline__ static OsdSortElem make(unsigned long long key, uint32_t idx) { return OsdSortElem{key, idx}; }
    __device__ __forceinline__ bool after(const OsdSortElem &o) const { return k > o.k || (k == o.k && i > o.i); }
    __device__ __forceinline__ OsdSortElem shfl_xor(int mask) const { return OsdSortElem{__shfl_xor_sync(0xffffffffu, k, mask), __shfl_xor_sync(0xffffffffu, i, mask)}; }
    __device__ __forceinline__ uint32_t index() const { return i; }
    __device__ __forceinline__ void store(unsigned char *base, int N2, int buf, int pos) const
    {
        reinterpret_cast<unsigned long long *>(base)[buf * N2 + pos] = k;
        reinterpret_cast<uint16_t *>(base + 16 * (size_t)N2)[buf * N2 + pos] = (uint16_t)i;
    }
    __device__ __forceinline__ static OsdSortElem load(const unsigned char *base, int N2, int buf, int pos)
    {
        return OsdSortElem{reinterpret_cast<const unsigned long long *>(base)[buf * N2 + pos], reinterpret_cast<const uint16_t *>(base + 16 * (size_t)N2)[buf * N2 + pos]};
    }
};
constexpr int OSDBF_EPT = 8;                 // elements per thread of the register-resident sort

template <int N> struct OsdIC { static constexpr int value = N; };

template <typename K, bool PACKED, int WMT>
__global__ void __launch_bounds__(OSDBF_THREADS, 2) osd0_block_fast_kernel(const OSDBlockParams P)
{
    typedef typename KeyBits<K>::type kbits;
    constexpr int NW = OSDBF_THREADS / 32, NWW = NW - 1;         // warps; warps that gather while warp 0 folds
    constexpr int KB = OSDB_BATCH, CPW = KB / NW;               // candidates per round / per warp
    static_assert(KB == 32 && KB % NW == 0, "one candidate per lane of the resolving warp");
    const int m = P.m, n = P.n, WM = WMT ? WMT : P.WM, WN = P.WN;       // WMT: the word count at compile time (loops over words unroll), 0: any
    const int tid = threadIdx.x, NT = OSDBF_THREADS, lane = tid & 31, warp = tid >> 5;
    const unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char smem[];
    uint32_t *TCP = reinterpret_cast<uint32_t *>(smem);                     // [m + 2][WM]  slot a: column prow[a] of T; slot m: syndrome column; slot m+1: zeros
    uint32_t *bw = TCP + (size_t)m * WM;
    uint16_t *prow = reinterpret_cast<uint16_t *>(TCP + (size_t)(m + 2) * WM);  // [m] pivot row of the k-th pivot
    uint16_t *pcolj = prow + m;                                            // [m] sorted position of the k-th pivot column
    uint16_t *rowpiv = pcolj + m;                                          // [m] slot of the column of row r: its pivot index, m+1 (the zero slot) while free
    uint32_t *used = reinterpret_cast<uint32_t *>((reinterpret_cast<uintptr_t>(rowpiv + m) + 3) & ~(uintptr_t)3);   // [WM]
    uint32_t *cand = used + WM;                                             // [KB][WM] free rows of the candidates
    uint32_t *solw = cand + (size_t)KB * WM;                                // [WN]
    kbits *keys = reinterpret_cast<kbits *>((reinterpret_cast<uintptr_t>(solw + WN) + 7) & ~(uintptr_t)7);
    uint16_t *ord = reinterpret_cast<uint16_t *>(reinterpret_cast<unsigned char *>(keys) + osdbf_key_area<K>(m, n));
    uint32_t *s_hx = reinterpret_cast<uint32_t *>((reinterpret_cast<uintptr_t>(ord + n) + 3) & ~(uintptr_t)3);   // [m + 1] coefficient bits of the columns that take part in the round
    uint16_t *s_hslot = reinterpret_cast<uint16_t *>(s_hx + (m + 1));                                             // [m + 1] and their slots
    __shared__ int s_nhit;
    __shared__ uint32_t s_pl[KB];                      // pivot rows accepted in this round, in order
    __shared__ int s_off[KB];                          // and where their S' vectors are (offset into cand)
    __shared__ int s_pex[KB], s_pany[KB];              // per candidate: lowest free row that no other candidate of the round has / lowest free row (INT_MAX: none)
    __shared__ uint2 s_g[KB];                          // and {word, 31 - bit} of the pivot row, as the coefficient gather wants them
    __shared__ int s_nacc;
    constexpr bool packed_chk = PACKED;               // m <= 1024 and column weight <= 3 (checked by the host)

    // Column c of T = the stored column of its slot if row c is a pivot row, the unit vector e_c otherwise.  Free rows point at the
    // zero slot, so "stored ^ e_c" is right on the FREE rows for every c (a stored column holds its own unit bit, which e_c cancels --
    // on a pivot row, masked off by the caller), and "stored" alone is right on the pivot rows (back-substitution).
    auto tcol_free = [&](int c, int w) -> uint32_t { return TCP[(size_t)rowpiv[c] * WM + w] ^ ((c >> 5) == w ? (1u << (c & 31)) : 0u); };
    auto tcol_piv = [&](int c, int w) -> uint32_t { return TCP[(size_t)rowpiv[c] * WM + w]; };
    // reduced column of sorted position jj on the pivot rows (XOR of the T columns of its checks), word w
    auto reduced_piv = [&](int jj, int w) -> uint32_t {
        const int col = ord[jj];
        uint32_t x = 0;
        if (packed_chk) {
            const uint32_t e = __ldg(P.colpack + col);
            const int cnt = (int)(e >> 30);
            if (cnt > 0) x = tcol_piv((int)(e & 1023u), w);
            if (cnt > 1) x ^= tcol_piv((int)((e >> 10) & 1023u), w);
            if (cnt > 2) x ^= tcol_piv((int)((e >> 20) & 1023u), w);
        } else {
            for (int a = P.var_ptr[col]; a < P.var_ptr[col + 1]; ++a) x ^= tcol_piv((int)P.vtab[2 * a + 1], w);
        }
        return x;
    };
    // packed check list of candidate `lane` of the round that starts at sorted position jb
    auto fetch = [&](int jb) -> uint32_t {
        const int jj = jb + lane;
        return (packed_chk && jj < n) ? __ldg(P.colpack + ord[jj]) : 0u;
    };

    const long long count = P.count_dev ? (long long)*P.count_dev : P.count_host;
    for (long long it = blockIdx.x; it < count; it += gridDim.x) {
        const long long shot = P.idx ? (long long)P.idx[it] : it;
        const K *llr = reinterpret_cast<const K *>(P.llr) + (size_t)shot * n;
        const uint32_t *hard = P.hard + (size_t)shot * WN;
        __syncthreads();
        // ---- stable ascending order of |llr| ----------------------------------------------------
        for (int w = tid; w < WM; w += NT) { bw[w] = P.synd[(size_t)shot * WM + w]; used[w] = 0; }
        for (int w = tid; w < WN; w += NT) solw[w] = hard[w];
        int N2 = 1;
        while (N2 < n) N2 <<= 1;
        constexpr int RE = 8;                                          // elements per thread of the radix sort, at most
        const int E = (n + NT - 1) / NT, NTOT = NT * E;
        if (sizeof(kbits) == 4 && E <= RE && (size_t)NTOT * (sizeof(kbits) + 2) + NW * 256 * 2 + 256 * 4 + 16 <= 4 * (size_t)m * WM) {
            // float keys: stable LSD radix sort of (key, index), 8 bits per pass, in the still unused transform area.  Thread (warp, lane) holds the elements at positions 32 E warp + 32 e + lane, e < E: within a warp the
            // rank of an element among the equal digits before it is popc(match_any & lanes below) plus a running per-warp count per
            // digit; a column scan over the warps and a 256-wide scan of the digit totals give the destinations.  ~200 instructions
            // per thread and pass (the bitonic network it replaces: 78 stages, ~4000 instructions per thread).
            unsigned char *sb = reinterpret_cast<unsigned char *>(TCP);
            kbits *kb = reinterpret_cast<kbits *>(sb);                                     // [NTOT]
            uint16_t *ib = reinterpret_cast<uint16_t *>(kb + NTOT);                        // [NTOT]
            uint16_t *hist = ib + NTOT;                                                    // [NW][256]
            uint32_t *base = reinterpret_cast<uint32_t *>((reinterpret_cast<uintptr_t>(hist + NW * 256) + 3) & ~(uintptr_t)3);   // [256]
            __shared__ uint32_t s_wsum[8];
            kbits key[RE];
            uint32_t idx[RE];
            const int p0 = warp * 32 * E + lane;
#pragma unroll
            for (int e = 0; e < RE; ++e)
                if (e < E) {
                    const int pos = p0 + 32 * e;
                    key[e] = (pos < n) ? KeyBits<K>::get(llr[pos]) : ~(kbits)0;           // padding sorts last
                    idx[e] = (pos < n) ? (uint32_t)pos : 0xFFFFu;
                }
            for (int pass = 0; pass < (int)sizeof(kbits); ++pass) {
                const int shift = 8 * pass;
                for (int i = tid; i < NW * 128; i += NT) reinterpret_cast<uint32_t *>(hist)[i] = 0u;
                __syncthreads();
                uint32_t rank[RE];
#pragma unroll
                for (int e = 0; e < RE; ++e)
                    if (e < E) {                                                           // (uniform)
                        const uint32_t d = (uint32_t)(key[e] >> shift) & 255u;
                        const unsigned peers = __match_any_sync(FULL, d);
                        const uint32_t r = __popc(peers & ((1u << lane) - 1u));
                        const uint32_t prev = hist[warp * 256 + d];
                        __syncwarp();
                        if (r == 0) hist[warp * 256 + d] = (uint16_t)(prev + __popc(peers));
                        __syncwarp();
                        rank[e] = prev + r;
                    }
                __syncthreads();
                uint32_t tot = 0, incl = 0;
                if (tid < 256) {
                    for (int w = 0; w < NW; ++w) {                                         // per digit: exclusive scan over the warps
                        const uint32_t c = hist[w * 256 + tid];
                        hist[w * 256 + tid] = (uint16_t)tot;
                        tot += c;
                    }
                    incl = tot;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const uint32_t o = __shfl_up_sync(FULL, incl, d);
                        if (lane >= d) incl += o;
                    }
                    if (lane == 31) s_wsum[warp] = incl;
                }
                __syncthreads();
                if (tid < 256) {
                    uint32_t off = 0;
                    for (int w = 0; w < warp; ++w) off += s_wsum[w];
                    base[tid] = off + incl - tot;
                }
                __syncthreads();
#pragma unroll
                for (int e = 0; e < RE; ++e)
                    if (e < E) {
                        const uint32_t d = (uint32_t)(key[e] >> shift) & 255u;
                        const uint32_t dst = base[d] + hist[warp * 256 + d] + rank[e];
                        kb[dst] = key[e];
                        ib[dst] = (uint16_t)idx[e];
                    }
                __syncthreads();
#pragma unroll
                for (int e = 0; e < RE; ++e)
                    if (e < E) {
                        key[e] = kb[p0 + 32 * e];
                        idx[e] = ib[p0 + 32 * e];
                    }
            }
#pragma unroll
            for (int e = 0; e < RE; ++e)
                if (e < E && p0 + 32 * e < n) ord[p0 + 32 * e] = (uint16_t)idx[e];
        } else if (sizeof(kbits) == 8 && N2 == OSDBF_EPT * NT && (size_t)2 * N2 * 10 <= 4 * (size_t)m * WM) {
            // bitonic sort with the elements in registers: thread (warp, lane) holds the elements 256 warp + 32 e + lane, e < 8.
            // Partners at distance < 32 come by shuffle, at 32 / 64 / 128 from the thread's own registers; only the 10 stages at
            // distance >= 256 (of 78) go through shared memory (double-buffered in the still unused transform area: one barrier each).
            typedef OsdSortElem<unsigned long long> El;
            constexpr int EPT = OSDBF_EPT;
            El el[EPT];
            const int bi = warp * (32 * EPT) + lane;
#pragma unroll
            for (int e = 0; e < EPT; ++e) {
                const int i = bi + 32 * e;
                el[e] = (i < n) ? El::make((unsigned long long)KeyBits<K>::get(llr[i]), (uint32_t)i) : El::make(~0ull, 0xFFFFu);
            }
            // element i takes its partner's value iff (mine > partner) ^ (i is the upper one of the pair) ^ (i lies in a descending block)
            auto reg_stage = [&](auto d_c, unsigned descm) {
                constexpr int D = decltype(d_c)::value;
#pragma unroll
                for (int e = 0; e < EPT; ++e)
                    if ((e & D) == 0) {
                        if (el[e].after(el[e | D]) != (bool)((descm >> e) & 1u)) { const El t = el[e]; el[e] = el[e | D]; el[e | D] = t; }
                    }
            };
            unsigned char *sbase = reinterpret_cast<unsigned char *>(TCP);
            int buf = 0;
            for (int k = 2; k <= N2; k <<= 1) {
                // bit e of descm: element e of this thread lies in a descending block of length k
                unsigned descm = 0;
#pragma unroll
                for (int e = 0; e < EPT; ++e) descm |= (((bi + 32 * e) & k) != 0 ? 1u : 0u) << e;
                for (int jd = k >> 1; jd > 0; jd >>= 1) {
                    if (jd >= 32 * EPT) {
#pragma unroll
                        for (int e = 0; e < EPT; ++e) el[e].store(sbase, N2, buf, bi + 32 * e);
                        __syncthreads();
                        const bool hi = (bi & jd) != 0;                                  // (a warp bit)
#pragma unroll
                        for (int e = 0; e < EPT; ++e) {
                            const El o = El::load(sbase, N2, buf, (bi + 32 * e) ^ jd);
                            if (el[e].after(o) != (hi != (bool)((descm >> e) & 1u))) el[e] = o;
                        }
                        buf ^= 1;
                    } else if (jd == 128) reg_stage(OsdIC<4>(), descm);
                    else if (jd == 64) reg_stage(OsdIC<2>(), descm);
                    else if (jd == 32) reg_stage(OsdIC<1>(), descm);
                    else {
                        const bool hi = (lane & jd) != 0;
#pragma unroll
                        for (int e = 0; e < EPT; ++e) {
                            const El o = el[e].shfl_xor(jd);
                            if (el[e].after(o) != (hi != (bool)((descm >> e) & 1u))) el[e] = o;
                        }
                    }
                }
            }
#pragma unroll
            for (int e = 0; e < EPT; ++e)
                if (bi + 32 * e < n) ord[bi + 32 * e] = (uint16_t)el[e].index();
        } else if (osdbf_bitonic_fits<K>(m, n)) {
            // bitonic sort of (key, index) pairs in the (still unused) transform area: O(n log^2 n) instead of the
            // O(n^2) rank counting -- 78 stages of 2048 compare-exchanges for n = 2592
            kbits *sk = reinterpret_cast<kbits *>(TCP);
            uint16_t *si = reinterpret_cast<uint16_t *>(sk + N2);
            for (int j = tid; j < N2; j += NT) {
                sk[j] = (j < n) ? KeyBits<K>::get(llr[j]) : ~(kbits)0;
                si[j] = (uint16_t)((j < n) ? j : 0xFFFF);
            }
            __syncthreads();
            for (int k = 2; k <= N2; k <<= 1)
                for (int jd = k >> 1; jd > 0; jd >>= 1) {
                    for (int t = tid; t < (N2 >> 1); t += NT) {
                        const int i = ((t & ~(jd - 1)) << 1) | (t & (jd - 1));       // element with bit jd clear
                        const int l = i | jd;
                        const kbits ka = sk[i], kb = sk[l];
                        const uint16_t ia = si[i], ib = si[l];
                        const bool a_after_b = (ka > kb) || (ka == kb && ia > ib);
                        const bool up = (i & k) == 0;                                // ascending block
                        if (a_after_b == up) { sk[i] = kb; sk[l] = ka; si[i] = ib; si[l] = ia; }
                    }
                    __syncthreads();
                }
            for (int j = tid; j < n; j += NT) ord[j] = si[j];
        } else {
            for (int j = tid; j < n; j += NT) keys[j] = KeyBits<K>::get(llr[j]);
            __syncthreads();
            for (int i = tid; i < n; i += NT) {
                const kbits ki = keys[i];
                int cnt = 0;
                int j = 0;
#pragma unroll 4
                for (; j < i; ++j) osd_count_lt(cnt, keys[j], ki + (kbits)1);        // j < i: key_j <= key_i
#pragma unroll 4
                for (; j < n; ++j) osd_count_lt(cnt, keys[j], ki);
                ord[-cnt] = (uint16_t)i;
            }
        }
        __syncthreads();                                              // keys are dead from here on
        // ---- residual syndrome s ^ H*hard; T = I (no column stored yet) ---------------------------
        for (int v = tid; v < n; v += NT)
            if ((hard[v >> 5] >> (v & 31)) & 1u)
                for (int a = P.var_ptr[v]; a < P.var_ptr[v + 1]; ++a) {
                    const int c = (int)P.vtab[2 * a + 1];
                    atomicXor(&bw[c >> 5], 1u << (c & 31));
                }
        for (int r = tid; r < m; r += NT) rowpiv[r] = (uint16_t)(m + 1);
        if (tid < KB) { s_pex[tid] = 0x7fffffff; s_pany[tid] = 0x7fffffff; }
        for (int w = tid; w < WM; w += NT) TCP[(size_t)(m + 1) * WM + w] = 0u;
        uint32_t e_cur = fetch(0);
        __syncthreads();

        // ---- forward elimination, KB candidate columns per round ------------------------------------
        int j = 0, npiv = 0;
        const int rank = P.rank;
        while (j < n && npiv < rank) {
            // evaluate: lane = candidate, warp w takes the words w, w + NW, ... of all 32 (the check lists were fetched a round ahead).
            // 3 loads per word and warp serve 32 candidates (bank conflicts among the lanes' slots included, a fraction of the
            // instructions of one warp per candidate).
            {
                const uint32_t e_nx = fetch(j + KB);
                // word w of candidate `lane` is x: store it, and note the candidate's lowest bit that no other candidate of the round has
                // (bits that occur twice = OR over the lanes of x & (OR of the lanes before): one prefix-OR scan and two warp reductions)
                auto finish_word = [&](int w, uint32_t x) {
                    cand[(size_t)lane * WM + w] = x;
                    uint32_t pre = x;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const uint32_t o = __shfl_up_sync(FULL, pre, d);
                        if (lane >= d) pre |= o;
                    }
                    uint32_t before = __shfl_up_sync(FULL, pre, 1);
                    if (lane == 0) before = 0;
                    const uint32_t twice = __reduce_or_sync(FULL, x & before);
                    if (x) {
                        atomicMin(&s_pany[lane], 32 * w + __ffs(x) - 1);
                        const uint32_t u = x & ~twice;
                        if (u) atomicMin(&s_pex[lane], 32 * w + __ffs(u) - 1);
                    }
                };
                if (packed_chk) {
                    const uint32_t e = e_cur;
                    const int cnt = (int)(e >> 30);
                    int base[3], uw[3];
                    uint32_t ub[3];
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        const int c = (int)((e >> (10 * i)) & 1023u);
                        const bool on = cnt > i;
                        base[i] = (on ? (int)rowpiv[c] : m + 1) * WM;
                        uw[i] = on ? (c >> 5) : -1;
                        ub[i] = 1u << (c & 31);
                    }
                    for (int w = warp; w < WM; w += NW) {
                        uint32_t x = TCP[base[0] + w] ^ TCP[base[1] + w] ^ TCP[base[2] + w];
                        x ^= (uw[0] == w ? ub[0] : 0u) ^ (uw[1] == w ? ub[1] : 0u) ^ (uw[2] == w ? ub[2] : 0u);
                        finish_word(w, x & ~used[w]);
                    }
                } else {
                    const int jj = j + lane;
                    const int col = jj < n ? (int)ord[jj] : 0;
                    const int a0 = jj < n ? P.var_ptr[col] : 0, a1 = jj < n ? P.var_ptr[col + 1] : 0;
                    for (int w = warp; w < WM; w += NW) {
                        uint32_t x = 0;
                        for (int a = a0; a < a1; ++a) x ^= tcol_free((int)P.vtab[2 * a + 1], w);
                        finish_word(w, x & ~used[w]);
                    }
                }
                e_cur = e_nx;
            }
            __syncthreads();
            if (warp == 0) {
                // lane = candidate.  Pivot row: its lowest free row that no other candidate of the round has (found during the
                // evaluation) -- such a candidate neither has to be brought to anyone else nor changes when others pivot -- else its
                // lowest free row
                const int pex = s_pex[lane], pany = s_pany[lane];
                s_pex[lane] = 0x7fffffff;
                s_pany[lane] = 0x7fffffff;
                const bool excl = pex != 0x7fffffff;
                int p = excl ? pex : (pany != 0x7fffffff ? pany : -1);
                unsigned acc = __ballot_sync(FULL, p >= 0);
                const unsigned shared_piv = __ballot_sync(FULL, p >= 0 && !excl);
                const unsigned below = (1u << lane) - 1u;
                // interaction matrix: bit k of im = this candidate has the pivot row of candidate k (shared pivot rows only: an
                // exclusive one is in nobody else), bit `lane` = it has its own.  Reducing candidate k' by candidate k is
                // im[k'] ^= im[k] on this matrix -- the vectors themselves are only touched to carry the XOR out.
                unsigned im = 0;
                for (unsigned t = shared_piv; t;) {
                    int kk[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) { kk[u] = t ? __ffs(t) - 1 : -1; t &= t - 1; }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int pk = __shfl_sync(FULL, p, kk[u] < 0 ? 0 : kk[u]);
                        if (kk[u] >= 0) im |= ((cand[(size_t)lane * WM + (pk >> 5)] >> (pk & 31)) & 1u) << kk[u];
                    }
                }
                if (p >= 0) im |= 1u << lane;
                // forward: a candidate that has the pivot row of an accepted candidate before it is reduced by it first
                unsigned todo = __ballot_sync(FULL, (im & below & acc) != 0);
                while (todo) {
                    const int kp = __ffs(todo) - 1;
                    todo &= todo - 1;
                    const unsigned lowk = (1u << kp) - 1u;
                    unsigned row = __shfl_sync(FULL, im, kp);
                    unsigned imk = row & acc & lowk;
                    if (!imk) continue;
                    uint32_t v = (lane < WM) ? cand[(size_t)kp * WM + lane] : 0u;
                    do {
                        const int k = __ffs(imk) - 1;                 // rows of accepted candidates have nothing below their own bit
                        if (lane < WM) v ^= cand[(size_t)k * WM + lane];
                        row ^= __shfl_sync(FULL, im, k);
                        imk = row & acc & lowk;
                    } while (imk);
                    if (lane < WM) cand[(size_t)kp * WM + lane] = v;
                    if ((row >> kp) & 1u) {                            // it still has its pivot row
                        if (lane == kp) im = row;
                        __syncwarp();
                        continue;
                    }
                    const unsigned bal = __ballot_sync(FULL, v != 0);
                    if (!bal) {                                        // dependent on the candidates before it
                        acc &= ~(1u << kp);
                        if (lane == kp) { p = -1; im = 0; }
                        __syncwarp();
                        continue;
                    }
                    const int src = __ffs(bal) - 1;
                    const int newp = 32 * src + __ffs(__shfl_sync(FULL, v, src)) - 1;
                    __syncwarp();
                    // a new pivot row: column kp of the matrix changes
                    const uint32_t b = (cand[(size_t)lane * WM + (newp >> 5)] >> (newp & 31)) & 1u;
                    if (lane == kp) { p = newp; im = row | (1u << kp); }
                    else im = (im & ~(1u << kp)) | (b << kp);
                    todo |= __ballot_sync(FULL, b != 0 && lane > kp);
                }
                // never more pivots than the rank of H
                while (__popc(acc) > rank - npiv) acc &= ~(0x80000000u >> __clz(acc));
                const bool mine = (acc >> lane) & 1u;
                // publish the pivots: the other warps start gathering the coefficient bits of the stored columns (which needs the
                // pivot ROWS only) while this warp still folds the S vectors
                if (mine) {
                    const int a = __popc(acc & below);
                    s_pl[a] = (uint32_t)p;
                    s_g[a] = make_uint2((unsigned)(p >> 5), (unsigned)(31 - (p & 31)));
                    s_off[a] = lane * WM;
                    prow[npiv + a] = (uint16_t)p;
                    pcolj[npiv + a] = (uint16_t)(j + lane);
                    rowpiv[p] = (uint16_t)(npiv + a);
                    atomicOr(&used[p >> 5], 1u << (p & 31));
                }
                if (lane == 0) { s_nacc = __popc(acc); s_nhit = 0; }
                __threadfence_block();
                asm volatile("bar.arrive 1, %0;" ::"n"(NW * 32) : "memory");
                // S = free rows without the pivot row; fold the entries above the diagonal in (descending order)
                if (mine) cand[(size_t)lane * WM + (p >> 5)] &= ~(1u << (p & 31));
                __syncwarp();
                const unsigned jm = mine ? (im & acc & ~((2u << lane) - 1u)) : 0u;
                unsigned bt = __ballot_sync(FULL, jm != 0);
                while (bt) {
                    const int b = 31 - __clz(bt);
                    bt &= ~(1u << b);
                    unsigned jb = __shfl_sync(FULL, jm, b);
                    uint32_t v = 0;
                    while (jb) {
                        const int a = __ffs(jb) - 1;
                        jb &= jb - 1;
                        if (lane < WM) v ^= cand[(size_t)a * WM + lane];
                    }
                    if (lane < WM) cand[(size_t)b * WM + lane] ^= v;
                    __syncwarp();
                }
            } else {
                asm volatile("bar.sync 1, %0;" ::"n"(NW * 32) : "memory");      // the pivot rows of the round are known
            }
            const int nacc = s_nacc;
            // apply.  Row c of T is added to other rows only once c is a pivot row: the columns of free rows are unit vectors
            // and have no pivot row of this round, so only the stored columns (pivots 0 .. npiv-1) and the syndrome take part.
            // Column ^= sum over the accepted candidates a whose pivot row the column has (as it stands now) of S'_a.
            auto apply_group = [&](auto ni_c, int g0) {
                constexpr int NI = decltype(ni_c)::value;
                unsigned x[NI];
                int cb[NI];
#pragma unroll
                for (int i = 0; i < NI; ++i) {
                    const int idx = g0 + i * NWW * 32 + lane;
                    cb[i] = (idx < npiv ? idx : m) * WM;
                    x[i] = 0;
                }
#pragma unroll 4
                for (int a = 0; a < nacc; ++a) {                     // bit nacc-1-a of x: pivot a
                    const uint2 g = s_g[a];
#pragma unroll
                    for (int i = 0; i < NI; ++i) x[i] = __funnelshift_l(TCP[cb[i] + g.x] << g.y, x[i], 1);
                }
#pragma unroll
                for (int i = 0; i < NI; ++i) {
                    const int c0 = g0 + i * NWW * 32;
                    // the columns that take part go to a list; the XORs are dealt out evenly over the warps afterwards (the hits
                    // cluster: handled where they are found, the warp with the most of them held everybody up)
                    const bool h = x[i] != 0 && c0 + lane <= npiv;
                    const unsigned hit = __ballot_sync(FULL, h);
                    if (hit) {
                        int base = 0;
                        if (lane == 0) base = atomicAdd(&s_nhit, __popc(hit));
                        base = __shfl_sync(FULL, base, 0);
                        if (h) {
                            const int pos = base + __popc(hit & ((1u << lane) - 1u));
                            s_hslot[pos] = (uint16_t)(c0 + lane < npiv ? c0 + lane : m);
                            s_hx[pos] = x[i];
                        }
                    }
                }
            };
            for (int g0 = (warp - 1) * 32; warp > 0 && g0 <= npiv; g0 += 4 * NWW * 32) {
                const int ni = (npiv + 1 - g0 + NWW * 32 - 1) / (NWW * 32);
                if (ni >= 4) apply_group(OsdIC<4>(), g0);
                else if (ni == 3) apply_group(OsdIC<3>(), g0);
                else if (ni == 2) apply_group(OsdIC<2>(), g0);
                else apply_group(OsdIC<1>(), g0);
            }
            __syncthreads();                                          // warp 0 has folded: the S' vectors are final
            // the columns of the new pivot rows: unit vector ^ S' (lane = pivot of the round, warp w writes the words w, w + NW, ...)
            if (lane < nacc) {
                const uint32_t p = s_pl[lane];
                const int off = s_off[lane];
                for (int w = warp; w < WM; w += NW)
                    TCP[(size_t)(npiv + lane) * WM + w] = cand[off + w] | (w == (int)(p >> 5) ? (1u << (p & 31)) : 0u);
            }
            {
                const int nhit = s_nhit;
                for (int hI = warp; hI < nhit; hI += 2 * NW) {         // two columns in flight per warp
                    const bool two = hI + NW < nhit;
                    unsigned xa = s_hx[hI], xb = two ? s_hx[hI + NW] : 0u;
                    const int sa = s_hslot[hI], sb = two ? (int)s_hslot[hI + NW] : sa;
                    if (lane < WM) {
                        // most columns have a single pivot row of the round: first term of both, then the rest
                        uint32_t va = cand[s_off[nacc - __ffs(xa)] + lane];
                        uint32_t vb = two ? cand[s_off[nacc - __ffs(xb)] + lane] : 0u;
                        uint32_t ta = TCP[(size_t)sa * WM + lane], tb = TCP[(size_t)sb * WM + lane];
                        xa &= xa - 1;
                        xb &= xb - 1;
                        while (xa) { va ^= cand[s_off[nacc - __ffs(xa)] + lane]; xa &= xa - 1; }
                        while (xb) { vb ^= cand[s_off[nacc - __ffs(xb)] + lane]; xb &= xb - 1; }
                        TCP[(size_t)sa * WM + lane] = ta ^ va;
                        if (two) TCP[(size_t)sb * WM + lane] = tb ^ vb;
                    }
                }
            }
            npiv += nacc;
            j += KB;
            __syncthreads();
        }

        // ---- validity; back-substitution over the pivots in reverse order ---------------------------
        int bad = 0;
        for (int w = tid; w < WM; w += NT) {
            const int rows = m - 32 * w;
            const uint32_t live = rows >= 32 ? 0xffffffffu : ((1u << rows) - 1u);
            if (bw[w] & ~used[w] & live) bad = 1;
        }
        bad = __syncthreads_or(bad);
        if (bad) {                                                   // inconsistent syndrome: redo with the reference's pivot rule
            if (tid == 0) P.redo_idx[atomicAdd(P.redo_count, 1u)] = (int32_t)shot;
            continue;
        }
        if (warp == 0) {
            // x_k = b[p_k]; if set, b ^= reduced column j_k (its entries in the rows of the earlier pivots).  Every bit of
            // b is final once all later pivots have been applied, so the next pivot to fire is the LATEST pivot among
            // the set bits of b that belong to pivots before the current one: found with one pass over the set bits (a
            // word per lane) and a warp maximum -- as many steps as the solution has pivots, not as H has rows.
            uint32_t bword = (lane < WM) ? (bw[lane] & used[lane]) : 0u;
            int kcur = npiv;                                          // pivots >= kcur are done
            while (true) {
                int best = -1;
                for (uint32_t t = bword; t; t &= t - 1) {
                    const int kk = (int)rowpiv[32 * lane + __ffs(t) - 1];
                    if (kk < kcur && kk > best) best = kk;
                }
                best = __reduce_max_sync(FULL, best);
                if (best < 0) break;
                const int r = prow[best], jk = pcolj[best];
                if (lane < WM) {
                    uint32_t x = reduced_piv(jk, lane) & used[lane];
                    if (lane == (r >> 5)) x &= ~(1u << (r & 31));    // keep x_k itself
                    bword ^= x;
                }
                if (lane == 0) { const int v = ord[jk]; solw[v >> 5] ^= 1u << (v & 31); }
                kcur = best;
            }
        }
        __syncthreads();
        for (int w = tid; w < WN; w += NT) P.out[(size_t)shot * WN + w] = solw[w];
        if (tid == 0 && P.valid) P.valid[shot] = 1;
    }
}

}  // namespace qldpc
