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
    const uint32_t *synd;         // [B][WM]
    const void *llr;              // [B][n] float or double
    const uint32_t *hard;         // [B][WN]  BP hard decision
    uint32_t *out;                // [B][WN]  solution (may alias hard)
    uint8_t *valid;               // [B] solution satisfies the syndrome (may be null)
    // elimination record for the OSD-w sweep (all may be null), indexed by position in idx
    int32_t *rec_ordering;        // [count][n]
    int32_t *rec_pivcol;          // [count][m]  pivot column (permuted index) of position k, -1 if none
    uint8_t *rec_sred;            // [count][m]  reduced syndrome by position
    int32_t *rec_npiv;            // [count]
};

template <typename K> struct KeyBits;
template <> struct KeyBits<float> {
    typedef uint32_t type;
    __device__ static __forceinline__ type get(float x) { return __float_as_uint(x) & 0x7fffffffu; }
};
template <> struct KeyBits<double> {
    typedef unsigned long long type;
    __device__ static __forceinline__ type get(double x) { return (unsigned long long)__double_as_longlong(x) & 0x7fffffffffffffffull; }
};

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

template <typename K, int WM>
__global__ void __launch_bounds__(OSD_WARPS * 32)
osd0_kernel(const OSDParams P)
{
    typedef typename KeyBits<K>::type kbits;
    const int m = P.m, n = P.n, WN = P.WN;
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
    const unsigned FULL = 0xffffffffu;

    for (long long it = (long long)blockIdx.x * OSD_WARPS + warp; it < count; it += nwarps) {
        const long long shot = P.idx ? (long long)P.idx[it] : it;
        const K *llr = reinterpret_cast<const K *>(P.llr) + (size_t)shot * n;
        const uint32_t *hard = P.hard + (size_t)shot * WN;

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
        for (int w = 0; w < WM; ++w) rs[w] = __reduce_xor_sync(FULL, rs[w]) ^ P.synd[(size_t)shot * WM + w];

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
        const int rank = P.rank;
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
        for (int w = lane; w < WN; w += 32) P.out[(size_t)shot * WN + w] = solw[w];
        if (lane == 0 && P.valid) P.valid[shot] = any_bad ? 0 : 1;

        // ---- elimination record for the OSD-w sweep -----------------------------------------
        if (P.rec_ordering) {
            for (int j = lane; j < n; j += 32) P.rec_ordering[(size_t)it * n + j] = ord[j];
            for (int r = lane; r < m; r += 32) P.rec_pivcol[(size_t)it * m + r] = -1;
            __syncwarp();
#pragma unroll
            for (int i = 0; i < WM; ++i) {
                if (lane + 32 * i < m) {
                    P.rec_pivcol[(size_t)it * m + pos[i]] = pcol[i];
                    P.rec_sred[(size_t)it * m + pos[i]] = (uint8_t)bbit[i];
                }
            }
            if (lane == 0) P.rec_npiv[it] = row;
        }
        __syncwarp();
    }
}

}  // namespace qldpc
