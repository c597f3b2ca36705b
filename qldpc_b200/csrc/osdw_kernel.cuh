// OSD-w combination sweep (order > 0) on sm_100a -- one CTA per shot whose OSD-0 solution does
// not satisfy the syndrome.
//
// Replaces decoding/OSD_enhanced.py:66-131 (+ recompute_solution :134-155, compute_metric
// :158-177; verbatim copy in rework/decoding.py:254-347).  For syndromes of the form e * H^T the
// reference returns the OSD-0 solution before reaching this code (OSD_enhanced.py:58-60,
// SURVEY.md H5), so this kernel only ever runs on syndromes outside the column space of H; it is
// written for exactness, not speed.  Every quirk of the reference is kept:
//   * candidate positions = first min(#non-pivot, order + 10) non-pivot PERMUTED positions (:80-81);
//   * candidates in itertools.combinations order, weight 1..order, cut at max_combinations (:89-97);
//   * recompute_solution walks the pivots in order, Gauss-Seidel style, using the UNREDUCED
//     permuted row `r` of H together with the REDUCED syndrome bit s_reduced[r] (:144-153);
//   * metric = (1e10 + 1e8 * #unsatisfied checks, if any) + sum_i solution_i * |llr_i|, float64,
//     the sum in NumPy's pairwise order (:163-175);
//   * selection rule :117-127: first valid candidate replaces the incumbent unconditionally, then
//     only valid candidates with a strictly smaller metric; with no valid candidate at all, the
//     strictly-smallest metric wins, OSD-0 being the first incumbent.
// Candidates are independent, so the sequential rule collapses to a lexicographic minimum over
// (invalid?, metric, enumeration index), evaluated in parallel: one candidate per thread.
//
// The kernel runs on a device-side list of shots (idx / count_dev: the shots whose OSD-0 solution missed the syndrome,
// compacted by compact_invalid_kernel -- no host synchronisation) and RECOMPUTES the elimination record (ordering, pivot
// columns by position, reduced syndrome) in shared memory: stable argsort of |llr| and the reference's row-major
// Gauss-Jordan by warp 0, the very code of osd_kernel.cuh (template parameter WM, <= 5 words = 160 rows).  In the BP -> OSD
// pipelines the BP hard decision is taken from the sign of the posterior LLR (hard = llr < 0), because the OSD-0 stage has
// already overwritten it with its solution.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>

#include "osd_kernel.cuh"

namespace qldpc {

struct OSDWParams {
    int m, n, WM, WN;
    int rank;                     // GF(2) rank of H
    const uint32_t *Hrows;        // [m][WN] packed rows of H, original column order
    const uint32_t *colmask;      // [n][WM] packed columns of H
    const int32_t *idx;           // [count] shot ids to process (null: identity)
    const unsigned int *count_dev;// number of entries (device) ...
    long long count;              // ... or given by the host when count_dev == null
    const uint32_t *synd;         // [B][WM]
    const void *llr;              // [B][n] double, or float when llr_f32
    int llr_f32;
    const uint32_t *hard;         // [B][WN] BP hard decision; null: llr < 0
    uint32_t *sol;                // [B][WN] out: best solution
    const uint8_t *valid;         // [B]  1: OSD-0 solution satisfies the syndrome -> untouched (null: sweep every shot)
    int order;
    long long max_combinations;   // 0: no limit
};

constexpr int OSDW_THREADS = 256;
constexpr int OSDW_MAX_T = 64;     // order + 10 <= 64
constexpr int OSDW_MAX_ORDER = 16;

struct OSDWKey {
    int invalid;
    double metric;
    long long index;              // -1: OSD-0 incumbent
};
__device__ __forceinline__ bool osdw_less(const OSDWKey &a, const OSDWKey &b)
{
    if (a.invalid != b.invalid) return a.invalid < b.invalid;
    if (a.metric != b.metric) return a.metric < b.metric;
    return a.index < b.index;
}

// NumPy pairwise sum (loops_utils.h DOUBLE_pairwise_sum) of the virtual array
// a[i] = bit_i(sol) ? absllr[i] : 0.0, i in [lo, lo + cnt)
__device__ double osdw_pairwise(const double *absllr, const uint32_t *sol, int S, int lo, int cnt)
{
    auto a = [&](int i) -> double { return ((sol[(i >> 5) * S] >> (i & 31)) & 1u) ? absllr[i] : 0.0; };
    if (cnt < 8) {
        double res = 0.0;
        for (int i = 0; i < cnt; ++i) res = __dadd_rn(res, a(lo + i));
        return res;
    }
    if (cnt <= 128) {
        double r[8];
        for (int j = 0; j < 8; ++j) r[j] = a(lo + j);
        int i;
        for (i = 8; i < cnt - (cnt % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], a(lo + i + j));
        double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                               __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
        for (; i < cnt; ++i) res = __dadd_rn(res, a(lo + i));
        return res;
    }
    int n2 = cnt / 2;
    n2 -= n2 % 8;
    return __dadd_rn(osdw_pairwise(absllr, sol, S, lo, n2), osdw_pairwise(absllr, sol, S, lo + n2, cnt - n2));
}

template <int WME>
__global__ void __launch_bounds__(OSDW_THREADS) osdw_kernel(const OSDWParams P)
{
    const int m = P.m, n = P.n, WM = P.WM, WN = P.WN;
    const int tid = threadIdx.x, S = OSDW_THREADS;
    extern __shared__ __align__(16) unsigned char smem[];
    // carve-up
    double *absllr = reinterpret_cast<double *>(smem);                 // [n]
    long long *binom = reinterpret_cast<long long *>(absllr + n);      // [(MAX_T+1)][(MAX_ORDER+1)]
    uint32_t *Hp = reinterpret_cast<uint32_t *>(binom + (OSDW_MAX_T + 1) * (OSDW_MAX_ORDER + 1)); // [m][WN] permuted, unreduced
    uint32_t *eperm = Hp + (size_t)m * WN;                             // [WN]
    uint32_t *hardp = eperm + WN;                                      // [WN] hard decision, original order
    uint32_t *resid = hardp + WN;                                      // [WM] residual syndrome s ^ H*hard
    int *ord = reinterpret_cast<int *>(resid + WM);                    // [n]
    int *pivcol = ord + n;                                             // [m]
    int *testpos = pivcol + m;                                         // [MAX_T]
    uint32_t *sred = reinterpret_cast<uint32_t *>(testpos + OSDW_MAX_T);// [m] (one per word, simple)
    uint32_t *scratch_e = sred + m;                                    // [WN][S] per-thread permuted candidate
    uint32_t *scratch_s = scratch_e + (size_t)WN * S;                  // [WN][S] per-thread solution, original order
    // work area of the in-kernel elimination (warp 0)
    unsigned long long *keys = reinterpret_cast<unsigned long long *>((reinterpret_cast<uintptr_t>(scratch_s + (size_t)WN * S) + 7) & ~(uintptr_t)7);  // [n]
    uint32_t *cmask = reinterpret_cast<uint32_t *>(keys + n);          // [n][WM]
    uint32_t *solw = cmask + (size_t)n * WM;                           // [WN]
    uint16_t *ord16 = reinterpret_cast<uint16_t *>(solw + WN);         // [n]
    uint8_t *sred8 = reinterpret_cast<uint8_t *>(ord16 + n);           // [m]
    for (int i = tid; i < n * WM; i += S) cmask[i] = P.colmask[i];
    __shared__ OSDWKey s_key[OSDW_THREADS];
    __shared__ int s_npiv, s_T;
    __shared__ long long s_total;

    // Pascal triangle (exact in int64 for the sizes allowed)
    for (int i = tid; i < (OSDW_MAX_T + 1) * (OSDW_MAX_ORDER + 1); i += S) binom[i] = 0;
    __syncthreads();
    if (tid == 0) {
        for (int a = 0; a <= OSDW_MAX_T; ++a) {
            binom[a * (OSDW_MAX_ORDER + 1)] = 1;
            for (int b = 1; b <= OSDW_MAX_ORDER && b <= a; ++b)
                binom[a * (OSDW_MAX_ORDER + 1) + b] =
                    binom[(a - 1) * (OSDW_MAX_ORDER + 1) + b - 1] + (b <= a - 1 ? binom[(a - 1) * (OSDW_MAX_ORDER + 1) + b] : 0);
        }
    }
    __syncthreads();
    auto C = [&](int a, int b) -> long long { return (b < 0 || b > a) ? 0 : binom[a * (OSDW_MAX_ORDER + 1) + b]; };

    const long long count = P.count_dev ? (long long)*P.count_dev : P.count;
    for (long long it = blockIdx.x; it < count; it += gridDim.x) {
        const long long shot = P.idx ? (long long)P.idx[it] : it;
        if (P.valid && P.valid[shot]) continue;                        // OSD_enhanced.py:58-60
        __syncthreads();
        auto llr_at = [&](int j) -> double {
            return P.llr_f32 ? (double)reinterpret_cast<const float *>(P.llr)[(size_t)shot * n + j]
                             : reinterpret_cast<const double *>(P.llr)[(size_t)shot * n + j];
        };
        for (int j = tid; j < n; j += S) absllr[j] = fabs(llr_at(j));
        for (int w = tid; w < WN; w += S) {
            uint32_t hw = 0;
            if (P.hard) hw = P.hard[(size_t)shot * WN + w];
            else
                for (int b = 0; b < 32 && 32 * w + b < n; ++b) hw |= (uint32_t)(llr_at(32 * w + b) < 0.0) << b;   // hard = values < 0
            hardp[w] = hw;
            eperm[w] = 0;
        }
        {
            // ---- elimination record: stable argsort of |llr| (OSD_enhanced.py:34-35), gf2_elimination (:180-224) by warp 0 ----
            for (int j = tid; j < n; j += S) keys[j] = KeyBits<double>::get(llr_at(j));
            __syncthreads();
            if (tid < 32) {
                for (int i = tid; i < n; i += 32) {
                    const unsigned long long ki = keys[i];
                    int cnt = 0;
                    for (int j = 0; j < n; ++j) cnt += (keys[j] < ki) || (keys[j] == ki && j < i);
                    ord16[cnt] = (uint16_t)i;
                }
                __syncwarp();
                OSDShotIO io;
                io.hard = hardp;
                io.synd = P.synd + (size_t)shot * WM;
                io.out = solw;
                io.valid = nullptr;
                io.rec_ordering = ord;
                io.rec_pivcol = pivcol;
                io.rec_sred = sred8;
                io.rec_npiv = &s_npiv;
                osd0_rowmajor_shot<WME>(m, n, WN, P.rank, io, cmask, ord16, solw, tid);
            }
            __syncthreads();
            for (int r = tid; r < m; r += S) sred[r] = sred8[r];
        }
        __syncthreads();
        const int npiv = s_npiv;
        // permuted unreduced rows: Hp[r] bit j = H[r][ord[j]]
        for (int t = tid; t < m * WN; t += S) {
            const int r = t / WN, w = t - r * WN;
            uint32_t x = 0;
            const int hi = min(32, n - 32 * w);
            for (int b = 0; b < hi; ++b) {
                const int o = ord[32 * w + b];
                x |= ((P.Hrows[(size_t)r * WN + (o >> 5)] >> (o & 31)) & 1u) << b;
            }
            Hp[t] = x;
        }
        // residual syndrome: s ^ H * hard
        for (int w = tid; w < WM; w += S) {
            uint32_t x = P.synd[(size_t)shot * WM + w];
            const int hi = min(32, m - 32 * w);
            for (int b = 0; b < hi; ++b) {
                uint32_t par = 0;
                for (int k = 0; k < WN; ++k) par ^= P.Hrows[(size_t)(32 * w + b) * WN + k] & hardp[k];
                x ^= (uint32_t)(__popc(par) & 1) << b;
            }
            resid[w] = x;
        }
        if (tid == 0) {
            // e_permuted (OSD_enhanced.py:46-50) and the candidate positions (:68-81)
            for (int k = 0; k < npiv; ++k)
                if (sred[k]) eperm[pivcol[k] >> 5] |= 1u << (pivcol[k] & 31);
            int T = 0, k = 0;
            const int want = min(P.order + 10, OSDW_MAX_T);
            for (int j = 0; j < n && T < want; ++j) {
                while (k < npiv && pivcol[k] < j) ++k;
                if (k < npiv && pivcol[k] == j) continue;
                testpos[T++] = j;
            }
            s_T = T;
            long long total = 0;
            const int wmax = min(P.order, T);
            for (int w = 1; w <= wmax; ++w) total += C(T, w);
            if (P.max_combinations > 0 && total > P.max_combinations) total = P.max_combinations;
            s_total = total;
        }
        __syncthreads();
        const int T = s_T;
        const long long total = s_total;
        if (T == 0) continue;                                          // :71-72

        uint32_t *e = scratch_e + tid, *so = scratch_s + tid;
        // evaluates candidate `ci` (-1: the OSD-0 incumbent); leaves its solution in `so`
        auto evaluate = [&](long long ci) -> OSDWKey {
            for (int w = 0; w < WN; ++w) e[w * S] = eperm[w];
            if (ci >= 0) {
                // unrank: weight class, then lexicographic combination of T choose wgt
                long long idx = ci;
                int wgt = 1;
                while (idx >= C(T, wgt)) { idx -= C(T, wgt); ++wgt; }
                int x = 0;
                for (int i = 0; i < wgt; ++i) {
                    while (true) {
                        const long long cnt = C(T - 1 - x, wgt - 1 - i);
                        if (idx < cnt) break;
                        idx -= cnt;
                        ++x;
                    }
                    const int pp = testpos[x];
                    e[(pp >> 5) * S] ^= 1u << (pp & 31);                // :100-102
                    ++x;
                }
                // recompute_solution (:134-155)
                for (int k = 0; k < npiv; ++k) {
                    const int c = pivcol[k];
                    uint32_t par = 0;
                    for (int w = 0; w < WN; ++w) par ^= Hp[k * WN + w] & e[w * S];
                    uint32_t bit = __popc(par) & 1u;
                    const uint32_t hb = (Hp[k * WN + (c >> 5)] >> (c & 31)) & 1u, eb = (e[(c >> 5) * S] >> (c & 31)) & 1u;
                    bit ^= hb & eb;                                     // `col != c`
                    const uint32_t nb = (sred[k] & 1u) ^ bit;
                    e[(c >> 5) * S] = (e[(c >> 5) * S] & ~(1u << (c & 31))) | (nb << (c & 31));
                }
            }
            // unpermute + xor hard (:109-111)
            for (int w = 0; w < WN; ++w) so[w * S] = hardp[w];
            for (int j = 0; j < n; ++j)
                if ((e[(j >> 5) * S] >> (j & 31)) & 1u) so[(ord[j] >> 5) * S] ^= 1u << (ord[j] & 31);
            // unsatisfied checks: residual ^ Hp * e
            int sw = 0;
            for (int r = 0; r < m; ++r) {
                uint32_t par = 0;
                for (int w = 0; w < WN; ++w) par ^= Hp[r * WN + w] & e[w * S];
                sw += ((__popc(par) & 1u) != ((resid[r >> 5] >> (r & 31)) & 1u));
            }
            OSDWKey key;
            key.invalid = sw > 0;
            double metric = sw > 0 ? __dadd_rn(1e10, __dmul_rn((double)sw, 1e8)) : 0.0;   // :167-170
            metric = __dadd_rn(metric, osdw_pairwise(absllr, so, S, 0, n));                // :174-175
            key.metric = metric;
            key.index = ci;
            return key;
        };

        OSDWKey best;
        best.invalid = 2; best.metric = 0.0; best.index = 0x7fffffffffffffffll;
        if (tid == 0) best = evaluate(-1);
        for (long long ci = tid; ci < total; ci += S) {
            const OSDWKey k = evaluate(ci);
            if (osdw_less(k, best)) best = k;
        }
        s_key[tid] = best;
        __syncthreads();
        for (int o = S / 2; o > 0; o >>= 1) {
            if (tid < o && osdw_less(s_key[tid + o], s_key[tid])) s_key[tid] = s_key[tid + o];
            __syncthreads();
        }
        if (tid == 0) {
            const OSDWKey win = s_key[0];
            evaluate(win.index);
            for (int w = 0; w < WN; ++w) P.sol[(size_t)shot * WN + w] = so[w * S];
        }
    }
}

inline size_t osdw_smem_bytes(int m, int n, int WM, int WN)
{
    size_t o = 8 * (size_t)n + 8 * (size_t)(OSDW_MAX_T + 1) * (OSDW_MAX_ORDER + 1);
    o += 4 * ((size_t)m * WN + 2 * (size_t)WN + WM + n + m + OSDW_MAX_T + m + 2 * (size_t)WN * OSDW_THREADS);
    o += 8 + 8 * (size_t)n + 4 * (size_t)n * WM + 4 * (size_t)WN + 2 * (size_t)n + (size_t)m;      // in-kernel elimination
    return o + 64;
}

template <int WME>
inline cudaError_t launch_osdw_inst(const OSDWParams &P, int num_sms, cudaStream_t st)
{
    const size_t smem = osdw_smem_bytes(P.m, P.n, P.WM, P.WN);
    cudaError_t e = cudaFuncSetAttribute(osdw_kernel<WME>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    long long grid = (long long)num_sms * 2;
    if (!P.count_dev) grid = std::min<long long>(grid, P.count);
    osdw_kernel<WME><<<(int)std::max<long long>(1, grid), OSDW_THREADS, smem, st>>>(P);
    return cudaGetLastError();
}

inline cudaError_t launch_osdw(const OSDWParams &P, int num_sms, cudaStream_t st)
{
    if (P.order > OSDW_MAX_ORDER) return cudaErrorInvalidValue;
    switch (P.WM) {
    case 1: return launch_osdw_inst<1>(P, num_sms, st);
    case 2: return launch_osdw_inst<2>(P, num_sms, st);
    case 3: return launch_osdw_inst<3>(P, num_sms, st);
    case 4: return launch_osdw_inst<4>(P, num_sms, st);
    case 5: return launch_osdw_inst<5>(P, num_sms, st);
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace qldpc
