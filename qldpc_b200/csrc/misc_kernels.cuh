// Small data-path kernels around the decoders: bit packing at the host boundary, the device
// error/syndrome sampler, and the packed XOR-popcount syndrome / logical checks with counters.
//
// Replaces (reference paths relative to michelebanfi/qLDPC):
//   decoding/beliefPropagationGPU.py:181-200  generate_errors_and_syndromes_batch  (sample_kernel)
//   paperResults.py:83-100, rework/main.py:90-112, BP_per_Iteration.py:66-77        (check_kernel)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace qldpc {

// ---- u8 [B][nbits] <-> packed u32 [B][W] -------------------------------------------------------
// One thread per output word: 32 byte loads (L1-resident rows), one coalesced word store.
__global__ void pack_bits_kernel(const uint8_t *__restrict__ in, uint32_t *__restrict__ out, long long B, int nbits, int W)
{
    const long long total = B * W;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long s = t / W;
        const int w = (int)(t - s * W);
        const uint8_t *row = in + (size_t)s * nbits;
        uint32_t x = 0;
        const int hi = min(32, nbits - 32 * w);
        for (int b = 0; b < hi; ++b) x |= (uint32_t)(row[32 * w + b] & 1u) << b;
        out[t] = x;
    }
}

__global__ void unpack_bits_kernel(const uint32_t *__restrict__ in, uint8_t *__restrict__ out, long long B, int nbits, int W)
{
    const long long total = B * (long long)nbits;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long s = t / nbits;
        const int j = (int)(t - s * nbits);
        out[t] = (uint8_t)((in[(size_t)s * W + (j >> 5)] >> (j & 31)) & 1u);
    }
}

// nbits % 16 == 0: one thread expands 16 bits into 16 bytes (one 128-bit store; the 16 bits never straddle a word)
__global__ void unpack_bits16_kernel(const uint32_t *__restrict__ in, uint4 *__restrict__ out, long long B, int nbits, int W)
{
    const int per_row = nbits >> 4;
    const long long total = B * (long long)per_row;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long s = t / per_row;
        const int j = (int)(t - s * per_row) << 4;
        const uint32_t bits = (in[(size_t)s * W + (j >> 5)] >> (j & 31)) & 0xffffu;
        uint32_t w[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t nib = (bits >> (4 * q)) & 0xfu;
            w[q] = (nib & 1u) | ((nib & 2u) << 7) | ((nib & 4u) << 14) | ((nib & 8u) << 21);
        }
        out[t] = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

template <typename TI, typename TO>
__global__ void cast_kernel(const TI *__restrict__ in, TO *__restrict__ out, long long N)
{
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < N; t += (long long)gridDim.x * blockDim.x)
        out[t] = (TO)in[t];
}

// ---- Philox4x32-10 (Salmon et al., SC'11), counter = (shot id, block), key = seed --------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}

struct SampleParams {
    int m, n, WM, WN;
    const uint32_t *colmask;     // [n][WM]
    long long B;
    unsigned long long first_shot;   // global id of shot 0 of this launch (multi-GPU: rank offset)
    unsigned long long seed;
    uint32_t threshold;          // bit j of a shot is 1  <=>  u32 draw < threshold  (= floor(p * 2^32))
    int draws;                   // 1, or 2 = XOR of two independent draws (paperResults.py:61-63)
    uint32_t *err;               // [B][WN]
    uint32_t *synd;              // [B][WM]
};

// One thread per shot.  Variable j uses word (j & 3) of Philox block (j >> 2) of stream `draw`.
// WM == 0 (more than 160 checks): only the errors are produced here, the syndromes by syndrome_kernel.
// The result depends only on (seed, global shot id), never on the launch shape or the rank.
template <int WM>
__global__ void sample_kernel(const SampleParams P)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= P.B) return;
    const unsigned long long sid = P.first_shot + (unsigned long long)t;
    const uint2 key = make_uint2((uint32_t)P.seed, (uint32_t)(P.seed >> 32));
    constexpr int WR = WM > 0 ? WM : 1;
    uint32_t sy[WR];
#pragma unroll
    for (int w = 0; w < WR; ++w) sy[w] = 0;
    for (int w = 0; w < P.WN; ++w) {
        uint32_t bits = 0;
        const int hi = min(32, P.n - 32 * w);
        for (int d = 0; d < P.draws; ++d) {
            uint32_t bd = 0;
            for (int q = 0; q < 8; ++q) {
                const uint4 r = philox4x32_10(make_uint4((uint32_t)sid, (uint32_t)(sid >> 32), (uint32_t)(w * 8 + q), (uint32_t)d), key);
                bd |= (uint32_t)(r.x < P.threshold) << (4 * q);
                bd |= (uint32_t)(r.y < P.threshold) << (4 * q + 1);
                bd |= (uint32_t)(r.z < P.threshold) << (4 * q + 2);
                bd |= (uint32_t)(r.w < P.threshold) << (4 * q + 3);
            }
            bits ^= bd;
        }
        if (hi < 32) bits &= (1u << hi) - 1u;
        P.err[(size_t)t * P.WN + w] = bits;
        if (WM > 0) {
            uint32_t x = bits;
            while (x) {
                const int b = __ffs(x) - 1;
                x &= x - 1;
                const int v = 32 * w + b;
#pragma unroll
                for (int k = 0; k < WR; ++k) sy[k] ^= P.colmask[v * WR + k];
            }
        }
    }
    if (WM > 0) {
#pragma unroll
        for (int k = 0; k < WR; ++k) P.synd[(size_t)t * WR + k] = sy[k];
    }
}

// ---- measurement errors: every syndrome bit flipped with probability q, the "simple phenomenological error model" the
// reference keeps one uncomment away (paperResults.py:66-68).  One thread per (shot, 32-check word); Philox stream 0x4d454153
// ("MEAS") of the same (seed, global shot id) counter space as sample_kernel, so results are shard-invariant too.
__global__ void meas_noise_kernel(uint32_t *synd, long long B, int m, int WM, uint32_t threshold, unsigned long long seed,
                                  unsigned long long first_shot)
{
    const long long total = B * WM;
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long s = t / WM;
        const int w = (int)(t - s * WM);
        const unsigned long long sid = first_shot + (unsigned long long)s;
        uint32_t bits = 0;
        for (int q = 0; q < 8; ++q) {
            const uint4 r = philox4x32_10(make_uint4((uint32_t)sid, (uint32_t)(sid >> 32), (uint32_t)(w * 8 + q), 0x4d454153u), key);
            bits |= (uint32_t)(r.x < threshold) << (4 * q);
            bits |= (uint32_t)(r.y < threshold) << (4 * q + 1);
            bits |= (uint32_t)(r.z < threshold) << (4 * q + 2);
            bits |= (uint32_t)(r.w < threshold) << (4 * q + 3);
        }
        const int hi = min(32, m - 32 * w);
        if (hi < 32) bits &= (1u << hi) - 1u;
        synd[t] ^= bits;
    }
}

// ---- syndromes of given errors: synd = err * H^T mod 2 (paperResults.py:65, beliefPropagationGPU.py:198) -------
// One thread per (shot, 32-check word): parity of the packed error bits over the columns of each row (CSR of H).
__global__ void syndrome_kernel(const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col_idx, int m, int WM, int WN,
                                long long B, const uint32_t *__restrict__ err, uint32_t *__restrict__ synd)
{
    const long long total = B * WM;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long s = t / WM;
        const int w = (int)(t - s * WM);
        const uint32_t *e = err + (size_t)s * WN;
        uint32_t out = 0;
        const int hi = min(32, m - 32 * w);
        for (int b = 0; b < hi; ++b) {
            const int c = 32 * w + b;
            uint32_t par = 0;
            for (int a = row_ptr[c]; a < row_ptr[c + 1]; ++a) {
                const int v = col_idx[a];
                par ^= e[v >> 5] >> (v & 31);
            }
            out |= (par & 1u) << b;
        }
        synd[t] = out;
    }
}

// ---- posterior-LLR histograms (BP_per_Iteration.py:56,60; rework/Alvarado.py:159-162) -----------------------------
// hist[0]: LLRs of variables whose true error bit is 0, hist[1]: true bit 1, hist[2]: all LLRs of BP-failed shots.
// Bins are uniform over [lo, hi); values outside go to the first / last bin.  Shared-memory privatised counters.
template <typename T>
__global__ void __launch_bounds__(256) llr_hist_kernel(const T *__restrict__ llr, const uint32_t *__restrict__ err, const uint8_t *__restrict__ conv,
                                                       long long B, int n, int WN, double lo, double hi, int nbins, unsigned long long *hist,
                                                       unsigned long long *n_failed)
{
    extern __shared__ unsigned int s_h[];                 // [3][nbins]
    for (int i = threadIdx.x; i < 3 * nbins; i += blockDim.x) s_h[i] = 0;
    __syncthreads();
    const double scale = nbins / (hi - lo);
    const long long total = B * (long long)n;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long s = t / n;
        const int v = (int)(t - s * n);
        const double x = (double)llr[t];
        int bin = (int)floor((x - lo) * scale);
        bin = max(0, min(nbins - 1, bin));
        const int bit = err ? (int)((err[(size_t)s * WN + (v >> 5)] >> (v & 31)) & 1u) : 0;
        atomicAdd(&s_h[bit * nbins + bin], 1u);
        if (conv && !conv[s]) {
            atomicAdd(&s_h[2 * nbins + bin], 1u);
            if (v == 0) atomicAdd(n_failed, 1ull);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * nbins; i += blockDim.x)
        if (s_h[i]) atomicAdd(&hist[i], (unsigned long long)s_h[i]);
}

// ---- checks + counters ------------------------------------------------------------------------
enum {
    CNT_SHOTS = 0,        // shots counted
    CNT_BP_FAILED = 1,    // BP did not converge (== OSD invocations when OSD is on)
    CNT_LOGICAL = 2,      // any(L @ (corr ^ err))                         (paperResults.py:93)
    CNT_LOGICAL_OSD = 3,  // logical && BP failed                          (rework/main.py:99-101)
    CNT_DEGENERATE = 4,   // valid && !logical && corr != err              (paperResults.py:90)
    CNT_MISCORRECTED = 5, // logical && wt(err) <  d//2                    (paperResults.py:97-98)
    CNT_INCORRECTABLE = 6,// logical && wt(err) >= d//2                    (paperResults.py:99-100)
    CNT_INVALID = 7,      // H @ corr != syndrome
    CNT_ITER_SUM = 8,     // sum of 0-based exit iterations                (rework/main.py:84,117)
    CNT_LOGICAL_BP = 9,   // logical && BP converged
    CNT_RESID_WEIGHT = 10,// sum of wt(corr ^ err)
    CNT_ERR_WEIGHT = 11,  // sum of wt(err)
    CNT_NUM = 16
};
enum { FLAG_LOGICAL = 1, FLAG_VALID = 2, FLAG_DEGENERATE = 4, FLAG_BP_CONVERGED = 8 };

struct CheckParams {
    int m, n, k, WM, WN;
    const uint32_t *colmask;     // [n][WM]
    const uint32_t *Lrows;       // [k][WN] packed logical operators (may be null when k == 0)
    long long B;
    const uint32_t *err;         // [B][WN]
    const uint32_t *corr;        // [B][WN]
    const uint32_t *synd;        // [B][WM]
    const uint8_t *conv;         // [B] (may be null: treated as converged)
    const int32_t *iters;        // [B] (may be null)
    int half_distance;           // distance // 2
    unsigned long long *counters;// [CNT_NUM] (may be null)
    uint8_t *flags;              // [B] (may be null)
    int32_t *weight;             // [B] residual weight (may be null)
    const uint32_t *corr_synd;   // [B][WM] syndrome of the correction, used by the WM == 0 instantiation (more than 160 checks)
};

template <int WM>
__global__ void __launch_bounds__(256) check_kernel(const CheckParams P)
{
    __shared__ unsigned long long s_cnt[CNT_NUM];
    if (threadIdx.x < CNT_NUM) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    unsigned long long c[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) c[i] = 0;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < P.B; t += (long long)gridDim.x * blockDim.x) {
        const uint32_t *e = P.err + (size_t)t * P.WN, *x = P.corr + (size_t)t * P.WN;
        constexpr int WR = WM > 0 ? WM : 1;
        uint32_t sy[WR];
#pragma unroll
        for (int k = 0; k < WR; ++k) sy[k] = (WM > 0) ? P.synd[(size_t)t * WR + k] : 0u;
        int wres = 0, werr = 0;
        bool differs = false;
        for (int w = 0; w < P.WN; ++w) {
            const uint32_t ew = e[w], xw = x[w];
            wres += __popc(ew ^ xw);
            werr += __popc(ew);
            differs |= (ew != xw);
            if (WM > 0) {
                uint32_t y = xw;             // syndrome of the correction: XOR of the packed columns
                while (y) {
                    const int b = __ffs(y) - 1;
                    y &= y - 1;
                    const int v = 32 * w + b;
#pragma unroll
                    for (int k = 0; k < WR; ++k) sy[k] ^= P.colmask[v * WR + k];
                }
            }
        }
        bool valid = true;
        if (WM > 0) {
#pragma unroll
            for (int k = 0; k < WR; ++k) valid = valid && (sy[k] == 0);
        } else {
            for (int k = 0; k < P.WM; ++k) valid = valid && (P.corr_synd[(size_t)t * P.WM + k] == P.synd[(size_t)t * P.WM + k]);
        }
        bool logical = false;
        for (int r = 0; r < P.k; ++r) {
            uint32_t par = 0;
            for (int w = 0; w < P.WN; ++w) par ^= P.Lrows[r * P.WN + w] & (e[w] ^ x[w]);
            logical |= (__popc(par) & 1);
        }
        const bool conv = P.conv ? (P.conv[t] != 0) : true;
        const bool degenerate = valid && !logical && differs;
        c[CNT_SHOTS] += 1;
        c[CNT_BP_FAILED] += !conv;
        c[CNT_LOGICAL] += logical;
        c[CNT_LOGICAL_OSD] += logical && !conv;
        c[CNT_DEGENERATE] += degenerate;
        c[CNT_MISCORRECTED] += logical && (werr < P.half_distance);
        c[CNT_INCORRECTABLE] += logical && (werr >= P.half_distance);
        c[CNT_INVALID] += !valid;
        c[CNT_ITER_SUM] += P.iters ? (unsigned long long)P.iters[t] : 0ull;
        c[CNT_LOGICAL_BP] += logical && conv;
        c[CNT_RESID_WEIGHT] += wres;
        c[CNT_ERR_WEIGHT] += werr;
        if (P.flags) P.flags[t] = (uint8_t)((logical ? FLAG_LOGICAL : 0) | (valid ? FLAG_VALID : 0) |
                                            (degenerate ? FLAG_DEGENERATE : 0) | (conv ? FLAG_BP_CONVERGED : 0));
        if (P.weight) P.weight[t] = wres;
    }
    if (P.counters) {
#pragma unroll
        for (int i = 0; i < 12; ++i) {
            unsigned long long v = c[i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if ((threadIdx.x & 31) == 0 && v) atomicAdd(&s_cnt[i], v);
        }
        __syncthreads();
        if (threadIdx.x < 12 && s_cnt[threadIdx.x]) atomicAdd(&P.counters[threadIdx.x], s_cnt[threadIdx.x]);
    }
}

// ---- alpha estimation (rework/Alvarado.py:10-66) without the messages ----------------------------------------------
// The estimator histograms R_new / alpha of performMinSum_Symmetric after the FIRST check pass (decoding.py:58-59) by the
// true value of the edge's bit.  With one prior value L on every variable all incoming messages equal L, so the message
// on every edge of check c is (1 - 2 s_c) L: the histograms are four numbers, counts[2 * bit + s] = number of edges whose
// variable has error bit `bit` and whose check has syndrome bit s.  One thread per shot: k = |row_c & error|, s = k mod 2.
__global__ void alpha_counts_kernel(const uint32_t *__restrict__ Hrows, const int32_t *__restrict__ row_ptr, int m, int WN, long long B,
                                    const uint32_t *__restrict__ err, unsigned long long *counts)
{
    unsigned int c4[4] = {0u, 0u, 0u, 0u};
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < B; t += (long long)gridDim.x * blockDim.x) {
        const uint32_t *e = err + (size_t)t * WN;
        for (int c = 0; c < m; ++c) {
            int k = 0;
            for (int w = 0; w < WN; ++w) k += __popc(e[w] & Hrows[(size_t)c * WN + w]);
            const int rw = row_ptr[c + 1] - row_ptr[c], s = k & 1;
            c4[2 + s] += (unsigned int)k;
            c4[s] += (unsigned int)(rw - k);
        }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        unsigned int v = c4[q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&counts[q], (unsigned long long)v);
    }
}

// Shots of a list (idx / count, identity when idx is null) whose OSD-0 solution missed the syndrome (valid == 0): compacted
// into out_idx, count in out_count (zeroed by the caller).  The OSD-w sweep runs on that list (osdw_kernel.cuh).
__global__ void compact_invalid_kernel(const int32_t *idx, const unsigned int *count_dev, long long count_host, const uint8_t *valid,
                                       int32_t *out_idx, unsigned int *out_count)
{
    const long long count = count_dev ? (long long)*count_dev : count_host;
    for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < count; it += (long long)gridDim.x * blockDim.x) {
        const int32_t shot = idx ? idx[it] : (int32_t)it;
        if (!valid[shot]) out_idx[atomicAdd(out_count, 1u)] = shot;
    }
}

}  // namespace qldpc
