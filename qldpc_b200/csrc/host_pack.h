// Host side of the uint8-per-bit interface (the reference's dtypes: syndromes [B][m] and corrections [B][n] as 0/1 bytes).
//
// qldpc_bposd_decode_host moves 221 bytes per [[144,12,12]] shot over PCIe when the byte rows themselves are copied (72 in,
// 149 out) and 37 when the rows cross the bus bit-packed.  With enough host threads the packing is done HERE, on the CPU,
// between the caller's arrays and pinned staging buffers: SSE2 movemask for bytes -> bits, a 256-entry table of 8-byte
// expansions and non-temporal stores for bits -> bytes (the output is written once and never read back: no read-for-ownership
// traffic).  A small persistent thread pool splits a chunk into contiguous shot ranges.
//
// Bit order and semantics equal pack_bits_kernel / unpack_bits_kernel (misc_kernels.cuh): bit b of a row = byte b & 1, little-
// endian uint32 words, unused bits of the last word zero.
#pragma once
#include <stdint.h>
#include <string.h>

#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#if defined(__SSE2__) || defined(__x86_64__)
#include <emmintrin.h>
#define QLDPC_HOST_SSE2 1
#endif

namespace qldpc {

class HostPool {
public:
    explicit HostPool(int n) : n_(n < 1 ? 1 : n)
    {
        for (int i = 1; i < n_; ++i) th_.emplace_back([this, i] { loop(i); });
    }
    ~HostPool()
    {
        {
            std::lock_guard<std::mutex> g(mu_);
            stop_ = true;
            ++gen_;
        }
        cv_.notify_all();
        for (auto &t : th_) t.join();
    }
    int size() const { return n_; }
    // f(tid, nthreads) on every thread of the pool; the caller is thread 0; returns when all are done
    void run(const std::function<void(int, int)> &f)
    {
        if (n_ > 1) {
            {
                std::lock_guard<std::mutex> g(mu_);
                fn_ = &f;
                pending_ = n_ - 1;
                ++gen_;
            }
            cv_.notify_all();
        }
        f(0, n_);
        if (n_ > 1) {
            std::unique_lock<std::mutex> l(mu_);
            done_.wait(l, [&] { return pending_ == 0; });
        }
    }

private:
    void loop(int tid)
    {
        unsigned seen = 0;
        for (;;) {
            const std::function<void(int, int)> *f;
            {
                std::unique_lock<std::mutex> l(mu_);
                cv_.wait(l, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
                f = fn_;
            }
            (*f)(tid, n_);
            {
                std::lock_guard<std::mutex> g(mu_);
                if (--pending_ == 0) done_.notify_one();
            }
        }
    }
    int n_;
    std::vector<std::thread> th_;
    std::mutex mu_;
    std::condition_variable cv_, done_;
    const std::function<void(int, int)> *fn_ = nullptr;
    int pending_ = 0;
    unsigned gen_ = 0;
    bool stop_ = false;
};

// ---- bytes -> bits ---------------------------------------------------------------------------------
static inline void host_pack_row(const uint8_t *r, uint32_t *o, int nbits, int W)
{
    int b = 0, w = 0;
#ifdef QLDPC_HOST_SSE2
    for (; b + 32 <= nbits; b += 32, ++w) {
        const __m128i v0 = _mm_loadu_si128(reinterpret_cast<const __m128i *>(r + b));
        const __m128i v1 = _mm_loadu_si128(reinterpret_cast<const __m128i *>(r + b + 16));
        // bit 0 of every byte to its bit 7 (a 16-bit shift moves bit 8 to bit 15 as well), then one bit per byte
        o[w] = (uint32_t)_mm_movemask_epi8(_mm_slli_epi16(v0, 7)) | ((uint32_t)_mm_movemask_epi8(_mm_slli_epi16(v1, 7)) << 16);
    }
#endif
    for (; w < W; ++w) {
        uint32_t x = 0;
        int i = 0;
        for (; i + 8 <= 32 && b + 8 <= nbits; i += 8, b += 8) {      // 8 bytes -> 8 bits with one multiplication
            uint64_t v;
            memcpy(&v, r + b, 8);
            x |= (uint32_t)(((v & 0x0101010101010101ull) * 0x0102040810204080ull) >> 56) << i;
        }
        for (; i < 32 && b < nbits; ++i, ++b) x |= (uint32_t)(r[b] & 1u) << i;
        o[w] = x;
    }
}

// ---- bits -> bytes ---------------------------------------------------------------------------------
struct HostExpandTable {
    uint64_t t[256];
    HostExpandTable()
    {
        for (int v = 0; v < 256; ++v) {
            uint64_t x = 0;
            for (int i = 0; i < 8; ++i) x |= (uint64_t)((v >> i) & 1) << (8 * i);
            t[v] = x;
        }
    }
};
static inline const uint64_t *host_expand_table()
{
    static const HostExpandTable T;
    return T.t;
}

// rows [s0, s1) of `in` ([B][W] words) into `out` ([B][nbits] bytes)
static inline void host_unpack_rows(const uint32_t *in, uint8_t *out, long long s0, long long s1, int nbits, int W)
{
    const uint64_t *T = host_expand_table();
#ifdef QLDPC_HOST_SSE2
    if (nbits % 16 == 0 && (reinterpret_cast<uintptr_t>(out) & 15u) == 0) {
        // every 16-bit piece is one aligned 16-byte non-temporal store
        const int pieces = nbits / 16;
        for (long long s = s0; s < s1; ++s) {
            const uint16_t *ib = reinterpret_cast<const uint16_t *>(in + (size_t)s * W);
            __m128i *o = reinterpret_cast<__m128i *>(out + (size_t)s * nbits);
            for (int j = 0; j < pieces; ++j) {
                const uint32_t h = ib[j];
                _mm_stream_si128(o + j, _mm_set_epi64x((long long)T[h >> 8], (long long)T[h & 0xffu]));
            }
        }
        _mm_sfence();
        return;
    }
#endif
    for (long long s = s0; s < s1; ++s) {
        const uint8_t *ib = reinterpret_cast<const uint8_t *>(in + (size_t)s * W);
        uint8_t *o = out + (size_t)s * nbits;
        int j = 0;
        for (; 8 * j + 8 <= nbits; ++j) memcpy(o + 8 * j, &T[ib[j]], 8);
        for (int b = 8 * j; b < nbits; ++b) o[b] = (uint8_t)((ib[b >> 3] >> (b & 7)) & 1u);
    }
}

static inline void host_pack_rows(HostPool &pool, const uint8_t *in, uint32_t *out, long long B, int nbits, int W)
{
    pool.run([&](int tid, int nt) {
        const long long s0 = B * tid / nt, s1 = B * (tid + 1) / nt;
        for (long long s = s0; s < s1; ++s) host_pack_row(in + (size_t)s * nbits, out + (size_t)s * W, nbits, W);
    });
}

static inline void host_unpack_rows(HostPool &pool, const uint32_t *in, uint8_t *out, long long B, int nbits, int W)
{
    pool.run([&](int tid, int nt) { host_unpack_rows(in, out, B * tid / nt, B * (tid + 1) / nt, nbits, W); });
}

}  // namespace qldpc
