// Host build of qldpc_b200/csrc/host_pack.h for tests/test_host.py and tools: array entry points + a throughput probe.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include "../../qldpc_b200/csrc/host_pack.h"
extern "C" {
void hp_pack(const uint8_t *in, uint32_t *out, long long B, int nbits, int threads)
{
    qldpc::HostPool pool(threads);
    qldpc::host_pack_rows(pool, in, out, B, nbits, (nbits + 31) / 32);
}
void hp_unpack(const uint32_t *in, uint8_t *out, long long B, int nbits, int threads)
{
    qldpc::HostPool pool(threads);
    qldpc::host_unpack_rows(pool, in, out, B, nbits, (nbits + 31) / 32);
}
// seconds per call of pack (m bits) + unpack (n bits) over B shots, `reps` times on one pool
double hp_time(long long B, int m, int n, int threads, int reps)
{
    const int WM = (m + 31) / 32, WN = (n + 31) / 32;
    uint8_t *a = (uint8_t *)aligned_alloc(64, (size_t)B * m + 64), *c = (uint8_t *)aligned_alloc(64, (size_t)B * n + 64);
    uint32_t *p = (uint32_t *)aligned_alloc(64, (size_t)B * WM * 4 + 64), *q = (uint32_t *)aligned_alloc(64, (size_t)B * WN * 4 + 64);
    for (size_t i = 0; i < (size_t)B * m; ++i) a[i] = (uint8_t)((i * 2654435761u >> 13) & 1u);
    for (size_t i = 0; i < (size_t)B * WN; ++i) q[i] = (uint32_t)(i * 2654435761u);
    qldpc::HostPool pool(threads);
    qldpc::host_pack_rows(pool, a, p, B, m, WM);
    qldpc::host_unpack_rows(pool, q, c, B, n, WN);
    const auto t0 = std::chrono::steady_clock::now();
    for (int r = 0; r < reps; ++r) {
        qldpc::host_pack_rows(pool, a, p, B, m, WM);
        qldpc::host_unpack_rows(pool, q, c, B, n, WN);
    }
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / reps;
    free(a); free(c); free(p); free(q);
    return dt;
}
}
