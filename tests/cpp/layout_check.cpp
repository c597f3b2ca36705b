// Host-only check of the lane labelling of the warp- / CTA-per-shot BP kernels (qldpc_b200/csrc/bp_warp_layout.h):
// builds the labelling for a check matrix read from a text file and verifies the tables the kernels consume.
//   layout_check <graph.txt> <RW> <check_slots> <var_slots> [lanes = 32 | 16]      graph.txt: m n / row_ptr / col_idx / var_ptr / var_edge / edge_check
#include "../../qldpc_b200/csrc/bp_warp_layout.h"

#include <cstdio>
#include <cstdlib>
#include <set>

#define REQUIRE(c)                                                             \
    do {                                                                       \
        if (!(c)) { std::printf("FAILED %s (line %d)\n", #c, __LINE__); return 1; } \
    } while (0)

int main(int argc, char **argv)
{
    if (argc < 5) return 2;
    FILE *f = std::fopen(argv[1], "r");
    if (!f) return 2;
    int m, n;
    if (std::fscanf(f, "%d %d", &m, &n) != 2) return 2;
    std::vector<int32_t> rp(m + 1);
    for (auto &x : rp) if (std::fscanf(f, "%d", &x) != 1) return 2;
    const int E = rp[m];
    std::vector<int32_t> ci(E), vp(n + 1), ve(E), ec(E);
    for (auto &x : ci) if (std::fscanf(f, "%d", &x) != 1) return 2;
    for (auto &x : vp) if (std::fscanf(f, "%d", &x) != 1) return 2;
    for (auto &x : ve) if (std::fscanf(f, "%d", &x) != 1) return 2;
    for (auto &x : ec) if (std::fscanf(f, "%d", &x) != 1) return 2;
    const int RW = std::atoi(argv[2]), cs = std::atoi(argv[3]), vs = std::atoi(argv[4]), W = argc > 5 ? std::atoi(argv[5]) : 32;

    qldpc::WarpLayoutBuilder b(m, n, rp.data(), ci.data(), vp.data(), ve.data(), ve.data(), ec.data(), RW, cs, vs, W);
    const bool ok = b.construct();
    const qldpc::WarpLayout L = b.tables();
    std::printf("m=%d n=%d RW=%d slots %d/%d conflict_free=%d natural=%d cost=%d floor=%d\n", m, n, RW, L.CPL, L.VPL, (int)ok, L.cost_natural,
                L.cost, L.floor);
    REQUIRE(ok && L.cost == L.floor);
    const int CPL = L.CPL, VPL = L.VPL;
    // every check and every variable sits at exactly one position
    std::vector<int> cpos(m, -1), vpos(n, -1);
    for (int p = 0; p < CPL * 32; ++p)
        if (L.cinfo[p] != 0xffffffffu) { REQUIRE((int)L.cinfo[p] < m && cpos[L.cinfo[p]] < 0); cpos[L.cinfo[p]] = p; }
    for (int p = 0; p < VPL * 32; ++p)
        if (L.vorig[p] != 0xffffffffu) { REQUIRE((int)L.vorig[p] < n && vpos[L.vorig[p]] < 0); vpos[L.vorig[p]] = p; }
    for (int c = 0; c < m; ++c) REQUIRE(cpos[c] >= 0);
    for (int v = 0; v < n; ++v) { REQUIRE(vpos[v] >= 0); REQUIRE(L.vpos[v] == 4u * (uint32_t)vpos[v]); }
    // every edge of H occupies exactly one edge slot of its check: it reads the posterior of its variable and delivers
    // into the plane of its position in the variable's addition order; everything else is padding (+inf row, dump row)
    std::vector<int> t_of_edge(E, -1);
    for (int v = 0; v < n; ++v)
        for (int q = vp[v]; q < vp[v + 1]; ++q) t_of_edge[ve[q]] = q - vp[v];
    std::multiset<std::pair<int, int>> want, got;          // (check, variable)
    for (int c = 0; c < m; ++c)
        for (int e = rp[c]; e < rp[c + 1]; ++e) want.insert({c, ci[e]});
    std::set<uint32_t> targets;
    for (int i = 0; i < CPL; ++i)
        for (int k = 0; k < RW; ++k) {
            int bank_rd[32] = {0}, bank_wr[32] = {0};
            for (int l = 0; l < 32; ++l) {
                const size_t at = (size_t)(i * RW + k) * 32 + l;
                const uint32_t vword = L.vidx[at] / 4, sword = L.sidx[at] / 4;
                REQUIRE(L.vidx[at] % 4 == 0 && L.sidx[at] % 4 == 0);
                // bank-conflict free: 32 lanes in 32 different columns (32-bit words), or -- 64-bit words, lanes = 16 -- the 16
                // lanes of each half-warp in 16 different bank pairs
                if (W == 32) REQUIRE(++bank_rd[vword & 31] == 1 && ++bank_wr[sword & 31] == 1);
                else REQUIRE(++bank_rd[(l / 16) * 16 + (vword & 15)] == 1 && ++bank_wr[(l / 16) * 16 + (sword & 15)] == 1);
                const uint32_t c = L.cinfo[i * 32 + l];
                if (vword >= (uint32_t)VPL * 32) {                                       // padding slot
                    REQUIRE(vword < (uint32_t)(VPL + 1) * 32);
                    REQUIRE(sword >= (uint32_t)3 * VPL * 32 && sword < (uint32_t)(3 * VPL + 1) * 32);
                    continue;
                }
                REQUIRE(c != 0xffffffffu);
                const uint32_t v = L.vorig[vword];
                REQUIRE(v != 0xffffffffu);
                got.insert({(int)c, (int)v});
                int e = -1;
                for (int q = rp[c]; q < rp[c + 1]; ++q) if (ci[q] == (int)v) e = q;
                REQUIRE(e >= 0);
                REQUIRE(sword == (uint32_t)(t_of_edge[e] * VPL * 32) + vword);           // plane t, column of the variable
                REQUIRE(targets.insert(sword).second);                                   // no two edges deliver into the same word
                REQUIRE(L.sidx0[at] == L.sidx[at]);                                      // (one addition order given)
            }
        }
    REQUIRE(want == got);
    std::printf("LAYOUT-OK\n");
    return 0;
}
