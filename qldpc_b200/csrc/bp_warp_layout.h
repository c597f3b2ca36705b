// Host side of the warp-per-shot BP kernel (bp_warp_kernel.cuh): which lane owns which check / variable, in which
// register slot each edge of a check sits, and the gather tables that follow from that choice.
//
// The kernel's shared-memory traffic per shot-iteration is CPL*RW + VPL conflict-free stores plus two families of
// gathers whose bank conflicts depend only on the labelling:
//   * variable pass, instruction (i, t): lane l reads the t-th added message of the variable at position 32 i + l; the
//     word lives in the bank of the lane that owns the check -> wavefronts = max number of lanes hitting one bank;
//   * message update, instruction (i, k): lane l reads the posterior of the variable of edge slot k of the check at
//     position 32 i + l; bank = lane of the variable; equal words broadcast.
// The labelling is free (the arithmetic of a check does not depend on the order of its edges, the addition order of a
// variable is carried by t and is not touched), so it is chosen by a short deterministic simulated annealing that
// minimises the total number of wavefronts.  Natural labelling of [[144,12,12]]: 55 gather wavefronts against the
// floor of 33 (one per instruction).
#pragma once
#include <stdint.h>

#include <algorithm>
#include <cmath>
#include <vector>

namespace qldpc {

struct WarpLayout {
    int CPL = 0, VPL = 0, RW = 0;
    int cost_natural = 0, cost = 0, floor = 0;  // gather wavefronts per shot-iteration
    // device tables
    std::vector<uint32_t> ridx, ridx0, vidx;    // see bp_warp_kernel.cuh
    std::vector<uint32_t> cinfo;                // [CPL][32] original check index of the position, 0xffffffff = padding
    std::vector<uint32_t> vorig;                // [VPL][32] original variable index of the position, 0xffffffff = padding
    std::vector<uint32_t> vpos;                 // [VPL][32] byte offset in the posterior buffer of variable 32 i + lane
};

class WarpLayoutBuilder {
public:
    double t0 = 2.0, t1 = 0.3;                 // annealing temperatures, in units of the smoothed objective

    WarpLayoutBuilder(int m, int n, const int32_t *row_ptr, const int32_t *col_idx, const int32_t *var_ptr, const int32_t *var_edge0,
                      const int32_t *var_edge1, const int32_t *edge_check, int RW)
        : m(m), n(n), RW(RW), row_ptr(row_ptr, row_ptr + m + 1), col_idx(col_idx, col_idx + row_ptr[m]), var_ptr(var_ptr, var_ptr + n + 1),
          ve0(var_edge0, var_edge0 + row_ptr[m]), ve1(var_edge1, var_edge1 + row_ptr[m]), edge_check(edge_check, edge_check + row_ptr[m])
    {
        CPL = (m + 31) / 32;
        VPL = (n + 31) / 32;
        NI = VPL * 3 + CPL * RW;
        const int E = row_ptr[m];
        t_of_edge.assign(E, 0);
        for (int v = 0; v < n; ++v)
            for (int t = 0; t < 3; ++t) t_of_edge[ve1[var_ptr[v] + t]] = t;
        natural();
    }

    // natural labelling: check c at position c, variable v at position v, edges in CSR order
    void natural()
    {
        cpos.resize(m); vpos.resize(n); cat.assign(CPL * 32, -1); vat.assign(VPL * 32, -1); ks.resize((size_t)m * RW);
        for (int c = 0; c < m; ++c) { cpos[c] = c; cat[c] = c; for (int k = 0; k < RW; ++k) ks[(size_t)c * RW + k] = k; }
        for (int v = 0; v < n; ++v) { vpos[v] = v; vat[v] = v; }
        cost_natural = total_cost();
    }

    // explicit labelling; false (state unchanged) if the arrays are not permutations of the right shape
    bool set(const int32_t *check_pos, const int32_t *var_pos, const int32_t *kslot)
    {
        std::vector<int> na(CPL * 32, -1), nv(VPL * 32, -1);
        for (int c = 0; c < m; ++c) { if (check_pos[c] < 0 || check_pos[c] >= CPL * 32 || na[check_pos[c]] >= 0) return false; na[check_pos[c]] = c; }
        for (int v = 0; v < n; ++v) { if (var_pos[v] < 0 || var_pos[v] >= VPL * 32 || nv[var_pos[v]] >= 0) return false; nv[var_pos[v]] = v; }
        for (int c = 0; c < m; ++c) {
            unsigned seen = 0;
            for (int k = 0; k < RW; ++k) { const int x = kslot[(size_t)c * RW + k]; if (x < 0 || x >= RW || (seen >> x) & 1u) return false; seen |= 1u << x; }
        }
        cpos.assign(check_pos, check_pos + m); vpos.assign(var_pos, var_pos + n); ks.assign(kslot, kslot + (size_t)m * RW);
        cat = na; vat = nv;
        return true;
    }

    int cost() const { return total_cost(); }
    int cost_of_natural() const { return cost_natural; }
    int floor_cost() const { return NI; }
    const std::vector<int> &check_positions() const { return cpos; }
    const std::vector<int> &var_positions() const { return vpos; }
    const std::vector<int> &kslots() const { return ks; }

    // simulated annealing from the current labelling; keeps the best labelling seen (never worse than the start)
    void anneal(long long steps)
    {
        smooth = true;
        icost.resize(NI);
        for (int id = 0; id < NI; ++id) icost[id] = instr_cost(id);
        long long cur = 0;
        for (int id = 0; id < NI; ++id) cur += icost[id];
        smooth = false;
        int best = total_cost();
        smooth = true;
        std::vector<int> bc = cpos, bv = vpos, bk = ks;
        std::vector<int> stamp(NI, -1), touched, saved;
        kslot_of_edge.assign(col_idx.size(), 0);
        for (int c = 0; c < m; ++c) for (int k = 0; k < RW; ++k) kslot_of_edge[row_ptr[c] + ks[(size_t)c * RW + k]] = k;
        for (long long s = 0; s < steps && best > NI; ++s) {
            const double T = t0 * std::pow(t1 / t0, (double)s / (double)steps);
            const uint32_t kind = rnd() % 3;
            int a = 0, b = 0, c = 0;
            touched.clear();
            auto touch = [&](int id) { if (stamp[id] != (int)(s & 0x7fffffff)) { stamp[id] = (int)(s & 0x7fffffff); touched.push_back(id); } };
            auto touch_check = [&](int chk) {             // variable-pass instructions reading the messages of this check
                if (chk < 0) return;
                for (int e = row_ptr[chk]; e < row_ptr[chk + 1]; ++e) touch((vpos[col_idx[e]] / 32) * 3 + t_of_edge[e]);
            };
            auto touch_var = [&](int var) {               // message-update instructions reading the posterior of this variable
                if (var < 0) return;
                for (int q = var_ptr[var]; q < var_ptr[var + 1]; ++q) { const int e = ve1[q]; touch(VPL * 3 + (cpos[edge_check[e]] / 32) * RW + kslot_of_edge[e]); }
            };
            if (kind == 0) {                      // swap two check positions (one may be padding)
                a = rnd() % (CPL * 32); b = rnd() % (CPL * 32);
                if (a == b || (cat[a] < 0 && cat[b] < 0)) continue;
                for (int k = 0; k < RW; ++k) { touch(VPL * 3 + (a / 32) * RW + k); touch(VPL * 3 + (b / 32) * RW + k); }
                touch_check(cat[a]); touch_check(cat[b]);
                swap_pos(cat, cpos, a, b);
            } else if (kind == 1) {               // swap two variable positions
                a = rnd() % (VPL * 32); b = rnd() % (VPL * 32);
                if (a == b || (vat[a] < 0 && vat[b] < 0)) continue;
                for (int t = 0; t < 3; ++t) { touch((a / 32) * 3 + t); touch((b / 32) * 3 + t); }
                touch_var(vat[a]); touch_var(vat[b]);
                swap_pos(vat, vpos, a, b);
            } else {                              // swap two register slots of a check
                c = rnd() % m; a = rnd() % RW; b = rnd() % RW;
                if (a == b) continue;
                touch(VPL * 3 + (cpos[c] / 32) * RW + a); touch(VPL * 3 + (cpos[c] / 32) * RW + b);
                swap_k(c, a, b);
            }
            saved.resize(touched.size());
            long long nc = cur;
            for (size_t q = 0; q < touched.size(); ++q) {
                saved[q] = icost[touched[q]];
                icost[touched[q]] = instr_cost(touched[q]);
                nc += icost[touched[q]] - saved[q];
            }
            const double u = (rnd() + 1.0) / 4294967297.0;
            if (nc <= cur || u < std::exp((double)(cur - nc) / T)) {
                const bool improved = nc < cur;
                cur = nc;
                if (improved) {
                    int w = 0;
                    for (int id = 0; id < NI; ++id) w += icost[id] / 64;
                    if (w < best) { best = w; bc = cpos; bv = vpos; bk = ks; }
                }
            } else {                              // undo
                if (kind == 0) swap_pos(cat, cpos, a, b);
                else if (kind == 1) swap_pos(vat, vpos, a, b);
                else swap_k(c, a, b);
                for (size_t q = 0; q < touched.size(); ++q) icost[touched[q]] = saved[q];
            }
        }
        smooth = false;
        cpos = bc; vpos = bv; ks = bk;
        std::fill(cat.begin(), cat.end(), -1);
        std::fill(vat.begin(), vat.end(), -1);
        for (int c = 0; c < m; ++c) cat[cpos[c]] = c;
        for (int v = 0; v < n; ++v) vat[vpos[v]] = v;
    }

    WarpLayout tables() const
    {
        WarpLayout L;
        L.CPL = CPL; L.VPL = VPL; L.RW = RW;
        L.cost_natural = cost_natural; L.cost = total_cost(); L.floor = NI;
        L.ridx.assign((size_t)VPL * 3 * 32, 0u);
        L.ridx0.assign((size_t)VPL * 3 * 32, 0u);
        L.vidx.assign((size_t)CPL * RW * 32, 0u);
        L.cinfo.assign((size_t)CPL * 32, 0xffffffffu);
        L.vorig.assign((size_t)VPL * 32, 0xffffffffu);
        L.vpos.assign((size_t)VPL * 32, 0u);
        std::vector<int> slot_of_edge(col_idx.size(), 0);             // register slot of CSR edge e within its check
        for (int c = 0; c < m; ++c)
            for (int k = 0; k < RW; ++k) slot_of_edge[row_ptr[c] + ks[(size_t)c * RW + k]] = k;
        for (int i = 0; i < VPL; ++i)
            for (int l = 0; l < 32; ++l) {
                const int v = vat[i * 32 + l];
                if (v >= 0) L.vorig[(size_t)i * 32 + l] = (uint32_t)v;
                for (int t = 0; t < 3; ++t)
                    for (int which = 0; which < 2; ++which) {
                        uint32_t word = (uint32_t)(CPL * RW) * 32 + l;                           // the zero row
                        if (v >= 0) {
                            const int e = (which ? ve1 : ve0)[var_ptr[v] + t], c = edge_check[e];
                            word = (uint32_t)((cpos[c] / 32) * RW + slot_of_edge[e]) * 32 + (cpos[c] % 32);
                        }
                        (which ? L.ridx : L.ridx0)[(size_t)(i * 3 + t) * 32 + l] = 4u * word;
                    }
                const int vn = i * 32 + l;                                                        // natural index
                if (vn < n) L.vpos[(size_t)i * 32 + l] = 4u * (uint32_t)vpos[vn];
            }
        for (int i = 0; i < CPL; ++i)
            for (int l = 0; l < 32; ++l) {
                const int c = cat[i * 32 + l];
                if (c >= 0) L.cinfo[(size_t)i * 32 + l] = (uint32_t)c;
                for (int k = 0; k < RW; ++k)
                    L.vidx[(size_t)(i * RW + k) * 32 + l] = (c >= 0) ? 4u * (uint32_t)vpos[col_idx[row_ptr[c] + ks[(size_t)c * RW + k]]] : 4u * (uint32_t)l;
            }
        return L;
    }

private:
    int m, n, RW, CPL = 0, VPL = 0, NI = 0, cost_natural = 0;
    std::vector<int32_t> row_ptr, col_idx, var_ptr, ve0, ve1, edge_check;
    std::vector<int> cpos, vpos, cat, vat, ks, t_of_edge, kslot_of_edge, icost;
    uint64_t rng = 0x9e3779b97f4a7c15ull;
    bool smooth = false;

    uint32_t rnd()
    {
        rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17;
        return (uint32_t)(rng >> 32);
    }

    void swap_k(int c, int a, int b)
    {
        std::swap(ks[(size_t)c * RW + a], ks[(size_t)c * RW + b]);
        if (!kslot_of_edge.empty()) {
            kslot_of_edge[row_ptr[c] + ks[(size_t)c * RW + a]] = a;
            kslot_of_edge[row_ptr[c] + ks[(size_t)c * RW + b]] = b;
        }
    }

    // wavefronts of one instruction, plus (smooth) a small term that rewards flatter bank loads so that the search has
    // a slope to follow on the plateaus of the max
    int score(int mx, const int *cnt) const
    {
        if (!smooth) return mx;
        int sq = 0;
        for (int b = 0; b < 32; ++b) sq += cnt[b] * cnt[b];
        return 64 * mx + std::max(0, std::min(63, sq - 32));
    }

    int instr_cost(int id) const { return id < VPL * 3 ? var_instr_cost(id / 3, id % 3) : check_instr_cost((id - VPL * 3) / RW, (id - VPL * 3) % RW); }

    int var_instr_cost(int i, int t) const
    {
        int cnt[32] = {0}, mx = 1;
        for (int l = 0; l < 32; ++l) {
            const int v = vat[i * 32 + l];
            const int b = (v < 0) ? l : (cpos[edge_check[ve1[var_ptr[v] + t]]] & 31);    // padding reads the zero row, own bank
            mx = std::max(mx, ++cnt[b]);
        }
        return score(mx, cnt);
    }

    int check_instr_cost(int i, int k) const
    {
        int words[32][32], cnt[32] = {0}, mx = 1;
        for (int l = 0; l < 32; ++l) {
            const int c = cat[i * 32 + l];
            const int w = (c < 0) ? l : vpos[col_idx[row_ptr[c] + ks[(size_t)c * RW + k]]];   // padding: own bank
            const int b = w & 31;
            bool seen = false;
            for (int j = 0; j < cnt[b]; ++j) if (words[b][j] == w) { seen = true; break; }
            if (!seen) {
                words[b][cnt[b]] = w;
                mx = std::max(mx, ++cnt[b]);
            }
        }
        return score(mx, cnt);
    }

    int total_cost() const
    {
        int t = 0;
        for (int id = 0; id < NI; ++id) t += instr_cost(id);
        return t;
    }

    static void swap_pos(std::vector<int> &at, std::vector<int> &pos, int a, int b)
    {
        std::swap(at[a], at[b]);
        if (at[a] >= 0) pos[at[a]] = a;
        if (at[b] >= 0) pos[at[b]] = b;
    }
};

}  // namespace qldpc
