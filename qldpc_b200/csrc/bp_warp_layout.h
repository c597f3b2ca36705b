// Host side of the warp-per-shot BP kernel (bp_warp_kernel.cuh): which lane owns which check / variable, in which
// register slot each edge of a check sits, and the scatter / gather tables that follow from that choice.
//
// The kernel's shared-memory traffic per shot-iteration is
//   * CPL*RW scatter stores: the lane that owns a check writes the message of edge slot k straight into the column of
//     the destination variable (plane t = position of the edge in the variable's addition order) -> bank = lane of
//     the variable;
//   * 3*VPL + VPL conflict-free accesses by the variable's owner (its own column);
//   * CPL*RW gathers: the check's owner reads the posterior of the variable of edge slot k -> bank = lane of the
//     variable again.
// Both conflict patterns are those of the instruction (check slot i, edge slot k): the variables it touches must sit
// in 32 different lanes.  The labelling is free (the arithmetic of a check does not depend on the order of its edges;
// the addition order of a variable is carried by the plane index), and a conflict-free one exists whenever no lane
// receives more than RW edges from one check slot (Koenig: a bipartite multigraph of maximum degree RW splits into RW
// matchings).  construct() builds it:
//   1. checks go to CPL balanced slots (72 checks -> 24 + 24 + 24, not 32 + 32 + 8: the spare lanes are the slack);
//   2. variables are moved between lanes until every (check slot, lane) pair carries at most RW edges
//      (deterministic local search on the overflow, a few thousand cheap steps);
//   3. the edges of every check slot are coloured with RW colours (alternating-path edge colouring) = the register
//      slot of each edge; the padding lanes of a slot are pointed at the banks a round leaves unused.
// lanes = 16 (float64 kernel, 64-bit shared-memory words): a 64-bit access is served one HALF-warp at a time and two lanes of a
// half collide when their columns are equal mod 16 (same pair of banks).  The same construction with 16-lane conflict domains
// -- a "slot" is then half a kernel slot, positions stay slot * lanes + lane -- makes every half-warp access conflict-free.
#pragma once
#include <stdint.h>

#include <algorithm>
#include <cmath>
#include <vector>

namespace qldpc {

struct WarpLayout {
    int CPL = 0, VPL = 0, RW = 0;
    int cost_natural = 0, cost = 0, floor = 0;  // scatter + gather wavefronts per shot-iteration
    // device tables
    std::vector<uint32_t> sidx, sidx0, vidx;    // see bp_warp_kernel.cuh
    std::vector<uint32_t> cinfo;                // [CPL][32] original check index of the position, 0xffffffff = padding
    std::vector<uint32_t> vorig;                // [VPL][32] original variable index of the position, 0xffffffff = padding
    std::vector<uint32_t> vpos;                 // [ceil(n/32)][32] byte offset in the posterior buffer of variable 32 i + lane
};

class WarpLayoutBuilder {
public:
    WarpLayoutBuilder(int m, int n, const int32_t *row_ptr, const int32_t *col_idx, const int32_t *var_ptr, const int32_t *var_edge0,
                      const int32_t *var_edge1, const int32_t *edge_check, int RW, int check_slots = 0, int var_slots = 0, int lanes = 32)
        : m(m), n(n), RW(RW), W(lanes), row_ptr(row_ptr, row_ptr + m + 1), col_idx(col_idx, col_idx + row_ptr[m]), var_ptr(var_ptr, var_ptr + n + 1),
          ve0(var_edge0, var_edge0 + row_ptr[m]), ve1(var_edge1, var_edge1 + row_ptr[m]), edge_check(edge_check, edge_check + row_ptr[m])
    {
        // rows may have fewer than RW edges (padding edge slots); more slots than ceil(m/32) / ceil(n/32) may be asked for
        // (W = 16: conflict domains are the half-warps, a slot is half a kernel slot -- see the header comment)
        CPL = std::max((m + 31) / 32, check_slots) * (32 / W);
        VPL = std::max((n + 31) / 32, var_slots) * (32 / W);
        NI = CPL * RW;
        natural();
        cost_natural = total_cost();
    }

    // natural labelling: check c at position c, variable v at position v, edges in CSR order
    void natural()
    {
        cpos.resize(m); vpos.resize(n); cat.assign(CPL * W, -1); vat.assign(VPL * W, -1); ks.resize((size_t)m * RW);
        for (int c = 0; c < m; ++c) {
            cpos[c] = c; cat[c] = c;
            for (int k = 0; k < RW; ++k) ks[(size_t)c * RW + k] = (k < row_ptr[c + 1] - row_ptr[c]) ? k : -1;     // -1: padding edge slot
        }
        for (int v = 0; v < n; ++v) { vpos[v] = v; vat[v] = v; }
        padbank.assign((size_t)NI * W, -1);
    }

    int cost() const { return total_cost(); }
    int cost_of_natural() const { return cost_natural; }
    int floor_cost() const { return 2 * NI; }

    // the labelling described above; true if it is conflict-free, false if step 2 did not get every (check slot, lane)
    // degree down to RW within `steps` moves (the overflowing edges then share a bank with another edge of their round)
    bool construct(long long steps = 8000000)
    {
        // 1. balanced check slots, natural order inside
        std::fill(cat.begin(), cat.end(), -1);
        for (int i = 0, c = 0; i < CPL; ++i) {
            const int cnt = m / CPL + (i < m % CPL ? 1 : 0);
            for (int l = 0; l < cnt; ++l, ++c) { cpos[c] = i * W + l; cat[i * W + l] = c; }
        }
        std::fill(vat.begin(), vat.end(), -1);
        for (int v = 0; v < n; ++v) {                                  // spread evenly over all variable slots
            const int p = (int)(((long long)v * VPL * W) / n);
            vpos[v] = p; vat[p] = v;
        }
        // 2. degree of (check slot, lane)
        std::vector<int> deg((size_t)CPL * W, 0);
        auto add_var = [&](int v, int lane, int sgn) {
            for (int q = var_ptr[v]; q < var_ptr[v + 1]; ++q) deg[(size_t)(cpos[edge_check[ve1[q]]] / W) * W + lane] += sgn;
        };
        auto lane_over = [&](int lane) {          // overflow of a lane, plus a small term that prefers flat loads
            long long o = 0;
            for (int i = 0; i < CPL; ++i) { const int d = deg[(size_t)i * W + lane]; o += 1024ll * std::max(0, d - RW) + (long long)d * d; }
            return o;
        };
        auto overflow = [&]() {
            long long o = 0;
            for (size_t q = 0; q < deg.size(); ++q) o += std::max(0, deg[q] - RW);
            return o;
        };
        for (int v = 0; v < n; ++v) add_var(v, vpos[v] % W, +1);
        long long of = overflow();
        for (long long s = 0; s < steps && of > 0; ++s) {
            // a: a variable that feeds an overflowing (check slot, lane) pair, found by a few random probes; b: anywhere
            int a = rnd() % (VPL * W);
            for (int probe = 0; probe < 16; ++probe) {
                const int q = rnd() % (int)deg.size();
                if (deg[q] <= RW) continue;
                const int cand = (rnd() % VPL) * W + (q % W);
                const int v = vat[cand];
                if (v < 0) continue;
                bool feeds = false;
                for (int e = var_ptr[v]; e < var_ptr[v + 1]; ++e) feeds = feeds || (cpos[edge_check[ve1[e]]] / W == q / W);
                if (feeds) { a = cand; break; }
            }
            const int b = rnd() % (VPL * W);
            const int la = a % W, lb = b % W;
            if (la == lb || (vat[a] < 0 && vat[b] < 0)) continue;
            const long long before = lane_over(la) + lane_over(lb);
            if (vat[a] >= 0) { add_var(vat[a], la, -1); add_var(vat[a], lb, +1); }
            if (vat[b] >= 0) { add_var(vat[b], lb, -1); add_var(vat[b], la, +1); }
            const long long after = lane_over(la) + lane_over(lb);
            // accept improvements and sideways moves; a rare uphill move (1 in 64) keeps the search from stalling
            if (after <= before || (rnd() & 63) == 0) {
                swap_pos(vat, vpos, a, b);
                of = overflow();
            } else {
                if (vat[a] >= 0) { add_var(vat[a], lb, -1); add_var(vat[a], la, +1); }
                if (vat[b] >= 0) { add_var(vat[b], la, -1); add_var(vat[b], lb, +1); }
            }
        }
        // 3. edge colouring of each check slot: vertices = checks of the slot and lanes, colours = register slots
        padbank.assign((size_t)NI * W, -1);
        std::vector<int> colour(col_idx.size(), -1);
        auto lane_of = [&](int e) { return vpos[col_idx[e]] % W; };
        auto chk_of = [&](int e) { return cpos[edge_check[e]] % W; };
        for (int i = 0; i < CPL; ++i) {
            std::vector<int> at_check((size_t)W * RW, -1), at_lane((size_t)W * RW, -1);      // [vertex][colour] -> edge
            for (int l = 0; l < W; ++l) {
                const int c = cat[i * W + l];
                if (c < 0) continue;
                for (int e = row_ptr[c]; e < row_ptr[c + 1]; ++e) {
                    const int u = l, w = lane_of(e);
                    int fa = -1, fb = -1;
                    for (int k = 0; k < RW; ++k) {
                        if (fa < 0 && at_check[(size_t)u * RW + k] < 0) fa = k;
                        if (fb < 0 && at_lane[(size_t)w * RW + k] < 0) fb = k;
                    }
                    if (fb < 0) {                                                 // overflowing lane: accept the conflict
                        colour[e] = fa;
                        at_check[(size_t)u * RW + fa] = e;
                        continue;
                    }
                    if (at_lane[(size_t)w * RW + fa] >= 0) {
                        // fa is free at the check but taken at the lane: flip the fa/fb alternating path that starts at
                        // the lane (bipartite: it cannot come back to this check)
                        std::vector<int> path;
                        int cur = w, col = fa;
                        bool on_lane = true;
                        bool broken = false;
                        while (true) {
                            const int pe = on_lane ? at_lane[(size_t)cur * RW + col] : at_check[(size_t)cur * RW + col];
                            if (pe < 0) break;
                            if (path.size() > 128) { broken = true; break; }       // only after an overflowing lane spoiled the colouring
                            path.push_back(pe);
                            cur = on_lane ? chk_of(pe) : lane_of(pe);
                            on_lane = !on_lane;
                            col = (col == fa) ? fb : fa;
                        }
                        if (broken) {                                                // accept the conflict
                            colour[e] = fa;
                            at_check[(size_t)u * RW + fa] = e;
                            continue;
                        }
                        for (int pe : path) { at_check[(size_t)chk_of(pe) * RW + colour[pe]] = -1; at_lane[(size_t)lane_of(pe) * RW + colour[pe]] = -1; }
                        for (int pe : path) {
                            colour[pe] = (colour[pe] == fa) ? fb : fa;
                            at_check[(size_t)chk_of(pe) * RW + colour[pe]] = pe;
                            at_lane[(size_t)lane_of(pe) * RW + colour[pe]] = pe;
                        }
                    }
                    colour[e] = fa;
                    at_check[(size_t)u * RW + fa] = e;
                    at_lane[(size_t)w * RW + fa] = e;
                }
            }
            for (int l = 0; l < W; ++l) {
                const int c = cat[i * W + l];
                if (c < 0) continue;
                for (int k = 0; k < RW; ++k) ks[(size_t)c * RW + k] = -1;
                for (int e = row_ptr[c]; e < row_ptr[c + 1]; ++e) ks[(size_t)c * RW + colour[e]] = e - row_ptr[c];
            }
            // padding (a lane without a check, or an edge slot its check does not use): one of the banks round k leaves unused
            for (int k = 0; k < RW; ++k) {
                int nb = 0;
                for (int l = 0; l < W; ++l) {
                    const int c = cat[i * W + l];
                    if (c >= 0 && ks[(size_t)c * RW + k] >= 0) continue;
                    while (nb < W && at_lane[(size_t)nb * RW + k] >= 0) ++nb;
                    padbank[(size_t)(i * RW + k) * W + l] = (nb < W) ? nb++ : l;
                }
            }
        }
        return of == 0;
    }

    WarpLayout tables() const
    {
        WarpLayout L;
        // kernel coordinates: position = slot * 32 + lane, whatever the width of the conflict domains
        const int CPLk = CPL * W / 32, VPLk = VPL * W / 32;
        L.CPL = CPLk; L.VPL = VPLk; L.RW = RW;
        L.cost_natural = cost_natural; L.cost = total_cost(); L.floor = 2 * NI;
        L.sidx.assign((size_t)CPLk * RW * 32, 0u);
        L.sidx0.assign((size_t)CPLk * RW * 32, 0u);
        L.vidx.assign((size_t)CPLk * RW * 32, 0u);
        L.cinfo.assign((size_t)CPLk * 32, 0xffffffffu);
        L.vorig.assign((size_t)VPLk * 32, 0xffffffffu);
        L.vpos.assign((size_t)((n + 31) / 32) * 32, 0u);
        std::vector<int> t0_of_edge(col_idx.size(), 0), t1_of_edge(col_idx.size(), 0);   // position of edge e in its variable's addition order
        for (int v = 0; v < n; ++v)
            for (int t = 0; t < var_ptr[v + 1] - var_ptr[v]; ++t) { t0_of_edge[ve0[var_ptr[v] + t]] = t; t1_of_edge[ve1[var_ptr[v] + t]] = t; }
        for (int i = 0; i < VPLk; ++i)
            for (int l = 0; l < 32; ++l) {
                const int v = vat[i * 32 + l];
                if (v >= 0) L.vorig[(size_t)i * 32 + l] = (uint32_t)v;
            }
        for (int vn = 0; vn < n; ++vn) L.vpos[vn] = 4u * (uint32_t)vpos[vn];                        // by index in H
        for (int i = 0; i < CPLk; ++i)
            for (int l = 0; l < 32; ++l) {
                const int c = cat[i * 32 + l];
                if (c >= 0) L.cinfo[(size_t)i * 32 + l] = (uint32_t)c;
                for (int k = 0; k < RW; ++k) {
                    const size_t at = (size_t)(i * RW + k) * 32 + l;
                    if (c < 0 || ks[(size_t)c * RW + k] < 0) {
                        // padding: reads the +inf row of the posterior buffer, delivers into the dump row (chosen bank of
                        // the lane's own conflict domain)
                        const int b = (l / W) * W + pad_bank((i * 32 + l) / W, k, l % W);
                        L.vidx[at] = 4u * (uint32_t)(VPLk * 32 + b);
                        L.sidx[at] = L.sidx0[at] = 4u * (uint32_t)(3 * VPLk * 32 + b);
                        continue;
                    }
                    const int e = row_ptr[c] + ks[(size_t)c * RW + k], p = vpos[col_idx[e]];
                    L.vidx[at] = 4u * (uint32_t)p;
                    L.sidx[at] = 4u * (uint32_t)(t1_of_edge[e] * VPLk * 32 + p);
                    L.sidx0[at] = 4u * (uint32_t)(t0_of_edge[e] * VPLk * 32 + p);
                }
            }
        return L;
    }

private:
    int m, n, RW, W, CPL = 0, VPL = 0, NI = 0, cost_natural = 0;      // CPL, VPL: slots of W lanes
    std::vector<int32_t> row_ptr, col_idx, var_ptr, ve0, ve1, edge_check;
    std::vector<int> cpos, vpos, cat, vat, ks, padbank;
    uint64_t rng = 0x9e3779b97f4a7c15ull;

    uint32_t rnd()
    {
        rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17;
        return (uint32_t)(rng >> 32);
    }

    int pad_bank(int i, int k, int l) const
    {
        const int b = padbank[(size_t)(i * RW + k) * W + l];
        return b < 0 ? l : b;
    }

    // instruction (check slot i, edge slot k): wavefronts of the scatter store (every edge is its own word) plus
    // wavefronts of the gather (edges to the same variable read the same word)
    int instr_cost(int id) const
    {
        const int i = id / RW, k = id % RW;
        int words[32][32], cnt[32] = {0}, scnt[32] = {0}, mx = 1, smx = 1;
        for (int l = 0; l < W; ++l) {
            const int c = cat[i * W + l];
            const int w = (c < 0 || ks[(size_t)c * RW + k] < 0) ? pad_bank(i, k, l) : vpos[col_idx[row_ptr[c] + ks[(size_t)c * RW + k]]];
            const int b = w % W;
            smx = std::max(smx, ++scnt[b]);
            bool seen = false;
            for (int j = 0; j < cnt[b]; ++j) if (words[b][j] == w) { seen = true; break; }
            if (!seen) {
                words[b][cnt[b]] = w;
                mx = std::max(mx, ++cnt[b]);
            }
        }
        return mx + smx;
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
