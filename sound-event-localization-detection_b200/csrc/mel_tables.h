// Host-side construction of the mel gather tables (shared by abi.cu and tests/host_sim).
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <utility>
#include <vector>

namespace seld {

// ---- mel gather schedule --------------------------------------------------------------------------
// Lane l owns mel l (slot A) and mel n_mels-1-l (slot B, only mels >= 32): with HTK triangles the non-zero
// counts then add up to a roughly constant number per lane.  Each slot is a loop whose trip count is the
// longest list of that slot; shorter lists are padded with zero-weight entries.  The ORDER of a lane's
// entries is free (a sum), so it is chosen such that the 8 lanes of each quarter-warp — one LDS.128
// wavefront — hit 8 different rows modulo 8, i.e. all 32 banks: a padding entry (weight 0) is inserted when
// a lane has spare iterations and no conflict-free row left.
struct MelTables {
    int la = 0, lb = 0;
    std::vector<int2> entries;  // [(la+lb)][32]: {byte offset of bin row, weight bits}
    std::vector<int> idx;       // [2][32]
};

inline void schedule_slot(const std::vector<std::vector<std::pair<int, float>>>& lists /*[32]*/, int len,
                          std::vector<int2>& out /* appended: [len][32] */) {
    const size_t base = out.size();
    out.resize(base + (size_t)len * 32);
    for (int q = 0; q < 4; ++q) {  // quarter-warps are independent wavefronts
        std::vector<std::vector<std::pair<int, float>>> rem(8);
        for (int l = 0; l < 8; ++l) rem[l] = lists[q * 8 + l];
        for (int it = 0; it < len; ++it) {
            bool used[8] = {false, false, false, false, false, false, false, false};
            int chosen_row[8];
            float chosen_w[8];
            // lanes with the least slack choose first
            int order[8];
            for (int l = 0; l < 8; ++l) order[l] = l;
            std::sort(order, order + 8, [&](int x, int y) {
                int sx = (len - it) - (int)rem[x].size(), sy = (len - it) - (int)rem[y].size();
                return sx < sy;
            });
            for (int oi = 0; oi < 8; ++oi) {
                const int l = order[oi];
                auto& r = rem[l];
                const int slack = (len - it) - (int)r.size();
                int pick = -1;
                // prefer the free residue class in which this lane has most entries left
                int cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                for (auto& e : r) cnt[e.first & 7]++;
                int best = -1;
                for (size_t i = 0; i < r.size(); ++i) {
                    const int res = r[i].first & 7;
                    if (used[res]) continue;
                    if (best < 0 || cnt[res] > cnt[best]) { best = res; pick = (int)i; }
                }
                if (pick < 0 && !r.empty() && slack <= 0) pick = 0;  // forced conflict
                if (pick >= 0) {
                    chosen_row[l] = r[pick].first;
                    chosen_w[l] = r[pick].second;
                    used[chosen_row[l] & 7] = true;
                    r.erase(r.begin() + pick);
                } else {  // padding entry on a free bank group
                    int res = 0;
                    while (res < 8 && used[res]) ++res;
                    if (res == 8) res = 0;
                    used[res] = true;
                    chosen_row[l] = res;  // rows 0..7 always exist
                    chosen_w[l] = 0.f;
                }
            }
            for (int l = 0; l < 8; ++l) {
                int2 e;
                e.x = chosen_row[l] * 16;  // byte offset of the float4 row
                std::memcpy(&e.y, &chosen_w[l], 4);
                out[base + (size_t)it * 32 + q * 8 + l] = e;
            }
        }
    }
}

inline MelTables build_mel_tables(const float* fb, int n_bins, int n_mels) {
    MelTables t;
    t.idx.assign(64, -1);
    std::vector<std::vector<std::pair<int, float>>> A(32), Bs(32);
    for (int l = 0; l < 32; ++l) {
        const int ma = l < n_mels ? l : -1;
        const int mb = (n_mels - 1 - l >= 32) ? n_mels - 1 - l : -1;
        t.idx[l] = ma;
        t.idx[32 + l] = mb;
        for (int k = 0; k < n_bins; ++k) {
            if (ma >= 0 && fb[(size_t)k * n_mels + ma] != 0.f) A[l].push_back({k, fb[(size_t)k * n_mels + ma]});
            if (mb >= 0 && fb[(size_t)k * n_mels + mb] != 0.f) Bs[l].push_back({k, fb[(size_t)k * n_mels + mb]});
        }
        t.la = std::max(t.la, (int)A[l].size());
        t.lb = std::max(t.lb, (int)Bs[l].size());
    }
    t.la = (t.la + 3) & ~3;  // the kernel unrolls the gather by 4
    t.lb = (t.lb + 3) & ~3;
    schedule_slot(A, t.la, t.entries);
    schedule_slot(Bs, t.lb, t.entries);
    return t;
}


}  // namespace seld
