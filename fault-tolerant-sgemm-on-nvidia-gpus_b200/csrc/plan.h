// plan.h -- host-side work planner for the persistent fused ABFT-SGEMM kernel.
//
// The reference launches one CTA per output tile and lets the hardware scheduler deal them out
// (/root/reference/kernel/ft_sgemm/sgemm.cu:110-199: grid = (M/ms, N/ns)).  A persistent tcgen05 kernel with one CTA
// (pair) per SM has to do that job itself, and with 256x256 tiles a 4096^3 problem is only 256 tiles on 74 CTA pairs:
// dealing whole tiles round-robin leaves 13.5 % of the machine idle in the last wave (and the ABFT checksum
// tile-columns make it 4 waves instead of 3.68).  The planner therefore
//   * orders the work as [checksum tiles][whole data tiles in raster order][split-K tail: the last H data tiles cut
//     into S k-slices, slice-major],
//   * assigns items in that order to the least-loaded unit (list scheduling with item costs in "tile-times"; checksum
//     tiles cost their narrowed width, slices cost 1/S plus a measured fold-in overhead),
//   * picks (H, S) from a small candidate set by simulated makespan.
// Every unit's list is increasing in the global order and every dependency points to an earlier item, which is what
// makes the in-kernel waits deadlock-free (see SegIter in ftsgemm_kernel.cuh, tests/test_schedule.py).
#pragma once
#include <algorithm>
#include <cstdint>
#include <queue>
#include <vector>

namespace ftsgemm {

struct PlanItem {
  int tile;        // decode order: checksum tiles first, then data tiles
  int kb_begin, kb_end;
  int kind;        // 0 whole tile, 1 split-K contributor, 2 split-K finisher
  int slice;
  int split_idx;   // index among the split tiles (workspace slot), -1 otherwise
};

struct Plan {
  int units = 0;
  int sk_tiles = 0;   // H: number of data tiles in the split-K tail
  int sk_slices = 1;  // S
  double makespan = 0.0;
  std::vector<int> offsets;          // units + 1
  std::vector<PlanItem> items;       // grouped by unit, in execution order
};

struct PlanInput {
  int units;            // CTAs or CTA pairs
  int n_chk_tiles;      // checksum tile-columns x tiles_m (first tiles in decode order)
  int n_data_tiles;
  int num_kb;           // k-blocks per tile
  int tiles_m;          // checksum tile t belongs to checksum tile-column t / tiles_m
  std::vector<double> chk_col_cost;  // tile-times of one tile of each checksum tile-column (narrowed UMMA N / BN)
  double chk_release = 0.0;          // tile-times before checksum items can start (encode pre-pass running concurrently)
  double slice_overhead;// tile-times added to every split item (partial-sum round trip)
  int max_slices;       // 1 disables the tail
  int force_slices;     // > 1: use exactly this S on the best H (tests)
  size_t slab_bytes;    // bytes of one partial accumulator tile x CTAs per unit (workspace sizing)
};

namespace plan_detail {

struct Cand {
  int H, S;
};

// list scheduling in global item order; returns makespan, optionally records the assignment.
// head_first: [checksum][split slices, slice-major][whole tiles] -- the finishers' fold-in is hidden behind the whole
// tiles that follow (cheaper per item) but small-items-first levels worse; otherwise [checksum][whole][split slices].
inline double schedule(const PlanInput &in, int H, int S, bool head_first, Plan *out) {
  typedef std::pair<double, int> LU;  // (load, unit): least load first, ties to the lowest unit id
  std::priority_queue<LU, std::vector<LU>, std::greater<LU>> pq;
  for (int u = 0; u < in.units; ++u) pq.push(LU(0.0, u));
  std::vector<std::vector<PlanItem>> lists;
  if (out) lists.resize(in.units);
  double makespan = 0.0;
  auto give = [&](const PlanItem &it, double cost, double release = 0.0) {
    LU lu = pq.top();
    pq.pop();
    if (lu.first < release) lu.first = release;  // the unit idles until the item's input exists
    lu.first += cost;
    if (lu.first > makespan) makespan = lu.first;
    if (out) lists[lu.second].push_back(it);
    pq.push(lu);
  };
  const int whole = in.n_data_tiles - H;
  // the split tiles are always the LAST H data tiles of the raster; only their position in the item order changes
  auto give_whole = [&]() {
    for (int d = 0; d < whole; ++d) give(PlanItem{in.n_chk_tiles + d, 0, in.num_kb, 0, 0, -1}, 1.0);
  };
  auto give_split = [&]() {
    // the fold-in is hidden only if a full wave of whole tiles follows the split items
    const double ovh = S > 1 ? in.slice_overhead * ((head_first && whole >= in.units) ? 0.35 : 1.0) : 0.0;
    for (int s = 0; s < S; ++s) {
      const int kb0 = static_cast<int>(static_cast<long long>(in.num_kb) * s / S);
      const int kb1 = static_cast<int>(static_cast<long long>(in.num_kb) * (s + 1) / S);
      for (int i = 0; i < H; ++i) {
        const int kind = (S == 1) ? 0 : (s == S - 1 ? 2 : 1);
        give(PlanItem{in.n_chk_tiles + whole + i, kb0, kb1, kind, s, S > 1 ? i : -1},
             static_cast<double>(kb1 - kb0) / in.num_kb + ovh);
      }
    }
  };
  for (int t = 0; t < in.n_chk_tiles; ++t)
    give(PlanItem{t, 0, in.num_kb, 0, 0, -1}, in.chk_col_cost[static_cast<size_t>(t / in.tiles_m)], in.chk_release);
  if (head_first) {
    give_split();
    give_whole();
  } else {
    give_whole();
    give_split();
  }
  if (out) {
    out->units = in.units;
    out->sk_tiles = S > 1 ? H : 0;
    out->sk_slices = S > 1 ? S : 1;
    out->makespan = makespan;
    out->offsets.assign(1, 0);
    out->items.clear();
    for (int u = 0; u < in.units; ++u) {
      out->items.insert(out->items.end(), lists[u].begin(), lists[u].end());
      out->offsets.push_back(static_cast<int>(out->items.size()));
    }
  }
  return makespan;
}

}  // namespace plan_detail

inline Plan build_plan(const PlanInput &in) {
  using plan_detail::schedule;
  int bestH = 0, bestS = 1;
  bool best_head = false;
  double best = schedule(in, 0, 1, false, nullptr);
  const int P = in.units, T = in.n_data_tiles;
  if (in.max_slices > 1 && T > 0) {
    const int hs[] = {T % P, T % P + P, P / 2, P, (3 * P) / 2, 2 * P, T};
    for (int hi = 0; hi < 7; ++hi) {
      const int H = std::min(hs[hi], T);
      if (H <= 0) continue;
      for (int S = 2; S <= in.max_slices; ++S) {
        if (in.num_kb / S < 4) break;  // keep slices at least 4 k-blocks long
        if (in.force_slices > 1 && S != in.force_slices) continue;
        if (static_cast<size_t>(H) * (S - 1) * in.slab_bytes > (static_cast<size_t>(256) << 20)) break;
        if (static_cast<size_t>(H) * (S - 1) * 8 * sizeof(int) > 65536) break;  // flag area
        for (int head = 0; head < 2; ++head) {
          const double t = schedule(in, H, S, head != 0, nullptr);
          const bool forced_first = in.force_slices > 1 && bestS == 1;
          if (t < best * 0.985 || forced_first) {  // a split must buy at least 1.5 %
            best = t;
            bestH = H;
            bestS = S;
            best_head = head != 0;
          }
        }
      }
    }
  }
  Plan p;
  schedule(in, bestH, bestS, best_head, &p);
  return p;
}

}  // namespace ftsgemm
