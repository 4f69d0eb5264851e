// Host-side graph plan for the staged EdgeConv kernel (edgeconv_tcgen05.cu).
//
// The reference builds its static kNN graph once per module in __init__ (checkerpose/model/pipeline.py:248,
// init.py:98) on farthest-point-sampled keypoints, whose order is spatially incoherent by construction: the
// 20 neighbours of 128 consecutive keypoints touch ~2000 distinct rows.  This routine, run once next to that
// knn() call, prepares everything the kernel needs that depends on the graph only:
//
//   1. renumbering: recursive coordinate bisection, so that every tile of 128 consecutive nodes is a compact
//      surface patch (its neighbour lists then touch ~245 distinct rows instead of ~2000);
//   2. per tile, the list of distinct neighbour rows ("ulist": what the kernel stages in shared memory);
//   3. per tile, 64 node PAIRS with their "program": neighbouring nodes share most of their neighbours (15 of 20
//      on the shipped clouds), so a pair is aggregated as  m = max(common rows);  a = max(m, rest of a);
//      b = max(m, rest of b)  -- 40 - C shared-memory row reads instead of 40.  Pairs are matched greedily by
//      overlap, sorted, and every 4 consecutive pairs (one warp of the kernel) use the same C (multiple of 4).
//
// Pure integer/geometry preprocessing on the host; nothing here is on the per-RoI path.
#include <algorithm>
#include <numeric>
#include <vector>

#include "common.cuh"

namespace {

constexpr int TILE = CP_PLAN_TILE;    // nodes per tile of the kernel = nodes per staging list
constexpr int PAIRS = CP_PLAN_PAIRS;  // pairs per tile

struct Rcb {
  const float* x;  // (3, N) coordinates of one graph
  int N;
  std::vector<int> order;

  void split(int* ids, int n, int lo) {
    if (n <= 1) {
      if (n == 1) order[lo] = ids[0];
      return;
    }
    float mn[3] = {1e30f, 1e30f, 1e30f}, mx[3] = {-1e30f, -1e30f, -1e30f};
    for (int i = 0; i < n; ++i)
      for (int a = 0; a < 3; ++a) {
        const float v = x[(size_t)a * N + ids[i]];
        mn[a] = std::min(mn[a], v);
        mx[a] = std::max(mx[a], v);
      }
    int ax = 0;
    for (int a = 1; a < 3; ++a)
      if (mx[a] - mn[a] > mx[ax] - mn[ax]) ax = a;
    // left part = a whole number of tiles while the box holds more than one tile, half of it below
    int left;
    if (n > TILE) {
      const int tiles = (n + TILE - 1) / TILE;
      left = ((tiles + 1) / 2) * TILE;
    } else {
      left = (n + 1) / 2;
    }
    const float* xa = x + (size_t)ax * N;
    std::nth_element(ids, ids + left, ids + n, [xa](int p, int q) { return xa[p] < xa[q] || (xa[p] == xa[q] && p < q); });
    split(ids, left, lo);
    split(ids + left, n - left, lo + left);
  }
};

struct Cand {
  int ov, a, b;
};

}  // namespace

extern "C" int cp_graph_plan_kp(int K) {
  static const int sizes[] = {8, 16, 20, 32, 40};
  for (int s : sizes)
    if (K >= 1 && K <= s) return s;
  return -1;
}

extern "C" int cp_graph_plan_build(const float* xyz, const int32_t* idx, int G, int N, int K, int umax, int32_t* perm,
                                   int32_t* idx_p, int32_t* ucount, uint16_t* ulist, uint16_t* prog) {
  CP_REQUIRE(idx && perm && idx_p && ucount && ulist && prog, CP_E_INVALID, "cp_graph_plan_build: null pointer");
  CP_REQUIRE(G > 0 && N > 0 && N < 65535 && K > 0 && umax > 0 && umax <= 512 && umax % (4 * CP_PLAN_LIST_LANES) == 0, CP_E_INVALID,
             "cp_graph_plan_build: bad sizes G=%d N=%d K=%d umax=%d (umax: multiple of 128, <= 512)", G, N, K, umax);
  const int KP = cp_graph_plan_kp(K);
  CP_REQUIRE(KP > 0, CP_E_UNSUPPORTED, "cp_graph_plan_build: K=%d > 40", K);
  const int T = (N + TILE - 1) / TILE;
  const int PW = 2 * KP + 8;
  const int LL = CP_PLAN_LIST_LANES, UI = umax / LL;
  int worst = 0;
  std::vector<int> ids(N), inv(N), stamp(N, -1), local(N), mark(N, 0), cmark(N, 0);
  int tick = 0;  // unique stamp per use of mark / cmark
  std::vector<Cand> cand;
  for (int g = 0; g < G; ++g) {
    int32_t* pg = perm + (size_t)g * N;
    if (xyz) {
      Rcb r{xyz + (size_t)g * 3 * N, N, std::vector<int>(N)};
      std::iota(ids.begin(), ids.end(), 0);
      r.split(ids.data(), N, 0);
      for (int i = 0; i < N; ++i) pg[i] = r.order[i];
    } else {
      for (int i = 0; i < N; ++i) pg[i] = i;  // no coordinates: keep the caller's numbering
    }
    for (int i = 0; i < N; ++i) inv[pg[i]] = i;
    const int32_t* ig = idx + (size_t)g * N * K;
    int32_t* ip = idx_p + (size_t)g * N * K;
    for (int i = 0; i < N; ++i)
      for (int k = 0; k < K; ++k) {
        const int j = ig[(size_t)pg[i] * K + k];
        CP_REQUIRE(j >= 0 && j < N, CP_E_INVALID, "cp_graph_plan_build: neighbour index %d outside [0,%d)", j, N);
        ip[(size_t)i * K + k] = inv[j];
      }
    std::fill(stamp.begin(), stamp.end(), -1);
    for (int t = 0; t < T; ++t) {
      const size_t gt = (size_t)g * T + t;
      const int n0 = t * TILE, n1 = std::min(N, n0 + TILE), nv = n1 - n0;
      // ---- distinct neighbour rows of the tile, ascending (sequential-ish source addresses for the copies) ----
      std::vector<int> u;
      for (int i = n0; i < n1; ++i)
        for (int k = 0; k < K; ++k) {
          const int j = ip[(size_t)i * K + k];
          if (stamp[j] != t) {
            stamp[j] = t;
            u.push_back(j);
          }
        }
      std::sort(u.begin(), u.end());
      const int U = (int)u.size();
      worst = std::max(worst, U);
      ucount[gt] = U;
      // kernel-friendly layout: quarter-warp q of the kernel copies list entries q, q+64, q+128, ...
      uint16_t* ul = ulist + gt * umax;
      for (int q = 0; q < LL; ++q)
        for (int i = 0; i < UI; ++i) {
          const int e = i * LL + q;
          ul[q * UI + i] = (uint16_t)(e < U ? u[e] : 0xFFFF);
        }
      for (int e = 0; e < U; ++e) local[u[e]] = e;
      auto off = [&](int row) { return (uint16_t)(std::min(local[row], umax - 1) * 128); };  // U > umax: plan unusable anyway

      // ---- pair matching: greedy by overlap of the neighbour sets ----
      cand.clear();
      for (int a = 0; a < nv; ++a) {
        ++tick;
        for (int k = 0; k < K; ++k) mark[ip[(size_t)(n0 + a) * K + k]] = tick;
        for (int b = a + 1; b < nv; ++b) {
          int ov = 0;
          for (int k = 0; k < K; ++k) ov += mark[ip[(size_t)(n0 + b) * K + k]] == tick;
          cand.push_back({ov, a, b});
        }
      }
      std::stable_sort(cand.begin(), cand.end(), [](const Cand& x, const Cand& y) { return x.ov > y.ov; });
      std::vector<char> used(TILE, 0);
      std::vector<Cand> pairs;
      for (const Cand& c : cand)
        if (!used[c.a] && !used[c.b]) {
          used[c.a] = used[c.b] = 1;
          pairs.push_back(c);
        }
      for (int a = 0; a < nv; ++a)
        if (!used[a]) pairs.push_back({0, a, 255});  // odd node out: no partner
      while ((int)pairs.size() < PAIRS) pairs.push_back({-1, 255, 255});  // nothing to do (ragged last tile)
      // pairs arrive sorted by overlap (greedy order); singles and empties last
      uint16_t* pt = prog + gt * PAIRS * PW;
      for (int w = 0; w < PAIRS / 4; ++w) {
        int cmin = KP;
        for (int q = 0; q < 4; ++q) cmin = std::min(cmin, std::max(pairs[w * 4 + q].ov, 0));
        const int C = cmin / 4 * 4;  // common rows taken by every pair of this warp
        for (int q = 0; q < 4; ++q) {
          const Cand& pr = pairs[w * 4 + q];
          uint16_t* e = pt + (size_t)(w * 4 + q) * PW;
          std::fill(e, e + PW, (uint16_t)0);
          e[2 * KP] = (uint16_t)((pr.a & 255) | ((pr.b & 255) << 8));
          e[2 * KP + 1] = (uint16_t)C;
          if (pr.a == 255) continue;  // empty slot: offsets 0 (always a staged row), results discarded
          const int32_t* na = ip + (size_t)(n0 + pr.a) * K;
          // a-list: C common rows first, then a's other neighbours, padded with its first entry
          int na_common = 0, pos = 0;
          std::vector<int> common;
          if (pr.b != 255) {
            const int32_t* nb = ip + (size_t)(n0 + pr.b) * K;
            ++tick;
            for (int k = 0; k < K; ++k) mark[nb[k]] = tick;
            for (int k = 0; k < K && na_common < C; ++k)
              if (mark[na[k]] == tick) {
                common.push_back(na[k]);
                ++na_common;
              }
          }
          // (a warp-uniform C never exceeds the pair's own overlap, so na_common == C for real pairs; a single
          //  node in a warp with C > 0 cannot happen because singles have overlap 0)
          ++tick;
          for (int r : common) cmark[r] = tick;
          for (int r : common) e[pos++] = off(r);
          for (int k = 0; k < K; ++k)
            if (cmark[na[k]] != tick) e[pos++] = off(na[k]);
          while (pos < KP) e[pos++] = e[0];
          // b-list: b's neighbours outside the common part, padded with its first entry (or a's when empty)
          pos = KP;
          if (pr.b != 255) {
            const int32_t* nb = ip + (size_t)(n0 + pr.b) * K;
            for (int k = 0; k < K; ++k)
              if (cmark[nb[k]] != tick) e[pos++] = off(nb[k]);
          }
          const uint16_t padv = pos > KP ? e[KP] : e[0];
          while (pos < 2 * KP) e[pos++] = padv;
        }
      }
    }
  }
  return worst;
}
