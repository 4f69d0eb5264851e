// Host-side graph plan for the staged EdgeConv kernel (edgeconv_tcgen05.cu).
//
// The reference builds its static kNN graph once per module in __init__ (checkerpose/model/pipeline.py:248,
// init.py:98) on farthest-point-sampled keypoints, whose order is spatially incoherent by construction: the
// 20 neighbours of 128 consecutive keypoints touch ~2000 distinct rows.  This routine, run once next to that
// knn() call, renumbers the keypoints by recursive coordinate bisection so that every tile of 128
// consecutive nodes is a compact surface patch (its neighbour lists then touch ~240 distinct rows), and
// precomputes per group of 64 nodes the list of distinct neighbour rows ("ulist", what the kernel stages in shared
// memory with bulk-async copies) and, per edge, the position of the neighbour in that list ("lidx").
// Pure integer/geometry preprocessing on the host; nothing here is on the per-RoI path.
#include <algorithm>
#include <numeric>
#include <vector>

#include "common.cuh"

namespace {

constexpr int TILE = 128;            // nodes per MMA tile: RCB boxes are whole tiles above this size
constexpr int GROUP = CP_PLAN_GROUP;  // nodes per staging group (one distinct-row list each)

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

}  // namespace

extern "C" int cp_graph_plan_build(const float* xyz, const int32_t* idx, int G, int N, int K, int umax, int32_t* perm,
                                   int32_t* idx_p, int32_t* ucount, int32_t* ulist, uint16_t* lidx) {
  CP_REQUIRE(idx && perm && idx_p && ucount && ulist && lidx, CP_E_INVALID, "cp_graph_plan_build: null pointer");
  CP_REQUIRE(G > 0 && N > 0 && K > 0 && umax > 0 && umax <= 511, CP_E_INVALID, "cp_graph_plan_build: bad sizes G=%d N=%d K=%d umax=%d", G, N, K, umax);
  const int T = (N + GROUP - 1) / GROUP;
  const int KP = (K + 7) / 8 * 8;
  int worst = 0;
  std::vector<int> ids(N), inv(N), stamp(N), local(N);
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
      int32_t* ul = ulist + ((size_t)g * T + t) * umax;
      const int n0 = t * GROUP, n1 = std::min(N, n0 + GROUP);
      // distinct neighbour rows of the tile, ascending (sequential-ish source addresses for the copies)
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
      ucount[(size_t)g * T + t] = U;
      for (int q = 0; q < U; ++q) {
        local[u[q]] = q;
        if (q < umax) ul[q] = u[q];
      }
      for (int q = U; q < umax; ++q) ul[q] = 0;
      for (int i = n0; i < n1; ++i) {
        uint16_t* li = lidx + ((size_t)g * N + i) * KP;
        // byte offset of the staged row slice; lists longer than umax are truncated and the plan is then unusable
        for (int k = 0; k < K; ++k) li[k] = (uint16_t)(std::min(local[ip[(size_t)i * K + k]], umax - 1) * 128);
        for (int k = K; k < KP; ++k) li[k] = li[0];  // padding repeats a real neighbour: harmless under max
      }
    }
  }
  return worst;
}
