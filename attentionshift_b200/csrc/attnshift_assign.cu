// Point-token <-> GT matching of seed_pseudo_gt on the device.
//
// The reference (mmdet/core/bbox/assigners/hungarian_point_assigner.py:95-99) copies the [proposals x GTs] cost matrix to
// the host and calls scipy.optimize.linear_sum_assignment: a D2H sync in the middle of seed_pseudo_gt.  scipy's solver is
// the shortest-augmenting-path algorithm for the rectangular assignment problem (D. F. Crouse, "On implementing 2D
// rectangular assignment algorithms", IEEE T-AES 52(4), 2016): rows are added one at a time, each by a Dijkstra search
// over the columns with dual variables u, v.  This kernel runs that algorithm in fp64, one warp per image: the lanes share
// the column scan of every search step (costs, tentative distances, arg-min), the short bookkeeping in between is
// uniform.  Ties are resolved exactly as the sequential scan of the published algorithm resolves them (see better()), so
// that integer-valued cost matrices give the same matching, not just the same total cost.
#include "common.cuh"

#include <math.h>

using namespace asb;

namespace {

constexpr int HUNG_MAX = 512;     // max(#proposals, #GTs) per image

struct Cand {                     // one column of the scan: tentative distance, "column is free", position in `remaining`
  double val;
  int free_col;
  int it;
};

// The sequential scan keeps the FIRST column of the smallest distance, except that a later FREE column of the same distance
// replaces it (it ends the search at once).  As a total order: smaller distance; then free before taken; then among free
// columns the last position, among taken ones the first.
__device__ __forceinline__ bool better(const Cand& a, const Cand& b) {
  if (a.it < 0) return false;
  if (b.it < 0) return true;
  if (a.val != b.val) return a.val < b.val;
  if (a.free_col != b.free_col) return a.free_col > b.free_col;
  return a.free_col ? a.it > b.it : a.it < b.it;
}

__global__ void __launch_bounds__(32)
hungarian_points_kernel(const float* __restrict__ cost, const int* __restrict__ g_first, const int* __restrict__ g_count,
                        int P, int* __restrict__ pos_inds, int* __restrict__ pos_gt, int* __restrict__ status) {
  __shared__ double u[HUNG_MAX], v[HUNG_MAX], spc[HUNG_MAX];
  __shared__ int path[HUNG_MAX], row4col[HUNG_MAX], col4row[HUNG_MAX], remaining[HUNG_MAX];
  __shared__ unsigned char SR[HUNG_MAX], SC[HUNG_MAX];
  const int img = blockIdx.x, lane = threadIdx.x;
  const int G = g_count[img], g0 = g_first[img];
  if (G <= 0) {
    if (status && lane == 0) status[img] = 0;
    return;
  }
  const float* c = cost + (size_t)P * g0;           // [G, P] row-major: one row of proposal costs per GT
  // the solver wants rows <= columns: with fewer GTs than proposals (the usual case) the GTs are the rows
  const bool tr = G < P;
  const int nr = tr ? G : P, nc = tr ? P : G;
  auto cst = [&](int i, int j) -> double { return (double)(tr ? c[(size_t)i * P + j] : c[(size_t)j * P + i]); };

  bool bad = false;                                 // scipy rejects NaN / -inf entries
  for (int e = lane; e < P * G; e += 32) {
    const float x = c[e];
    bad |= (x != x) || (x == -INFINITY);
  }
  bad = __any_sync(0xffffffffu, bad);
  for (int i = lane; i < nr; i += 32) { u[i] = 0.0; col4row[i] = -1; }
  for (int j = lane; j < nc; j += 32) { v[j] = 0.0; row4col[j] = -1; path[j] = -1; }
  __syncwarp();

  for (int cur = 0; cur < nr && !bad; ++cur) {
    for (int i = lane; i < nr; i += 32) SR[i] = 0;
    for (int j = lane; j < nc; j += 32) { SC[j] = 0; spc[j] = INFINITY; remaining[j] = nc - j - 1; }
    __syncwarp();
    int num_remaining = nc, i = cur, sink = -1;
    double min_val = 0.0;
    while (sink < 0) {
      if (lane == 0) SR[i] = 1;
      const double ui = u[i];
      Cand best{INFINITY, 0, -1};
      for (int it = lane; it < num_remaining; it += 32) {
        const int j = remaining[it];
        const double r = min_val + cst(i, j) - ui - v[j];
        double s = spc[j];
        if (r < s) { path[j] = i; spc[j] = r; s = r; }
        const Cand cand{s, row4col[j] < 0 ? 1 : 0, it};
        if (better(cand, best)) best = cand;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        Cand other;
        other.val = __shfl_xor_sync(0xffffffffu, best.val, o);
        other.free_col = __shfl_xor_sync(0xffffffffu, best.free_col, o);
        other.it = __shfl_xor_sync(0xffffffffu, best.it, o);
        if (better(other, best)) best = other;
      }
      min_val = best.val;
      if (best.it < 0 || min_val == INFINITY) { bad = true; break; }       // infeasible
      __syncwarp();
      const int j = remaining[best.it];
      if (row4col[j] < 0) sink = j; else i = row4col[j];
      __syncwarp();
      if (lane == 0) { SC[j] = 1; remaining[best.it] = remaining[num_remaining - 1]; }
      --num_remaining;
      __syncwarp();
    }
    if (bad) break;
    // dual variables
    for (int r = lane; r < nr; r += 32)
      if (r == cur) u[r] += min_val;
      else if (SR[r]) u[r] += min_val - spc[col4row[r]];
    for (int j = lane; j < nc; j += 32)
      if (SC[j]) v[j] -= min_val - spc[j];
    __syncwarp();
    // augment along the path back to the new row
    if (lane == 0) {
      int j = sink;
      while (true) {
        const int r = path[j];
        row4col[j] = r;
        const int t = col4row[r]; col4row[r] = j; j = t;
        if (r == cur) break;
      }
    }
    __syncwarp();
  }

  if (status) { if (lane == 0) status[img] = bad ? 1 : 0; }
  const int n_match = nr;                           // = min(P, G)
  if (bad) {                                        // keep every index in range: GT k <-> proposal k
    for (int k = lane; k < n_match; k += 32) { pos_inds[g0 + k] = k; pos_gt[g0 + k] = k; }
    return;
  }
  // PointPseudoSampler (point_pseudo_sampler.py:34-37): matched proposals in ascending order with the GT of each
  if (lane == 0) {
    int k = 0;
    for (int p = 0; p < P; ++p) {
      const int g = tr ? row4col[p] : col4row[p];
      if (g >= 0) { pos_inds[g0 + k] = p; pos_gt[g0 + k] = g; ++k; }
    }
  }
}

}  // namespace

extern "C" int as_hungarian_points(const float* cost, const int* g_first, const int* g_count, int n_img, int P, int max_g,
                                   int* pos_inds, int* pos_gt, int* status, cudaStream_t stream) {
  if (n_img <= 0) return 0;
  if (P < 1 || P > HUNG_MAX || max_g > HUNG_MAX) return AS_ERR_BAD_ARG;
  hungarian_points_kernel<<<n_img, 32, 0, stream>>>(cost, g_first, g_count, P, pos_inds, pos_gt, status);
  AS_LAUNCH_CHECK();
  return 0;
}
