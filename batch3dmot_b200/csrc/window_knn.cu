// Graph-construction k-NN (the step that builds the hot path's edge_index): for every current node of a window, the
// k nearest same-category nodes of the earlier frames under the reference's normalised motion metric
//   m = ( 1/2 t/max(t) + 1/4 y/max(y) + 1/4 v/max(v) ) / max(.)          (batch_3dmot/utils/graph_utils.py:33-88)
// with t = xy centre distance, y = |yaw difference| (geo_utils.py:8-21,46-57), v = |velocity difference| (3-vector),
// all float64, maxima over the node's candidate list (construct_detection_graph_disjoint_parallel_only_poses.py:
// 204-224). One warp per current node: three passes over its candidates (maxima, normaliser, selection) with
// the same float64 operation order as the reference — explicit __dmul_rn / __dadd_rn so that no multiply-add is
// fused — and a warp-level sorted insertion list (rank r in slot r / 32 of lane r % 32) for the k smallest
// (metric, candidate position) pairs. Ties are therefore broken by candidate position (ascending node id within
// the category), deterministically; torch.topk leaves their order unspecified upstream.
#include <math.h>

#include "b3d_common.cuh"

namespace b3d {

constexpr int WK_SLOTS = 2;          // k + 1 <= 64
constexpr int WK_WARPS = 4;

__device__ __forceinline__ double wk_angle_diff(double x, double y) {
  // (x - y + period/2) % period - period/2, Python/torch remainder semantics (result has the sign of the period)
  const double period = 6.283185307179586, half = 3.141592653589793;
  double r = fmod(__dadd_rn(__dadd_rn(x, -y), half), period);
  if (r < 0.0) r = __dadd_rn(r, period);
  return __dadd_rn(r, -half);
}

struct WkNode { double cx, cy, vx, vy, vz, yaw; };

__device__ __forceinline__ WkNode wk_load(const double* c, const double* v, const double* yaw, long long i) {
  WkNode n;
  n.cx = c[3 * i]; n.cy = c[3 * i + 1];
  n.vx = v[3 * i]; n.vy = v[3 * i + 1]; n.vz = v[3 * i + 2];
  n.yaw = yaw[i];
  return n;
}

__device__ __forceinline__ void wk_terms(const WkNode& a, const WkNode& b, double& t, double& y, double& v) {
  const double dx = __dadd_rn(b.cx, -a.cx), dy = __dadd_rn(b.cy, -a.cy);
  t = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
  const double ux = __dadd_rn(b.vx, -a.vx), uy = __dadd_rn(b.vy, -a.vy), uz = __dadd_rn(b.vz, -a.vz);
  v = fabs(sqrt(__dadd_rn(__dadd_rn(__dmul_rn(ux, ux), __dmul_rn(uy, uy)), __dmul_rn(uz, uz))));
  y = fabs(wk_angle_diff(a.yaw, b.yaw));
}

__device__ __forceinline__ double wk_warp_max(double x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x = fmax(x, __shfl_xor_sync(0xffffffffu, x, o));
  return x;
}

__device__ __forceinline__ bool wk_less(double d0, int i0, double d1, int i1) {
  return d0 < d1 || (d0 == d1 && i0 < i1);
}

__global__ void __launch_bounds__(WK_WARPS * 32) k_window_knn(
    const double* __restrict__ center, const double* __restrict__ velocity, const double* __restrict__ yaw,
    const int64_t* __restrict__ order, const int64_t* __restrict__ cur, const int64_t* __restrict__ first,
    const int64_t* __restrict__ cnt, long long R, int top_knn, int kmax, int64_t* __restrict__ ex,
    int32_t* __restrict__ flags) {
  const int lane = threadIdx.x & 31;
  const long long r = (long long)blockIdx.x * WK_WARPS + (threadIdx.x >> 5);
  if (r >= R) return;
  const long long c = cur[r], f0 = first[r];
  const int n = (int)cnt[r];
  const int k = n < top_knn ? n : top_knn;
  const int kk = k + 1 < n ? k + 1 : n;          // one extra rank: detects a tie at the cut
  const WkNode me = wk_load(center, velocity, yaw, c);
  // pass 1: maxima of the three terms
  double tmax = -INFINITY, ymax = -INFINITY, vmax = -INFINITY;
  for (int j = lane; j < n; j += 32) {
    double t, y, v;
    wk_terms(me, wk_load(center, velocity, yaw, order[f0 + j]), t, y, v);
    tmax = fmax(tmax, t); ymax = fmax(ymax, y); vmax = fmax(vmax, v);
  }
  tmax = wk_warp_max(tmax); ymax = wk_warp_max(ymax); vmax = wk_warp_max(vmax);
  // pass 2: maximum of the weighted sum
  double mmax = -INFINITY;
  bool nan = false;
  for (int j = lane; j < n; j += 32) {
    double t, y, v;
    wk_terms(me, wk_load(center, velocity, yaw, order[f0 + j]), t, y, v);
    const double m = __dadd_rn(__dadd_rn(__dmul_rn(0.5, t / tmax), __dmul_rn(0.25, y / ymax)), __dmul_rn(0.25, v / vmax));
    nan |= (m != m);
    mmax = fmax(mmax, m);
  }
  mmax = wk_warp_max(mmax);
  // pass 3: the kk smallest (m / mmax, position)
  double bd[WK_SLOTS];
  int bi[WK_SLOTS];
#pragma unroll
  for (int s = 0; s < WK_SLOTS; ++s) { bd[s] = INFINITY; bi[s] = 0x7fffffff; }
  const int kslot = (kk - 1) >> 5, klane = (kk - 1) & 31;
  for (int j0 = 0; j0 < n; j0 += 32) {
    const int j = j0 + lane;
    double q = INFINITY;
    if (j < n) {
      double t, y, v;
      wk_terms(me, wk_load(center, velocity, yaw, order[f0 + j]), t, y, v);
      const double m = __dadd_rn(__dadd_rn(__dmul_rn(0.5, t / tmax), __dmul_rn(0.25, y / ymax)), __dmul_rn(0.25, v / vmax));
      q = m / mmax;
      nan |= (q != q);
    }
    double kd = INFINITY;
    int ki = 0x7fffffff;
#pragma unroll
    for (int s = 0; s < WK_SLOTS; ++s)
      if (s == kslot) { kd = __shfl_sync(0xffffffffu, bd[s], klane); ki = __shfl_sync(0xffffffffu, bi[s], klane); }
    unsigned mset = __ballot_sync(0xffffffffu, j < n && wk_less(q, j, kd, ki));
    while (mset) {
      const int srcl = __ffs(mset) - 1;
      mset &= mset - 1;
      const double cd = __shfl_sync(0xffffffffu, q, srcl);
      const int ci = j0 + srcl;
      int pos = 0;
#pragma unroll
      for (int s = 0; s < WK_SLOTS; ++s) pos += __popc(__ballot_sync(0xffffffffu, wk_less(bd[s], bi[s], cd, ci)));
      if (pos < kk) {
        double carry_d = 0.0;
        int carry_i = 0;
#pragma unroll
        for (int s = 0; s < WK_SLOTS; ++s) {
          const double ud = __shfl_up_sync(0xffffffffu, bd[s], 1);
          const int ui = __shfl_up_sync(0xffffffffu, bi[s], 1);
          const double top_d = __shfl_sync(0xffffffffu, bd[s], 31);
          const int top_i = __shfl_sync(0xffffffffu, bi[s], 31);
          const int rank = s * 32 + lane;
          const double pd = lane == 0 ? carry_d : ud;
          const int pi = lane == 0 ? carry_i : ui;
          if (rank > pos) { bd[s] = pd; bi[s] = pi; }
          else if (rank == pos) { bd[s] = cd; bi[s] = ci; }
          carry_d = top_d; carry_i = top_i;
        }
      }
    }
  }
  // ties among the first kk sorted values (the reference's torch.topk order is unspecified there)
  bool tie = false;
#pragma unroll
  for (int s = 0; s < WK_SLOTS; ++s) {
    const int rank = s * 32 + lane;
    double prev = __shfl_up_sync(0xffffffffu, bd[s], 1);
    const double below = s > 0 ? __shfl_sync(0xffffffffu, bd[s > 0 ? s - 1 : 0], 31) : 0.0;
    if (lane == 0) prev = below;
    if (rank > 0 && rank < kk && bd[s] == prev) tie = true;
    // (a row whose metric is NaN inserted nothing: its list still holds the sentinels; the caller resolves it)
    if (rank < kmax) ex[r * kmax + rank] = (rank < k && bi[s] < n) ? order[f0 + bi[s]] : (int64_t)-1;
  }
  const unsigned any_nan = __ballot_sync(0xffffffffu, nan), any_tie = __ballot_sync(0xffffffffu, tie);
  if (lane == 0) flags[r] = (any_nan ? 1 : 0) | (any_tie ? 2 : 0);
}

}  // namespace b3d

using namespace b3d;

extern "C" int b3d_window_knn(const double* center, const double* velocity, const double* yaw, const int64_t* order,
                              const int64_t* cur, const int64_t* first, const int64_t* cnt, int64_t R, int32_t top_knn,
                              int32_t kmax, int64_t* ex, int32_t* flags, void* stream) {
  if (R == 0) return 0;
  if (!center || !velocity || !yaw || !order || !cur || !first || !cnt || !ex || !flags)
    return bad_arg("b3d_window_knn: null pointer");
  if (top_knn < 1 || top_knn > 63 || kmax < 1 || kmax > top_knn) return bad_arg("b3d_window_knn: 1 <= kmax <= top_knn <= 63");
  k_window_knn<<<(unsigned)ceil_div(R, WK_WARPS), WK_WARPS * 32, 0, (cudaStream_t)stream>>>(
      center, velocity, yaw, order, cur, first, cnt, R, top_knn, kmax, ex, flags);
  B3D_LAUNCH_CHECK("k_window_knn");
  return 0;
}
