// Frame-local brute-force k-NN with warp-level top-k selection, and the GATConv
// softmax-aggregate over the resulting padded neighbour table.
#include <limits.h>
#include <math.h>

#include "b3d_common.cuh"

namespace b3d {

constexpr int KNN_QB = 32;      // queries per block
constexpr int KNN_WARPS = 8;    // one warp owns KNN_QPW queries
constexpr int KNN_QPW = KNN_QB / KNN_WARPS;
constexpr int KNN_CT = 64;      // candidates staged per tile

__global__ void k_knn_blockptr(const int32_t* __restrict__ frame_ptr, int F, int32_t* __restrict__ blk_ptr) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int acc = 0;
    blk_ptr[0] = 0;
    for (int f = 0; f < F; ++f) {
      int n = frame_ptr[f + 1] - frame_ptr[f];
      acc += (n + KNN_QB - 1) / KNN_QB;
      blk_ptr[f + 1] = acc;
    }
  }
}

__device__ __forceinline__ bool pair_less(float d0, int i0, float d1, int i1) {
  return d0 < d1 || (d0 == d1 && i0 < i1);
}

// stride: floats per staged row, >= round_up(D,4), == 4 (mod 8) -> conflict-free LDS.128
// KS = sorted-list slots per lane: rank r of a query's current top-k lives in slot r / 32 of lane r % 32, so
// k <= 32 * KS (KS = 1 keeps the whole list in one register per lane; KS = 4 covers the k = 100 upstream allows).
template <int KS>
__global__ void __launch_bounds__(KNN_WARPS * 32) k_knn_frames(
    const float* __restrict__ x, int ldx, int D, int stride, const int32_t* __restrict__ frame_ptr, int F,
    const int32_t* __restrict__ blk_ptr, int k, int64_t* __restrict__ idx_out) {
  extern __shared__ __align__(16) float sm[];
  float* q_s = sm;                      // [KNN_QB][stride]
  float* c_s = sm + KNN_QB * stride;    // [KNN_CT][stride]
  const int b = blockIdx.x;
  if (b >= __ldg(blk_ptr + F)) return;
  int lo = 0, hi = F;  // find frame f: blk_ptr[f] <= b < blk_ptr[f+1]
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (__ldg(blk_ptr + mid) <= b) lo = mid; else hi = mid;
  }
  const int f = lo;
  const int fs = __ldg(frame_ptr + f), fe = __ldg(frame_ptr + f + 1);
  const int q0 = fs + (b - __ldg(blk_ptr + f)) * KNN_QB;
  const int nq = min(KNN_QB, fe - q0);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int D4 = (D + 3) & ~3;

  for (int i = tid; i < KNN_QB * D4; i += KNN_WARPS * 32) {
    int r = i / D4, d = i - r * D4;
    q_s[r * stride + d] = (r < nq && d < D) ? __ldg(x + (long long)(q0 + r) * ldx + d) : 0.f;
  }
  float bd[KNN_QPW][KS];
  int bi[KNN_QPW][KS];
#pragma unroll
  for (int qq = 0; qq < KNN_QPW; ++qq)
#pragma unroll
    for (int s = 0; s < KS; ++s) { bd[qq][s] = INFINITY; bi[qq][s] = INT_MAX; }
  const int kslot = (k - 1) >> 5, klane = (k - 1) & 31;   // where the current k-th best lives

  for (int c0 = fs; c0 < fe; c0 += KNN_CT) {
    const int nc = min(KNN_CT, fe - c0);
    __syncthreads();  // previous tile fully consumed (and q_s visible on first pass)
    for (int i = tid; i < KNN_CT * D4; i += KNN_WARPS * 32) {
      int r = i / D4, d = i - r * D4;
      c_s[r * stride + d] = (r < nc && d < D) ? __ldg(x + (long long)(c0 + r) * ldx + d) : 0.f;
    }
    __syncthreads();
#pragma unroll 1
    for (int half = 0; half < KNN_CT / 32; ++half) {
      const int j = half * 32 + lane;
      const int cand = c0 + j;
      const bool cvalid = j < nc;
      float dist[KNN_QPW];
#pragma unroll
      for (int qq = 0; qq < KNN_QPW; ++qq) dist[qq] = 0.f;
      const float4* cp = reinterpret_cast<const float4*>(c_s + j * stride);
      for (int d4 = 0; d4 < D4 / 4; ++d4) {
        const float4 cv = cp[d4];
#pragma unroll
        for (int qq = 0; qq < KNN_QPW; ++qq) {
          const float4 qv = reinterpret_cast<const float4*>(q_s + (warp * KNN_QPW + qq) * stride)[d4];
          float t;
          // ascending feature order, separate multiply and add (no FMA): the A.5 spec
          t = __fsub_rn(qv.x, cv.x); dist[qq] = __fadd_rn(dist[qq], __fmul_rn(t, t));
          t = __fsub_rn(qv.y, cv.y); dist[qq] = __fadd_rn(dist[qq], __fmul_rn(t, t));
          t = __fsub_rn(qv.z, cv.z); dist[qq] = __fadd_rn(dist[qq], __fmul_rn(t, t));
          t = __fsub_rn(qv.w, cv.w); dist[qq] = __fadd_rn(dist[qq], __fmul_rn(t, t));
        }
      }
#pragma unroll
      for (int qq = 0; qq < KNN_QPW; ++qq) {
        const int q = q0 + warp * KNN_QPW + qq;
        if (warp * KNN_QPW + qq >= nq) continue;  // warp-uniform
        float kd = INFINITY;
        int ki = INT_MAX;
#pragma unroll
        for (int s = 0; s < KS; ++s)
          if (s == kslot) { kd = __shfl_sync(0xffffffffu, bd[qq][s], klane); ki = __shfl_sync(0xffffffffu, bi[qq][s], klane); }
        const bool beats = cvalid && cand != q && pair_less(dist[qq], cand, kd, ki);
        unsigned m = __ballot_sync(0xffffffffu, beats);
        while (m) {
          const int srcl = __ffs(m) - 1;
          m &= m - 1;
          const float cd = __shfl_sync(0xffffffffu, dist[qq], srcl);
          const int ci = __shfl_sync(0xffffffffu, cand, srcl);
          int pos = 0;     // number of list entries ordered before the candidate == its insertion rank
#pragma unroll
          for (int s = 0; s < KS; ++s) pos += __popc(__ballot_sync(0xffffffffu, pair_less(bd[qq][s], bi[qq][s], cd, ci)));
          if (pos < k) {   // warp-uniform
            // shift ranks > pos up by one: rank r takes rank r-1 (lane-1 of the same slot, or lane 31 of the slot below)
            float carry_d = 0.f;
            int carry_i = 0;
#pragma unroll
            for (int s = 0; s < KS; ++s) {
              const float ud = __shfl_up_sync(0xffffffffu, bd[qq][s], 1);
              const int ui = __shfl_up_sync(0xffffffffu, bi[qq][s], 1);
              const float top_d = __shfl_sync(0xffffffffu, bd[qq][s], 31);   // leaves this slot (before the shift)
              const int top_i = __shfl_sync(0xffffffffu, bi[qq][s], 31);
              const int rank = s * 32 + lane;
              const float pd = lane == 0 ? carry_d : ud;
              const int pi = lane == 0 ? carry_i : ui;
              if (rank > pos) { bd[qq][s] = pd; bi[qq][s] = pi; }
              else if (rank == pos) { bd[qq][s] = cd; bi[qq][s] = ci; }
              carry_d = top_d; carry_i = top_i;
            }
          }
        }
      }
    }
  }
#pragma unroll
  for (int qq = 0; qq < KNN_QPW; ++qq) {
    if (warp * KNN_QPW + qq >= nq) continue;
    const long long q = q0 + warp * KNN_QPW + qq;
#pragma unroll
    for (int s = 0; s < KS; ++s) {
      const int r = s * 32 + lane;
      if (r < k) idx_out[q * k + r] = (bi[qq][s] == INT_MAX) ? (int64_t)-1 : (int64_t)bi[qq][s];
    }
  }
}

// ------------------------------------------------------------------ GAT
__global__ void __launch_bounds__(256) k_gat_scores(const float* __restrict__ h, int ldh, int D,
                                                    const float* __restrict__ att_src,
                                                    const float* __restrict__ att_dst, long long N,
                                                    float* __restrict__ a_s, float* __restrict__ a_d) {
  const int lane = threadIdx.x & 31;
  const long long n = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (n >= N) return;
  float s = 0.f, d = 0.f;
  for (int c = lane; c < D; c += 32) {
    float v = __ldg(h + n * ldh + c);
    s = fmaf(v, __ldg(att_src + c), s);
    d = fmaf(v, __ldg(att_dst + c), d);
  }
  s = warp_sum(s); d = warp_sum(d);
  if (lane == 0) { a_s[n] = s; a_d[n] = d; }
}

// Neighbour slot l of a target lives in register set l / 32 of lane l % 32 (k <= 32 * GAT_KS).
constexpr int GAT_KS = 4;

__global__ void __launch_bounds__(256) k_gat_aggregate(
    const float* __restrict__ h, int ldh, int D, const float* __restrict__ a_s, const float* __restrict__ a_d,
    const float* __restrict__ bias, const int64_t* __restrict__ nbr, int k, long long N, float slope,
    float* __restrict__ out, int ldo, float* __restrict__ alpha_out) {
  const int lane = threadIdx.x & 31;
  const long long t = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (t >= N) return;
  const int ks = (k + 31) >> 5;
  long long nb[GAT_KS];
  float z[GAT_KS], alpha[GAT_KS];
  float zmax = -INFINITY;
#pragma unroll
  for (int s = 0; s < GAT_KS; ++s) {
    const int l = s * 32 + lane;
    nb[s] = (s < ks && l < k) ? nbr[t * k + l] : -1;
    z[s] = -INFINITY;
    if (nb[s] >= 0) {
      const float v = __ldg(a_s + nb[s]) + __ldg(a_d + t);
      z[s] = v > 0.f ? v : slope * v;
    }
    zmax = fmaxf(zmax, z[s]);
  }
  zmax = warp_max(zmax);
  float den = 0.f;
#pragma unroll
  for (int s = 0; s < GAT_KS; ++s) {      // slot-major partial sums, then one warp reduction
    alpha[s] = nb[s] >= 0 ? expf(z[s] - zmax) : 0.f;
    den += alpha[s];
  }
  den = warp_sum(den);
#pragma unroll
  for (int s = 0; s < GAT_KS; ++s) {
    alpha[s] = nb[s] >= 0 ? alpha[s] / (den + 1e-16f) : 0.f;
    const int l = s * 32 + lane;
    if (alpha_out && s < ks && l < k) alpha_out[t * k + l] = alpha[s];
  }
  for (int c0 = 0; c0 < D; c0 += 32) {   // uniform trip count: the shuffles below need all lanes
    const int c = c0 + lane;
    float acc = 0.f;
#pragma unroll
    for (int s = 0; s < GAT_KS; ++s) {
      if (s >= ks) break;
      const int lim = min(32, k - s * 32);
      for (int l = 0; l < lim; ++l) {
        const float al = __shfl_sync(0xffffffffu, alpha[s], l);
        const long long nl = __shfl_sync(0xffffffffu, nb[s], l);
        if (nl >= 0 && c < D) acc = __fadd_rn(acc, __fmul_rn(al, __ldg(h + nl * ldh + c)));
      }
    }
    if (c < D) out[t * ldo + c] = acc + (bias ? __ldg(bias + c) : 0.f);
  }
}


// ------------------------------------------------------------------ GAT backward
// Target side (warp per target t, lane l = neighbour slot): d alpha_l = g[t] . h[nbr_l], softmax and
// LeakyReLU backward give ds[t,l] = d(a_s[nbr_l] + a_d[t]); dad[t] = sum_l ds[t,l].
__global__ void __launch_bounds__(256) k_gat_bwd_target(
    const float* __restrict__ g, int ldg, const float* __restrict__ h, int ldh, int D,
    const float* __restrict__ a_s, const float* __restrict__ a_d, const float* __restrict__ alpha,
    const int64_t* __restrict__ nbr, int k, long long N, float slope, float* __restrict__ ds,
    float* __restrict__ dad) {
  const int lane = threadIdx.x & 31;
  const long long t = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (t >= N) return;
  const int ks = (k + 31) >> 5;
  long long nb[GAT_KS];
  float al[GAT_KS], dal[GAT_KS], dsl[GAT_KS];
#pragma unroll
  for (int s = 0; s < GAT_KS; ++s) {
    const int l = s * 32 + lane;
    nb[s] = (s < ks && l < k) ? nbr[t * k + l] : -1;
    al[s] = nb[s] >= 0 ? alpha[t * k + l] : 0.f;
    dal[s] = 0.f;                        // slot (s, lane) ends up with g[t] . h[nbr]
  }
#pragma unroll
  for (int s = 0; s < GAT_KS; ++s) {
    if (s >= ks) break;
    const int lim = min(32, k - s * 32);
    for (int l = 0; l < lim; ++l) {
      const long long nl = __shfl_sync(0xffffffffu, nb[s], l);
      float part = 0.f;
      if (nl >= 0)
        for (int c = lane; c < D; c += 32) part = fmaf(__ldg(g + t * ldg + c), __ldg(h + nl * ldh + c), part);
      part = warp_sum(part);
      if (lane == l) dal[s] = part;
    }
  }
  float dot = 0.f;
#pragma unroll
  for (int s = 0; s < GAT_KS; ++s) dot += al[s] * dal[s];
  dot = warp_sum(dot);                   // sum_m alpha_m d alpha_m
  float tot = 0.f;
#pragma unroll
  for (int s = 0; s < GAT_KS; ++s) {
    dsl[s] = 0.f;
    if (nb[s] >= 0) {
      const float pre = __ldg(a_s + nb[s]) + __ldg(a_d + t);
      dsl[s] = al[s] * (dal[s] - dot) * (pre > 0.f ? 1.f : slope);
    }
    const int l = s * 32 + lane;
    if (s < ks && l < k) ds[t * k + l] = dsl[s];
    tot += dsl[s];
  }
  tot = warp_sum(tot);
  if (lane == 0) dad[t] = tot;
}

// Source side (warp per node n, lanes over feature columns): the entries (t, l) that list n as a
// neighbour come from the CSR of the flattened table grouped by source (b3d_csr_build; -1 padding is
// mapped to the dummy node N), visited in ascending position so the sums are deterministic:
//   dh[n] = sum alpha[t,l] g[t] + (sum ds[t,l]) att_src + dad[n] att_dst ;  das[n] = sum ds[t,l].
__global__ void __launch_bounds__(256) k_gat_bwd_source(
    const float* __restrict__ g, int ldg, int D, const float* __restrict__ alpha, const float* __restrict__ ds,
    const float* __restrict__ dad, const float* __restrict__ att_src, const float* __restrict__ att_dst,
    const int32_t* __restrict__ rowptr, const int32_t* __restrict__ perm, int k, long long N,
    float* __restrict__ dh, int lddh, float* __restrict__ das) {
  const int lane = threadIdx.x & 31;
  const long long n = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (n >= N) return;
  const int beg = __ldg(rowptr + n), end = __ldg(rowptr + n + 1);
  float s = 0.f;
  for (int q = beg; q < end; ++q) s += __ldg(ds + __ldg(perm + q));
  if (lane == 0) das[n] = s;
  const float d = __ldg(dad + n);
  for (int c = lane; c < D; c += 32) {
    float acc = 0.f;
    for (int q = beg; q < end; ++q) {
      const int p = __ldg(perm + q);
      acc = fmaf(__ldg(alpha + p), __ldg(g + (long long)(p / k) * ldg + c), acc);
    }
    dh[n * lddh + c] = acc + s * __ldg(att_src + c) + d * __ldg(att_dst + c);
  }
}

}  // namespace b3d

using namespace b3d;

extern "C" int b3d_knn_frames(const float* x, int32_t ldx, int32_t D, const int32_t* frame_ptr, int32_t F,
                              int64_t N, int32_t k, int64_t* idx_out, int32_t* scratch, void* stream) {
  if (!x || !frame_ptr || !idx_out || !scratch) return bad_arg("b3d_knn_frames: null pointer");
  if (k < 1 || k > 128) return bad_arg("b3d_knn_frames: k must be in [1,128]");
  if (D < 1 || D > 256) return bad_arg("b3d_knn_frames: D must be in [1,256]");
  if (N == 0 || F == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  int stride = (D + 3) & ~3;
  while ((stride & 7) != 4) stride += 4;
  size_t smem = sizeof(float) * (size_t)(KNN_QB + KNN_CT) * stride;
  static size_t smem_set = 0;
  if (smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(k_knn_frames<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_knn_frames<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_knn_frames<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail("knn smem attr", e);
    smem_set = smem;
  }
  k_knn_blockptr<<<1, 32, 0, st>>>(frame_ptr, F, scratch);
  B3D_LAUNCH_CHECK("k_knn_blockptr");
  unsigned grid = (unsigned)(ceil_div(N, KNN_QB) + F);
  if (k <= 32) k_knn_frames<1><<<grid, KNN_WARPS * 32, smem, st>>>(x, ldx, D, stride, frame_ptr, F, scratch, k, idx_out);
  else if (k <= 64) k_knn_frames<2><<<grid, KNN_WARPS * 32, smem, st>>>(x, ldx, D, stride, frame_ptr, F, scratch, k, idx_out);
  else k_knn_frames<4><<<grid, KNN_WARPS * 32, smem, st>>>(x, ldx, D, stride, frame_ptr, F, scratch, k, idx_out);
  B3D_LAUNCH_CHECK("k_knn_frames");
  return 0;
}

extern "C" int b3d_gat_aggregate(const float* h, int32_t ldh, int32_t D, const float* att_src,
                                 const float* att_dst, const float* bias, const int64_t* nbr, int32_t k,
                                 int64_t N, float slope, float* out, int32_t ldo, float* alpha_out,
                                 float* scratch, void* stream) {
  if (!h || !att_src || !att_dst || !nbr || !out || !scratch) return bad_arg("b3d_gat_aggregate: null pointer");
  if (k < 1 || k > 32 * GAT_KS) return bad_arg("b3d_gat_aggregate: k must be in [1,128]");
  if (N == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  float* a_s = scratch;
  float* a_d = scratch + N;
  k_gat_scores<<<(unsigned)ceil_div(N, 8), 256, 0, st>>>(h, ldh, D, att_src, att_dst, N, a_s, a_d);
  B3D_LAUNCH_CHECK("k_gat_scores");
  k_gat_aggregate<<<(unsigned)ceil_div(N, 8), 256, 0, st>>>(h, ldh, D, a_s, a_d, bias, nbr, k, N, slope, out, ldo, alpha_out);
  B3D_LAUNCH_CHECK("k_gat_aggregate");
  return 0;
}

extern "C" int b3d_gat_bwd(const float* dout, int32_t ldg, const float* h, int32_t ldh, int32_t D,
                           const float* att_src, const float* att_dst, const int64_t* nbr, int32_t k, int64_t N,
                           float slope, const float* alpha, const float* a_s_a_d, const int32_t* rowptr_src,
                           const int32_t* perm_src, float* dh, int32_t lddh, float* ds, float* das_dad,
                           void* stream) {
  if (!dout || !h || !att_src || !att_dst || !nbr || !alpha || !a_s_a_d || !rowptr_src || !perm_src || !dh || !ds ||
      !das_dad)
    return bad_arg("b3d_gat_bwd: null pointer");
  if (k < 1 || k > 32 * GAT_KS) return bad_arg("b3d_gat_bwd: k must be in [1,128]");
  if (N == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  float* das = das_dad;
  float* dad = das_dad + N;
  k_gat_bwd_target<<<(unsigned)ceil_div(N, 8), 256, 0, st>>>(dout, ldg, h, ldh, D, a_s_a_d, a_s_a_d + N, alpha, nbr, k, N,
                                                            slope, ds, dad);
  B3D_LAUNCH_CHECK("k_gat_bwd_target");
  k_gat_bwd_source<<<(unsigned)ceil_div(N, 8), 256, 0, st>>>(dout, ldg, D, alpha, ds, dad, att_src, att_dst, rowptr_src,
                                                            perm_src, k, N, dh, lddh, das);
  B3D_LAUNCH_CHECK("k_gat_bwd_source");
  return 0;
}
