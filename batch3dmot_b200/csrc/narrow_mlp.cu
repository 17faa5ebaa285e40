// Narrow MLP chains (every layer width <= 64): the edge encoders 4->16->32->64 / 4->8->16->32
// (clr_att_gnn.py:35-41, pose_gnn.py:29-35) and the edge classifiers 64->32->16->8->1 (+Sigmoid) /
// 32->16->8->4->1 (clr_att_gnn.py:49-57, pose_gnn.py:45-53). These widths do not fill a tensor-core
// tile, and as one GEMM launch per layer they stream every hidden activation through HBM. Here the
// whole nn.Sequential is ONE kernel: one thread per edge, the parameters (< 3k floats) in shared
// memory (warp-uniform LDS.128 broadcasts), every hidden activation in registers, so HBM sees only
// the chain's input and output. The backward kernel recomputes the hidden activations, walks the
// chain in reverse in registers and reduces the weight gradients of 256-row tiles through a
// transposed shared-memory stage (each thread owns up to 8 entries of dW); per-CTA partials are
// summed in fixed CTA order by a second kernel, so the result is deterministic. All arithmetic is
// fp32; bf16 is a storage format of X / Y / dY / dX only.
#include <cuda_bf16.h>

#include "b3d_common.cuh"

namespace b3d {

constexpr int NM_THREADS = 256;
constexpr int NM_RS = NM_THREADS + 4;   // row stride of the staging buffer (floats; keeps 16-byte alignment)

__host__ __device__ constexpr int cmax(int a, int b) { return a > b ? a : b; }

template <int D0_, int D1_, int D2_, int D3_, int D4_>
struct Chain {
  static constexpr int D0 = D0_, D1 = D1_, D2 = D2_, D3 = D3_, D4 = D4_;
  static constexpr int NL = D4_ > 0 ? 4 : (D3_ > 0 ? 3 : 2);
  static constexpr int DL = D4_ > 0 ? D4_ : (D3_ > 0 ? D3_ : D2_);
  // parameter layout in shared memory and in the per-CTA gradient partials
  static constexpr int W0 = 0, W1 = W0 + D1 * D0, W2 = W1 + D2 * D1, W3 = W2 + D3 * D2;
  static constexpr int B0 = W3 + D4 * D3, B1 = B0 + D1, B2 = B1 + D2, B3 = B2 + D3;
  static constexpr int NPARAM = B3 + D4;
  static constexpr int NPAD = (NPARAM + 3) & ~3;
  // second, TRANSPOSED copy of the weights in shared memory (Wt[k][j], row length padded to 4) for the
  // forward product: one LDS.128 then feeds 4 INDEPENDENT accumulators (4 outputs j..j+3 of one input k)
  static constexpr int P1 = (D1 + 3) & ~3, P2 = (D2 + 3) & ~3, P3 = (D3 + 3) & ~3, P4 = (D4 + 3) & ~3;
  static constexpr int T0 = NPAD, T1 = T0 + D0 * P1, T2 = T1 + D1 * P2, T3 = T2 + D2 * P3;
  static constexpr int NSMEM = T3 + D3 * P4;
  static constexpr int STAGE = cmax(cmax(D0 + D1, D1 + D2), cmax(D2 + D3, D3 + D4));
  static_assert(D0 % 4 == 0 && D1 % 4 == 0 && D2 % 4 == 0 && (D4 == 0 || D3 % 4 == 0), "K widths must be multiples of 4");
};

struct NMParams {
  const float* W[4];
  const float* b[4];
};
struct NMGradOut {
  float* ptr[8];     // dW0..dW3, db0..db3 (null = not wanted)
  int off[8], len[8];
  int n;
};

// ------------------------------------------------------------------ row I/O (runtime dtype)
template <int K>
__device__ __forceinline__ void nm_load_row(const void* base, int dtype, long long row, int ld, float (&v)[K]) {
  if (dtype == B3D_F64) {   // edge_attr as stored (float64): the reference's `.float()` cast folded into the load
    const double* p = reinterpret_cast<const double*>(base) + row * ld;
    if constexpr (K % 2 == 0) {
#pragma unroll
      for (int k = 0; k < K; k += 2) {
        const double2 q = __ldg(reinterpret_cast<const double2*>(p + k));
        v[k] = (float)q.x; v[k + 1] = (float)q.y;
      }
    } else {
#pragma unroll
      for (int k = 0; k < K; ++k) v[k] = (float)__ldg(p + k);
    }
    return;
  }
  if (dtype == B3D_BF16) {
    const __nv_bfloat16* p = reinterpret_cast<const __nv_bfloat16*>(base) + row * ld;
    if constexpr (K % 8 == 0) {
#pragma unroll
      for (int k = 0; k < K; k += 8) {
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(p + k));
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 f = __bfloat1622float2(h[i]);
          v[k + 2 * i] = f.x; v[k + 2 * i + 1] = f.y;
        }
      }
    } else {
#pragma unroll
      for (int k = 0; k < K; ++k) v[k] = __bfloat162float(p[k]);
    }
  } else {
    const float* p = reinterpret_cast<const float*>(base) + row * ld;
    if constexpr (K % 4 == 0) {
#pragma unroll
      for (int k = 0; k < K; k += 4) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(p + k));
        v[k] = q.x; v[k + 1] = q.y; v[k + 2] = q.z; v[k + 3] = q.w;
      }
    } else {
#pragma unroll
      for (int k = 0; k < K; ++k) v[k] = __ldg(p + k);
    }
  }
}

template <int N>
__device__ __forceinline__ void nm_store_row(void* base, int dtype, long long row, int ld, const float (&v)[N]) {
  if (dtype == B3D_BF16) {
    __nv_bfloat16* p = reinterpret_cast<__nv_bfloat16*>(base) + row * ld;
    if constexpr (N % 8 == 0) {
#pragma unroll
      for (int k = 0; k < N; k += 8) {
        uint4 q;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&q);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[k + 2 * i], v[k + 2 * i + 1]);
        *reinterpret_cast<uint4*>(p + k) = q;
      }
    } else {
#pragma unroll
      for (int k = 0; k < N; ++k) p[k] = __float2bfloat16_rn(v[k]);
    }
  } else {
    float* p = reinterpret_cast<float*>(base) + row * ld;
    if constexpr (N % 4 == 0) {
#pragma unroll
      for (int k = 0; k < N; k += 4) *reinterpret_cast<float4*>(p + k) = make_float4(v[k], v[k + 1], v[k + 2], v[k + 3]);
    } else {
#pragma unroll
      for (int k = 0; k < N; ++k) p[k] = v[k];
    }
  }
}

// ------------------------------------------------------------------ per-row layer math
// out[j] = act(b[j] + sum_k W[j][k] in[k]) from the TRANSPOSED weights Wt[k][j] (row stride NP = N padded
// to 4): k is the outer loop, so the N accumulators are independent FMA chains (with W[j][k] row-major and j
// outer, each output was one serial chain of K dependent FMAs: 25 % of the FFMA rate at 2 warps/scheduler).
// The weight reads are warp-uniform shared-memory broadcasts.
template <int K, int N, bool RELU>
__device__ __forceinline__ void nm_layer(const float* __restrict__ Wt, const float* __restrict__ bs,
                                         const float (&in)[K], float (&out)[N]) {
  constexpr int NP = (N + 3) & ~3;
  float acc[NP];
#pragma unroll
  for (int j = 0; j < NP; ++j) acc[j] = j < N ? bs[j] : 0.f;
#pragma unroll
  for (int k = 0; k < K; ++k) {
#pragma unroll
    for (int j = 0; j < NP; j += 4) {
      const float4 w = *reinterpret_cast<const float4*>(Wt + k * NP + j);
      acc[j] = fmaf(w.x, in[k], acc[j]);
      acc[j + 1] = fmaf(w.y, in[k], acc[j + 1]);
      acc[j + 2] = fmaf(w.z, in[k], acc[j + 2]);
      acc[j + 3] = fmaf(w.w, in[k], acc[j + 3]);
    }
  }
#pragma unroll
  for (int j = 0; j < N; ++j) out[j] = RELU ? fmaxf(acc[j], 0.f) : acc[j];
}

// din[k] = sum_j W[j][k] dz[j], with dz read back from this thread's own column of the staging buffer
// (rows 0..N-1 of S), so the gradient row does not stay live in registers across the weight-gradient step
template <int K, int N>
__device__ __forceinline__ void nm_layer_t(const float* __restrict__ Ws, const float* S, float (&din)[K]) {
#pragma unroll
  for (int k = 0; k < K; ++k) din[k] = 0.f;
#pragma unroll 4
  for (int j = 0; j < N; ++j) {
    const float d = S[j * NM_RS + threadIdx.x];
#pragma unroll
    for (int k = 0; k < K; k += 4) {
      const float4 w = *reinterpret_cast<const float4*>(Ws + j * K + k);
      din[k] = fmaf(w.x, d, din[k]);
      din[k + 1] = fmaf(w.y, d, din[k + 1]);
      din[k + 2] = fmaf(w.z, d, din[k + 2]);
      din[k + 3] = fmaf(w.w, d, din[k + 3]);
    }
  }
}

template <int K>
__device__ __forceinline__ void nm_relu_mask(float (&g)[K], const float (&act)[K]) {
#pragma unroll
  for (int k = 0; k < K; ++k) g[k] = act[k] > 0.f ? g[k] : 0.f;
}

template <class C>
__device__ __forceinline__ void nm_load_params(float* P, const NMParams& w) {
  constexpr int wo[4] = {C::W0, C::W1, C::W2, C::W3};
  constexpr int bo[4] = {C::B0, C::B1, C::B2, C::B3};
  constexpr int d[5] = {C::D0, C::D1, C::D2, C::D3, C::D4};
#pragma unroll
  for (int l = 0; l < C::NL; ++l) {
    const int nw = d[l] * d[l + 1];
    for (int i = threadIdx.x; i < nw; i += NM_THREADS) P[wo[l] + i] = __ldg(w.W[l] + i);
    for (int i = threadIdx.x; i < d[l + 1]; i += NM_THREADS) P[bo[l] + i] = w.b[l] ? __ldg(w.b[l] + i) : 0.f;
  }
  constexpr int to[4] = {C::T0, C::T1, C::T2, C::T3};
  constexpr int np[4] = {C::P1, C::P2, C::P3, C::P4};
#pragma unroll
  for (int l = 0; l < C::NL; ++l) {
    const int K = d[l], N = d[l + 1], NP = np[l];
    for (int i = threadIdx.x; i < K * NP; i += NM_THREADS) {
      const int k = i / NP, jj = i % NP;
      P[to[l] + i] = jj < N ? __ldg(w.W[l] + jj * K + k) : 0.f;
    }
  }
}

// ------------------------------------------------------------------ forward
template <class C>
__global__ void __launch_bounds__(NM_THREADS, 2) k_narrow_fwd(const void* __restrict__ X, int x_dtype, int ldx, long long M,
                                                          const NMParams w, int final_act, void* __restrict__ Y,
                                                          int y_dtype, int ldy) {
  __shared__ __align__(16) float P[C::NSMEM];
  nm_load_params<C>(P, w);
  __syncthreads();
  const int act = final_act & 0xFF;
  const bool in_relu = (final_act & B3D_NARROW_INPUT_RELU) != 0;
  constexpr int D3S = C::D3 > 0 ? C::D3 : 1;
  for (long long r = (long long)blockIdx.x * NM_THREADS + threadIdx.x; r < M; r += (long long)gridDim.x * NM_THREADS) {
    float x[C::D0], h1[C::D1], h2[C::D2];
    nm_load_row<C::D0>(X, x_dtype, r, ldx, x);
    if (in_relu) {
#pragma unroll
      for (int k = 0; k < C::D0; ++k) x[k] = fmaxf(x[k], 0.f);
    }
    nm_layer<C::D0, C::D1, true>(P + C::T0, P + C::B0, x, h1);
    if constexpr (C::NL == 2) {
      nm_layer<C::D1, C::D2, false>(P + C::T1, P + C::B1, h1, h2);
      if (act == B3D_ACT_RELU) {
#pragma unroll
        for (int j = 0; j < C::D2; ++j) h2[j] = fmaxf(h2[j], 0.f);
      }
      nm_store_row<C::D2>(Y, y_dtype, r, ldy, h2);
    } else if constexpr (C::NL == 3) {
      float h3[D3S];
      nm_layer<C::D1, C::D2, true>(P + C::T1, P + C::B1, h1, h2);
      nm_layer<C::D2, D3S, false>(P + C::T2, P + C::B2, h2, h3);
      if (act == B3D_ACT_SIGMOID) {
#pragma unroll
        for (int j = 0; j < D3S; ++j) h3[j] = 1.f / (1.f + expf(-h3[j]));
      }
      nm_store_row<D3S>(Y, y_dtype, r, ldy, h3);
    } else {
      float h3[D3S], h4[C::D4 > 0 ? C::D4 : 1];
      nm_layer<C::D1, C::D2, true>(P + C::T1, P + C::B1, h1, h2);
      nm_layer<C::D2, D3S, true>(P + C::T2, P + C::B2, h2, h3);
      nm_layer<D3S, (C::D4 > 0 ? C::D4 : 1), false>(P + C::T3, P + C::B3, h3, h4);
      if (act == B3D_ACT_SIGMOID) {
#pragma unroll
        for (int j = 0; j < C::D4; ++j) h4[j] = 1.f / (1.f + expf(-h4[j]));
      }
      nm_store_row<(C::D4 > 0 ? C::D4 : 1)>(Y, y_dtype, r, ldy, h4);
    }
  }
}

// ------------------------------------------------------------------ backward
__host__ __device__ constexpr int nm_ept(int K, int N) {   // dW entries per thread for a [N,K] layer over 256 threads
  int e = (N * K) / NM_THREADS;
  return e < 1 ? 1 : e;
}

// stage this thread's row of dz [N] and layer input a [K] as columns `tid` of S ([N+K][NM_RS])
template <int K, int N>
__device__ __forceinline__ void nm_stage(float* S, const float (&dz)[N], const float (&a)[K]) {
#pragma unroll
  for (int j = 0; j < N; ++j) S[j * NM_RS + threadIdx.x] = dz[j];
#pragma unroll
  for (int k = 0; k < K; ++k) S[(N + k) * NM_RS + threadIdx.x] = a[k];
}

// dW[j][kq + G i] += sum_r dz[r][j] a[r][kq + G i] over the 256 staged rows (ascending r)
template <int K, int N>
__device__ __forceinline__ void nm_accum(const float* S, float (&aw)[nm_ept(K, N)], float& ab) {
  constexpr int EPT = nm_ept(K, N);
  static_assert(K % EPT == 0, "entries per thread must divide K");
  constexpr int G = K / EPT;
  constexpr int ACTIVE = N * G;
  static_assert(ACTIVE <= NM_THREADS, "layer too large for one CTA");
  const int tid = threadIdx.x;
  if (tid < ACTIVE) {
    const int j = tid / G, kq = tid % G;
    const float* dz = S + j * NM_RS;
    const float* a = S + (N + kq) * NM_RS;
    float sb = 0.f;
#pragma unroll 2
    for (int r = 0; r < NM_THREADS; r += 4) {
      const float4 d = *reinterpret_cast<const float4*>(dz + r);
#pragma unroll
      for (int i = 0; i < EPT; ++i) {
        const float4 h = *reinterpret_cast<const float4*>(a + i * G * NM_RS + r);
        float t = aw[i];
        t = fmaf(d.x, h.x, t);
        t = fmaf(d.y, h.y, t);
        t = fmaf(d.z, h.z, t);
        t = fmaf(d.w, h.w, t);
        aw[i] = t;
      }
      sb += (d.x + d.y) + (d.z + d.w);
    }
    if (kq == 0) ab += sb;
  }
}

template <int K, int N>
__device__ __forceinline__ void nm_write_partial(float* part, int woff, int boff, const float (&aw)[nm_ept(K, N)],
                                                 float ab) {
  constexpr int EPT = nm_ept(K, N);
  constexpr int G = K / EPT;
  const int tid = threadIdx.x;
  if (tid < N * G) {
    const int j = tid / G, kq = tid % G;
#pragma unroll
    for (int i = 0; i < EPT; ++i) part[woff + j * K + kq + G * i] = aw[i];
    if (kq == 0) part[boff + j] = ab;
  }
}

template <class C>
__global__ void __launch_bounds__(NM_THREADS, 1)
k_narrow_bwd(const void* __restrict__ X, int x_dtype, int ldx, long long M, const NMParams w, int final_act,
             const void* __restrict__ dY, int dy_dtype, int lddy, void* __restrict__ dX, int dx_dtype, int lddx,
             float* __restrict__ partials) {
  extern __shared__ __align__(16) float nm_smem[];
  float* P = nm_smem;
  float* S = nm_smem + C::NSMEM;
  constexpr int DLAST = C::DL;
  constexpr int D4S = C::D4 > 0 ? C::D4 : 1;
  nm_load_params<C>(P, w);

  constexpr int D3S = C::D3 > 0 ? C::D3 : 1;
  const int act = final_act & 0xFF;
  const bool in_relu = (final_act & B3D_NARROW_INPUT_RELU) != 0;
  float aw0[nm_ept(C::D0, C::D1)], aw1[nm_ept(C::D1, C::D2)], aw2[nm_ept(C::D2, D3S)], aw3[nm_ept(D3S, D4S)];
  float ab0 = 0.f, ab1 = 0.f, ab2 = 0.f, ab3 = 0.f;
#pragma unroll
  for (int i = 0; i < nm_ept(C::D0, C::D1); ++i) aw0[i] = 0.f;
#pragma unroll
  for (int i = 0; i < nm_ept(C::D1, C::D2); ++i) aw1[i] = 0.f;
#pragma unroll
  for (int i = 0; i < nm_ept(C::D2, D3S); ++i) aw2[i] = 0.f;
#pragma unroll
  for (int i = 0; i < nm_ept(D3S, D4S); ++i) aw3[i] = 0.f;
  __syncthreads();

  const long long ntiles = (M + NM_THREADS - 1) / NM_THREADS;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long r = tile * NM_THREADS + threadIdx.x;
    const bool ok = r < M;
    float x[C::D0], h1[C::D1], h2[C::D2];
    if (ok) {
      nm_load_row<C::D0>(X, x_dtype, r, ldx, x);
      if (in_relu) {
#pragma unroll
        for (int k = 0; k < C::D0; ++k) x[k] = fmaxf(x[k], 0.f);
      }
    } else {
#pragma unroll
      for (int k = 0; k < C::D0; ++k) x[k] = 0.f;
    }
    nm_layer<C::D0, C::D1, true>(P + C::T0, P + C::B0, x, h1);
    // layer 1 is the last (linear, or ReLU when asked) layer of a 2-layer chain, a hidden ReLU layer otherwise
    if constexpr (C::NL == 2) nm_layer<C::D1, C::D2, false>(P + C::T1, P + C::B1, h1, h2);
    else nm_layer<C::D1, C::D2, true>(P + C::T1, P + C::B1, h1, h2);
    float g[DLAST];   // gradient of the chain's last pre-activation
    if (ok) {
      nm_load_row<DLAST>(dY, dy_dtype, r, lddy, g);
    } else {
#pragma unroll
      for (int j = 0; j < DLAST; ++j) g[j] = 0.f;
    }
    float g2[C::D2];  // gradient of layer 1's pre-activation (set below)
    if constexpr (C::NL == 2) {
#pragma unroll
      for (int j = 0; j < C::D2; ++j) g2[j] = (act == B3D_ACT_RELU && !(h2[j] > 0.f)) ? 0.f : g[j];
    } else if constexpr (C::NL == 3) {
      if (act == B3D_ACT_SIGMOID) {
        float z[D3S];
        nm_layer<C::D2, D3S, false>(P + C::T2, P + C::B2, h2, z);
#pragma unroll
        for (int j = 0; j < D3S; ++j) {
          const float sg = 1.f / (1.f + expf(-z[j]));
          g[j] *= sg * (1.f - sg);
        }
      }
      nm_stage<C::D2, D3S>(S, g, h2);
      __syncthreads();
      nm_accum<C::D2, D3S>(S, aw2, ab2);
      __syncthreads();
      nm_layer_t<C::D2, D3S>(P + C::W2, S, g2);
      nm_relu_mask<C::D2>(g2, h2);
    } else {
      float h3[D3S];
      nm_layer<C::D2, D3S, true>(P + C::T2, P + C::B2, h2, h3);
      if (act == B3D_ACT_SIGMOID) {
        float z[D4S];
        nm_layer<D3S, D4S, false>(P + C::T3, P + C::B3, h3, z);
#pragma unroll
        for (int j = 0; j < D4S; ++j) {
          const float sg = 1.f / (1.f + expf(-z[j]));
          g[j] *= sg * (1.f - sg);
        }
      }
      nm_stage<D3S, D4S>(S, g, h3);
      __syncthreads();
      nm_accum<D3S, D4S>(S, aw3, ab3);
      __syncthreads();
      float g3[D3S];
      nm_layer_t<D3S, D4S>(P + C::W3, S, g3);
      nm_relu_mask<D3S>(g3, h3);
      nm_stage<C::D2, D3S>(S, g3, h2);
      __syncthreads();
      nm_accum<C::D2, D3S>(S, aw2, ab2);
      __syncthreads();
      nm_layer_t<C::D2, D3S>(P + C::W2, S, g2);
      nm_relu_mask<C::D2>(g2, h2);
    }
    nm_stage<C::D1, C::D2>(S, g2, h1);
    __syncthreads();
    nm_accum<C::D1, C::D2>(S, aw1, ab1);
    __syncthreads();
    float g1[C::D1];
    nm_layer_t<C::D1, C::D2>(P + C::W1, S, g1);
    nm_relu_mask<C::D1>(g1, h1);
    nm_stage<C::D0, C::D1>(S, g1, x);
    __syncthreads();
    nm_accum<C::D0, C::D1>(S, aw0, ab0);
    __syncthreads();
    if (dX != nullptr && ok) {
      float gx[C::D0];
      nm_layer_t<C::D0, C::D1>(P + C::W0, S, gx);
      if (in_relu) nm_relu_mask<C::D0>(gx, x);      // x holds relu(input): > 0 exactly where the input was
      nm_store_row<C::D0>(dX, dx_dtype, r, lddx, gx);
    }
  }
  float* part = partials + (long long)blockIdx.x * C::NPAD;
  nm_write_partial<C::D0, C::D1>(part, C::W0, C::B0, aw0, ab0);
  nm_write_partial<C::D1, C::D2>(part, C::W1, C::B1, aw1, ab1);
  if constexpr (C::NL >= 3) nm_write_partial<C::D2, D3S>(part, C::W2, C::B2, aw2, ab2);
  if constexpr (C::NL == 4) nm_write_partial<D3S, D4S>(part, C::W3, C::B3, aw3, ab3);
}

// fixed-order sum of the per-CTA partials, scattered into the per-layer gradient tensors
__global__ void k_narrow_reduce(const float* __restrict__ partials, int nblocks, int npad, const NMGradOut o,
                                int accumulate) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npad) return;
  float* dst = nullptr;
  for (int t = 0; t < o.n; ++t)
    if (p >= o.off[t] && p < o.off[t] + o.len[t] && o.ptr[t]) dst = o.ptr[t] + (p - o.off[t]);
  if (!dst) return;
  float s = 0.f;
  for (int b = 0; b < nblocks; ++b) s += partials[(long long)b * npad + p];
  *dst = accumulate ? *dst + s : s;
}

// ------------------------------------------------------------------ host dispatch
using ChainEncMM = Chain<4, 16, 32, 64, 0>;
using ChainEncPose = Chain<4, 8, 16, 32, 0>;
using ChainClsMM = Chain<64, 32, 16, 8, 1>;
using ChainClsPose = Chain<32, 16, 8, 4, 1>;

using ChainClsTail = Chain<32, 16, 8, 1, 0>;    // 64 -> 32 of the classifier runs as a tensor-core layer in front of it
using ChainEncHead = Chain<4, 16, 32, 0, 0>;    // 32 -> 64 of the edge encoder runs as a tensor-core layer behind it

static int chain_id(int nl, const int32_t* d) {
  auto eq = [&](int n, int a, int b, int c, int e, int f) {
    return nl == n && d[0] == a && d[1] == b && d[2] == c && (n < 3 || d[3] == e) && (n < 4 || d[4] == f);
  };
  if (eq(3, 4, 16, 32, 64, 0)) return 0;
  if (eq(3, 4, 8, 16, 32, 0)) return 1;
  if (eq(4, 64, 32, 16, 8, 1)) return 2;
  if (eq(4, 32, 16, 8, 4, 1)) return 3;
  if (eq(3, 32, 16, 8, 1, 0)) return 4;
  if (eq(2, 4, 16, 32, 0, 0)) return 5;
  return -1;
}

// final_act = activation of the last layer (low byte) | B3D_NARROW_INPUT_RELU: Sigmoid on chains of >= 3 layers, ReLU
// on 2-layer chains (the chain is then the head of a longer nn.Sequential)
static bool act_ok(int nl, int final_act) {
  const int act = final_act & 0xFF;
  if (final_act & ~(0xFF | B3D_NARROW_INPUT_RELU)) return false;
  return act == B3D_ACT_NONE || (act == B3D_ACT_SIGMOID && nl >= 3) || (act == B3D_ACT_RELU && nl == 2);
}

static int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

static bool row_aligned(const void* p, int dtype, int ld, int width) {
  if (dtype == B3D_F64) return (reinterpret_cast<uintptr_t>(p) & 15) == 0 && ld % 2 == 0;   // double2 loads
  const int q = dtype == B3D_BF16 ? 8 : 4;
  if (width % q) return true;   // scalar path
  return (reinterpret_cast<uintptr_t>(p) & 15) == 0 && ld % q == 0;
}

template <class C>
static int launch_fwd(const void* X, int xd, int ldx, long long M, const NMParams& w, int act, void* Y, int yd, int ldy,
                      cudaStream_t st) {
  const long long tiles = ceil_div(M, NM_THREADS);
  const int grid = (int)(tiles < (long long)sm_count() * 8 ? tiles : (long long)sm_count() * 8);
  k_narrow_fwd<C><<<grid, NM_THREADS, 0, st>>>(X, xd, ldx, M, w, act, Y, yd, ldy);
  B3D_LAUNCH_CHECK("k_narrow_fwd");
  return 0;
}

template <class C>
static size_t bwd_smem() { return (size_t)(C::NSMEM + C::STAGE * NM_RS) * sizeof(float); }

static int bwd_grid(long long M) {
  const long long tiles = ceil_div(M > 0 ? M : 1, NM_THREADS);
  return (int)(tiles < sm_count() ? tiles : sm_count());
}

template <class C>
static int launch_bwd(const void* X, int xd, int ldx, long long M, const NMParams& w, int act, const void* dY, int gd,
                      int lddy, void* dX, int dxd, int lddx, float* const* dW, float* const* db, int flags,
                      float* partials, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(k_narrow_bwd<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bwd_smem<C>());
    if (e != cudaSuccess) return fail("k_narrow_bwd smem attribute", e);
    attr = true;
  }
  const int grid = bwd_grid(M);
  k_narrow_bwd<C><<<grid, NM_THREADS, bwd_smem<C>(), st>>>(X, xd, ldx, M, w, act, dY, gd, lddy, dX, dxd, lddx, partials);
  B3D_LAUNCH_CHECK("k_narrow_bwd");
  NMGradOut o{};
  const int wo[4] = {C::W0, C::W1, C::W2, C::W3}, bo[4] = {C::B0, C::B1, C::B2, C::B3};
  const int d[5] = {C::D0, C::D1, C::D2, C::D3, C::D4};
  o.n = 0;
  for (int l = 0; l < C::NL; ++l) {
    o.ptr[o.n] = dW ? dW[l] : nullptr; o.off[o.n] = wo[l]; o.len[o.n] = d[l] * d[l + 1]; ++o.n;
    o.ptr[o.n] = db ? db[l] : nullptr; o.off[o.n] = bo[l]; o.len[o.n] = d[l + 1]; ++o.n;
  }
  k_narrow_reduce<<<(C::NPAD + 127) / 128, 128, 0, st>>>(partials, grid, C::NPAD, o, flags & B3D_FLAG_ACCUMULATE);
  B3D_LAUNCH_CHECK("k_narrow_reduce");
  return 0;
}

static int npad_of(int id) {
  switch (id) {
    case 0: return ChainEncMM::NPAD;
    case 1: return ChainEncPose::NPAD;
    case 2: return ChainClsMM::NPAD;
    case 3: return ChainClsPose::NPAD;
    case 4: return ChainClsTail::NPAD;
    default: return ChainEncHead::NPAD;
  }
}

}  // namespace b3d

using namespace b3d;

extern "C" int b3d_narrow_mlp_supported(int32_t nl, const int32_t* dims) {
  return dims && nl >= 2 && nl <= 4 && chain_id(nl, dims) >= 0;
}

extern "C" int b3d_narrow_mlp_fwd(const void* X, int32_t x_dtype, int32_t ldx, int64_t M, int32_t nl,
                                  const int32_t* dims, const float* const* W, const float* const* b,
                                  int32_t final_act, void* Y, int32_t y_dtype, int32_t ldy, void* stream) {
  if (!X || !Y || !dims || !W || M < 0 || nl < 2 || nl > 4) return bad_arg("b3d_narrow_mlp_fwd");
  const int id = chain_id(nl, dims);
  if (id < 0) return bad_arg("b3d_narrow_mlp_fwd: unsupported layer widths");
  if (!act_ok(nl, final_act)) return bad_arg("b3d_narrow_mlp_fwd: final_act (Sigmoid: >= 3 layers, ReLU: 2 layers)");
  if (!row_aligned(X, x_dtype, ldx, dims[0]) || !row_aligned(Y, y_dtype, ldy, dims[nl]))
    return bad_arg("b3d_narrow_mlp_fwd: rows must be 16-byte aligned");
  if (M == 0) return 0;
  NMParams w{};
  for (int l = 0; l < nl; ++l) {
    if (!W[l]) return bad_arg("b3d_narrow_mlp_fwd: null weight");
    w.W[l] = W[l];
    w.b[l] = b ? b[l] : nullptr;
  }
  cudaStream_t st = (cudaStream_t)stream;
  switch (id) {
    case 0: return launch_fwd<ChainEncMM>(X, x_dtype, ldx, M, w, final_act, Y, y_dtype, ldy, st);
    case 1: return launch_fwd<ChainEncPose>(X, x_dtype, ldx, M, w, final_act, Y, y_dtype, ldy, st);
    case 2: return launch_fwd<ChainClsMM>(X, x_dtype, ldx, M, w, final_act, Y, y_dtype, ldy, st);
    case 3: return launch_fwd<ChainClsPose>(X, x_dtype, ldx, M, w, final_act, Y, y_dtype, ldy, st);
    case 4: return launch_fwd<ChainClsTail>(X, x_dtype, ldx, M, w, final_act, Y, y_dtype, ldy, st);
    default: return launch_fwd<ChainEncHead>(X, x_dtype, ldx, M, w, final_act, Y, y_dtype, ldy, st);
  }
}

extern "C" size_t b3d_narrow_mlp_bwd_workspace_bytes(int64_t M, int32_t nl, const int32_t* dims) {
  if (!dims || nl < 2 || nl > 4) return 0;
  const int id = chain_id(nl, dims);
  if (id < 0) return 0;
  return (size_t)bwd_grid(M) * npad_of(id) * sizeof(float);
}

extern "C" int b3d_narrow_mlp_bwd(const void* X, int32_t x_dtype, int32_t ldx, int64_t M, int32_t nl,
                                  const int32_t* dims, const float* const* W, const float* const* b,
                                  int32_t final_act, const void* dY, int32_t dy_dtype, int32_t lddy, void* dX,
                                  int32_t dx_dtype, int32_t lddx, float* const* dW, float* const* db,
                                  int32_t flags, void* workspace, size_t workspace_bytes, void* stream) {
  if (!X || !dY || !dims || !W || M <= 0 || nl < 2 || nl > 4 || !workspace) return bad_arg("b3d_narrow_mlp_bwd");
  const int id = chain_id(nl, dims);
  if (id < 0) return bad_arg("b3d_narrow_mlp_bwd: unsupported layer widths");
  if (!act_ok(nl, final_act)) return bad_arg("b3d_narrow_mlp_bwd: final_act (Sigmoid: >= 3 layers, ReLU: 2 layers)");
  if (workspace_bytes < b3d_narrow_mlp_bwd_workspace_bytes(M, nl, dims)) return bad_arg("b3d_narrow_mlp_bwd: workspace");
  if (!row_aligned(X, x_dtype, ldx, dims[0]) || !row_aligned(dY, dy_dtype, lddy, dims[nl]) ||
      (dX && !row_aligned(dX, dx_dtype, lddx, dims[0])))
    return bad_arg("b3d_narrow_mlp_bwd: rows must be 16-byte aligned");
  NMParams w{};
  for (int l = 0; l < nl; ++l) {
    if (!W[l]) return bad_arg("b3d_narrow_mlp_bwd: null weight");
    w.W[l] = W[l];
    w.b[l] = b ? b[l] : nullptr;
  }
  cudaStream_t st = (cudaStream_t)stream;
  float* part = reinterpret_cast<float*>(workspace);
  switch (id) {
    case 0: return launch_bwd<ChainEncMM>(X, x_dtype, ldx, M, w, final_act, dY, dy_dtype, lddy, dX, dx_dtype, lddx, dW, db, flags, part, st);
    case 1: return launch_bwd<ChainEncPose>(X, x_dtype, ldx, M, w, final_act, dY, dy_dtype, lddy, dX, dx_dtype, lddx, dW, db, flags, part, st);
    case 2: return launch_bwd<ChainClsMM>(X, x_dtype, ldx, M, w, final_act, dY, dy_dtype, lddy, dX, dx_dtype, lddx, dW, db, flags, part, st);
    case 3: return launch_bwd<ChainClsPose>(X, x_dtype, ldx, M, w, final_act, dY, dy_dtype, lddy, dX, dx_dtype, lddx, dW, db, flags, part, st);
    case 4: return launch_bwd<ChainClsTail>(X, x_dtype, ldx, M, w, final_act, dY, dy_dtype, lddy, dX, dx_dtype, lddx, dW, db, flags, part, st);
    default: return launch_bwd<ChainEncHead>(X, x_dtype, ldx, M, w, final_act, dY, dy_dtype, lddy, dX, dx_dtype, lddx, dW, db, flags, part, st);
  }
}
