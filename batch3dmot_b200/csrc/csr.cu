// One-time CSR/CSC construction: stable LSD radix sort of edge endpoints (8-bit digits),
// degree histogram + exclusive scan for rowptr. Integer work only; bit-exact against
// torch.argsort(stable=True) / bincount+cumsum (oracle/ref_restated.py: csr_build).
#include "b3d_common.cuh"

namespace b3d {

thread_local char g_err[256] = "ok";
long long g_launches = 0;

constexpr int RS_THREADS = 256;
constexpr int RS_ROUNDS = 8;
constexpr int RS_CHUNK = RS_THREADS * RS_ROUNDS;  // keys per block

__global__ void k_convert(const int64_t* __restrict__ ei, int64_t E, int64_t N,
                          int32_t* __restrict__ src32, int32_t* __restrict__ dst32,
                          int32_t* __restrict__ deg_src, int32_t* __restrict__ deg_dst,
                          int32_t* __restrict__ status) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  int64_t s = ei[e], d = ei[E + e];
  if (s < 0 || s >= N || d < 0 || d >= N) {
    atomicOr(status, 1);
    s = 0; d = 0;
  }
  src32[e] = (int32_t)s;
  dst32[e] = (int32_t)d;
  atomicAdd(&deg_src[s], 1);  // integer atomics: order-independent result
  atomicAdd(&deg_dst[d], 1);
}

// Single-block exclusive scan of `in[0..n)` into `out[0..n]` (out[n] = total): tiles of 4096 elements,
// 4 consecutive elements per thread (coalesced), shuffle scan inside warps, one smem hop across warps,
// running carry between tiles. (The previous chunk-per-thread version was uncoalesced: 700 us at n = 245k.)
__global__ void __launch_bounds__(1024) k_exclusive_scan(const int32_t* __restrict__ in, int32_t* __restrict__ out,
                                                        int64_t n) {
  __shared__ int32_t s_warp[32];
  __shared__ int32_t s_carry;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  if (t == 0) s_carry = 0;
  __syncthreads();
  for (int64_t base = 0; base < n; base += 4096) {
    const int64_t i0 = base + 4 * (int64_t)t;
    int32_t v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = (i0 + j < n) ? in[i0 + j] : 0;
    const int32_t mine = v[0] + v[1] + v[2] + v[3];
    int32_t inc = mine;                                  // inclusive scan of thread sums within the warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int32_t u = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += u;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      int32_t w = s_warp[lane], winc = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int32_t u = __shfl_up_sync(0xffffffffu, winc, o);
        if (lane >= o) winc += u;
      }
      s_warp[lane] = winc - w;                           // exclusive prefix of warp totals
    }
    __syncthreads();
    int32_t run = s_carry + s_warp[warp] + (inc - mine);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (i0 + j < n) out[i0 + j] = run;
      run += v[j];
    }
    __syncthreads();
    if (t == 1023) s_carry = run;                        // last thread's running total = carry of the next tile
    __syncthreads();
  }
  if (t == 0) out[n] = s_carry;
}

__global__ void k_radix_hist(const int32_t* __restrict__ keys, int64_t E, int shift,
                             int32_t* __restrict__ counts, int nblocks) {
  __shared__ int32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  int64_t base = (int64_t)blockIdx.x * RS_CHUNK;
#pragma unroll
  for (int r = 0; r < RS_ROUNDS; ++r) {
    int64_t i = base + r * RS_THREADS + threadIdx.x;
    if (i < E) atomicAdd(&h[(keys[i] >> shift) & 255], 1);
  }
  __syncthreads();
  counts[(int64_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
}

// Stable scatter: offsets[d*nblocks+b] = global start of digit d for block b.
// vals_in == nullptr means the identity permutation (first pass).
__global__ void k_radix_scatter(const int32_t* __restrict__ keys_in, const int32_t* __restrict__ vals_in,
                                int32_t* __restrict__ keys_out, int32_t* __restrict__ vals_out,
                                int64_t E, int shift, const int32_t* __restrict__ offsets, int nblocks) {
  __shared__ int32_t whist[RS_THREADS / 32][256];
  __shared__ int32_t run[256];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  run[t] = offsets[(int64_t)t * nblocks + blockIdx.x];
  for (int w = 0; w < RS_THREADS / 32; ++w) whist[w][t] = 0;
  __syncthreads();
  int64_t base = (int64_t)blockIdx.x * RS_CHUNK;
  for (int r = 0; r < RS_ROUNDS; ++r) {
    int64_t i = base + r * RS_THREADS + t;
    bool valid = i < E;
    int32_t key = valid ? keys_in[i] : 0;
    int32_t val = valid ? (vals_in ? vals_in[i] : (int32_t)i) : 0;
    int d = valid ? ((key >> shift) & 255) : (256 + lane);  // invalid lanes match only themselves
    unsigned peers = __match_any_sync(0xffffffffu, d);
    int rank = __popc(peers & ((1u << lane) - 1u));
    if (valid && rank == 0) whist[warp][d] = __popc(peers);
    __syncthreads();
    {  // thread t owns digit t: turn per-warp counts into per-warp start offsets
      int32_t acc = run[t];
#pragma unroll
      for (int w = 0; w < RS_THREADS / 32; ++w) {
        int32_t c = whist[w][t];
        whist[w][t] = acc;
        acc += c;
      }
      run[t] = acc;
    }
    __syncthreads();
    if (valid) {
      int32_t pos = whist[warp][d] + rank;
      keys_out[pos] = key;
      vals_out[pos] = val;
    }
    __syncthreads();
    for (int w = 0; w < RS_THREADS / 32; ++w) whist[w][t] = 0;
    __syncthreads();
  }
}

struct CsrWs {
  int32_t *deg_src, *deg_dst, *keys_a, *keys_b, *vals_a, *counts, *offsets;
  size_t bytes;
};

static CsrWs carve(void* ws, int64_t E, int64_t N) {
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  int64_t nblocks = ceil_div(E > 0 ? E : 1, RS_CHUNK);
  char* p = (char*)ws;
  CsrWs w;
  size_t o = 0;
  w.deg_src = (int32_t*)(p + o); o += al(sizeof(int32_t) * (N + 1));
  w.deg_dst = (int32_t*)(p + o); o += al(sizeof(int32_t) * (N + 1));
  w.keys_a = (int32_t*)(p + o); o += al(sizeof(int32_t) * E);
  w.keys_b = (int32_t*)(p + o); o += al(sizeof(int32_t) * E);
  w.vals_a = (int32_t*)(p + o); o += al(sizeof(int32_t) * E);
  w.counts = (int32_t*)(p + o); o += al(sizeof(int32_t) * 256 * nblocks);
  w.offsets = (int32_t*)(p + o); o += al(sizeof(int32_t) * (256 * nblocks + 1));
  w.bytes = o;
  return w;
}

static int radix_sort_perm(const int32_t* keys, int64_t E, int64_t N, int32_t* perm_out, CsrWs& w,
                           cudaStream_t st) {
  int bits = 1;
  while (((int64_t)1 << bits) < N) ++bits;
  int npass = (bits + 7) / 8;
  int nblocks = (int)ceil_div(E, RS_CHUNK);
  // ping-pong so that the LAST pass writes values into perm_out
  const int32_t* kin = keys;
  const int32_t* vin = nullptr;
  for (int p = 0; p < npass; ++p) {
    bool last = (p == npass - 1);
    int32_t* kout = (p & 1) ? w.keys_b : w.keys_a;
    int32_t* vout = last ? perm_out : (((npass - 1 - p) & 1) ? w.vals_a : perm_out);
    k_radix_hist<<<nblocks, RS_THREADS, 0, st>>>(kin, E, 8 * p, w.counts, nblocks);
    B3D_LAUNCH_CHECK("radix_hist");
    k_exclusive_scan<<<1, 1024, 0, st>>>(w.counts, w.offsets, (int64_t)256 * nblocks);
    B3D_LAUNCH_CHECK("radix_scan");
    k_radix_scatter<<<nblocks, RS_THREADS, 0, st>>>(kin, vin, kout, vout, E, 8 * p, w.offsets, nblocks);
    B3D_LAUNCH_CHECK("radix_scatter");
    kin = kout;
    vin = vout;
  }
  return 0;
}

}  // namespace b3d

using namespace b3d;

extern "C" const char* b3d_last_error(void) { return g_err; }
extern "C" int64_t b3d_launch_count(void) { return g_launches; }
extern "C" void b3d_reset_launch_count(void) { g_launches = 0; }

extern "C" size_t b3d_csr_workspace_bytes(int64_t E, int64_t N) { return carve(nullptr, E, N).bytes; }

extern "C" int b3d_csr_build(const int64_t* edge_index, int64_t E, int64_t N, int32_t* src32,
                             int32_t* dst32, int32_t* rowptr_dst, int32_t* perm_dst,
                             int32_t* rowptr_src, int32_t* perm_src, void* workspace,
                             size_t workspace_bytes, int32_t* status, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (E < 0 || N < 0 || E >= (int64_t)1 << 31 || N >= (int64_t)1 << 31) return bad_arg("E/N out of int32 range");
  CsrWs w = carve(workspace, E, N);
  if (workspace_bytes < w.bytes) return bad_arg("csr workspace too small");
  cudaError_t e;
  if ((e = cudaMemsetAsync(w.deg_src, 0, sizeof(int32_t) * (N + 1), st)) != cudaSuccess) return fail("memset", e);
  if ((e = cudaMemsetAsync(w.deg_dst, 0, sizeof(int32_t) * (N + 1), st)) != cudaSuccess) return fail("memset", e);
  if ((e = cudaMemsetAsync(status, 0, sizeof(int32_t), st)) != cudaSuccess) return fail("memset", e);
  if (E > 0) {
    k_convert<<<(unsigned)ceil_div(E, 256), 256, 0, st>>>(edge_index, E, N, src32, dst32, w.deg_src, w.deg_dst, status);
    B3D_LAUNCH_CHECK("csr_convert");
  }
  k_exclusive_scan<<<1, 1024, 0, st>>>(w.deg_dst, rowptr_dst, N);
  B3D_LAUNCH_CHECK("scan_dst");
  k_exclusive_scan<<<1, 1024, 0, st>>>(w.deg_src, rowptr_src, N);
  B3D_LAUNCH_CHECK("scan_src");
  if (E > 0) {
    int r;
    if ((r = radix_sort_perm(dst32, E, N, perm_dst, w, st))) return r;
    if ((r = radix_sort_perm(src32, E, N, perm_src, w, st))) return r;
  }
  return 0;
}
