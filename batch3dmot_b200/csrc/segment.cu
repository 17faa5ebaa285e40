// Deterministic warp-per-node segmented reduction (replaces torch_scatter atomics),
// row gather (its backward) and the modality-present row predicate.
#include <cuda_bf16.h>

#include "b3d_common.cuh"

namespace b3d {

constexpr int SEG_WARPS = 8;
static bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// One warp per node; lanes cover columns (float4 when aligned); edges summed in ascending
// k = the order sequential CPU scatter_add_ uses, so results match index_add_ bit-for-bit.
template <bool VEC>
__global__ void __launch_bounds__(SEG_WARPS * 32) k_segment_sum(
    const float* __restrict__ src, int ld, const int32_t* __restrict__ perm,
    const int32_t* __restrict__ rowptr, long long N, int C, float* __restrict__ out, int ldo, int accumulate) {
  const int lane = threadIdx.x & 31;
  const long long node = (long long)blockIdx.x * SEG_WARPS + (threadIdx.x >> 5);
  if (node >= N) return;
  const int beg = __ldg(rowptr + node), end = __ldg(rowptr + node + 1);
  if (VEC) {
    for (int c = lane * 4; c < C; c += 128) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      int k = beg;
      for (; k + 4 <= end; k += 4) {  // 4 independent loads in flight, summed in order
        long long e0 = perm ? __ldg(perm + k) : k, e1 = perm ? __ldg(perm + k + 1) : k + 1;
        long long e2 = perm ? __ldg(perm + k + 2) : k + 2, e3 = perm ? __ldg(perm + k + 3) : k + 3;
        float4 v0 = __ldg(reinterpret_cast<const float4*>(src + e0 * ld + c));
        float4 v1 = __ldg(reinterpret_cast<const float4*>(src + e1 * ld + c));
        float4 v2 = __ldg(reinterpret_cast<const float4*>(src + e2 * ld + c));
        float4 v3 = __ldg(reinterpret_cast<const float4*>(src + e3 * ld + c));
        acc.x += v0.x; acc.y += v0.y; acc.z += v0.z; acc.w += v0.w;
        acc.x += v1.x; acc.y += v1.y; acc.z += v1.z; acc.w += v1.w;
        acc.x += v2.x; acc.y += v2.y; acc.z += v2.z; acc.w += v2.w;
        acc.x += v3.x; acc.y += v3.y; acc.z += v3.z; acc.w += v3.w;
      }
      for (; k < end; ++k) {
        long long e = perm ? __ldg(perm + k) : k;
        float4 v = __ldg(reinterpret_cast<const float4*>(src + e * ld + c));
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      float4* o = reinterpret_cast<float4*>(out + node * ldo + c);
      if (accumulate) { float4 p = *o; acc.x += p.x; acc.y += p.y; acc.z += p.z; acc.w += p.w; }
      *o = acc;
    }
  } else {
    for (int c = lane; c < C; c += 32) {
      float acc = 0.f;
      for (int k = beg; k < end; ++k) {
        long long e = perm ? __ldg(perm + k) : k;
        acc += __ldg(src + e * ld + c);
      }
      float* o = out + node * ldo + c;
      *o = accumulate ? *o + acc : acc;
    }
  }
}

// fp32 rows; MASK: zero where the fp32 ReLU activation `mask` (same shape as out) is <= 0 (backward of an fp32
// segment sum whose source was a ReLU output: gather + mask in one pass)
template <bool VEC, bool MASK>
__global__ void __launch_bounds__(SEG_WARPS * 32) k_gather_rows(
    const float* __restrict__ src, int ld, const int32_t* __restrict__ idx, long long M, int C,
    float* __restrict__ out, int ldo, const float* __restrict__ mask, int ldm) {
  const int lane = threadIdx.x & 31;
  const long long r = (long long)blockIdx.x * SEG_WARPS + (threadIdx.x >> 5);
  if (r >= M) return;
  const long long g = __ldg(idx + r);
  if (VEC) {
    for (int c = lane * 4; c < C; c += 128) {
      float4 v = __ldg(reinterpret_cast<const float4*>(src + g * ld + c));
      if (MASK) {
        const float4 m = __ldg(reinterpret_cast<const float4*>(mask + r * ldm + c));
        v.x = m.x > 0.f ? v.x : 0.f; v.y = m.y > 0.f ? v.y : 0.f; v.z = m.z > 0.f ? v.z : 0.f; v.w = m.w > 0.f ? v.w : 0.f;
      }
      *reinterpret_cast<float4*>(out + r * ldo + c) = v;
    }
  } else {
    for (int c = lane; c < C; c += 32) {
      float v = __ldg(src + g * ld + c);
      if (MASK) v = __ldg(mask + r * ldm + c) > 0.f ? v : 0.f;
      out[r * ldo + c] = v;
    }
  }
}

__global__ void __launch_bounds__(SEG_WARPS * 32) k_row_nonzero(const float* __restrict__ f, long long row_len,
                                                               long long N, uint8_t* __restrict__ mask) {
  const int lane = threadIdx.x & 31;
  const long long r = (long long)blockIdx.x * SEG_WARPS + (threadIdx.x >> 5);
  if (r >= N) return;
  const float* p = f + r * row_len;
  float s = 0.f;
  for (long long c = lane; c < row_len; c += 32) s += __ldg(p + c);
  s = warp_sum(s);
  if (lane == 0) mask[r] = (s != 0.f) ? 1 : 0;
}


// out[r, c] = sum_k in_k[r, c] over up to 8 row-major inputs (fp32 or bf16, each with its own leading
// dimension), summed in fp32 in argument order; one pass instead of the n-1 pairwise additions the
// autograd engine would run when a tensor feeds several consumers.
struct AddNArgs {
  SegDev in[B3D_MAX_SEGS];
  int n;
  long long M;
  int C;
  void* out;
  int out_bf16, ldo;
};

__device__ __forceinline__ void addn_load8(const SegDev& S, long long r, int c, float (&acc)[8]) {
  if (S.dtype == B3D_BF16) {
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(S.ptr) + r * S.ld + c));
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      acc[2 * j] += __uint_as_float(w[j] << 16);
      acc[2 * j + 1] += __uint_as_float(w[j] & 0xFFFF0000u);
    }
  } else {
    const float4 a = __ldg(reinterpret_cast<const float4*>(S.ptr + r * S.ld + c));
    const float4 b = __ldg(reinterpret_cast<const float4*>(S.ptr + r * S.ld + c + 4));
    acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w;
    acc[4] += b.x; acc[5] += b.y; acc[6] += b.z; acc[7] += b.w;
  }
}

__global__ void __launch_bounds__(256) k_add_n(const AddNArgs a) {
  const int groups = a.C / 8;
  const long long total = a.M * groups;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / groups;
    const int c = (int)(i % groups) * 8;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < a.n; ++k) addn_load8(a.in[k], r, c, acc);
    if (a.out_bf16) {
      uint4 q;
      __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&q);
#pragma unroll
      for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(acc[2 * j], acc[2 * j + 1]);
      *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(a.out) + r * a.ldo + c) = q;
    } else {
      float* o = reinterpret_cast<float*>(a.out) + r * a.ldo + c;
      *reinterpret_cast<float4*>(o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
      *reinterpret_cast<float4*>(o + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
  }
}

}  // namespace b3d

using namespace b3d;

// bf16 rows, fp32 accumulation: warp per node, 8 columns (16 B) per lane per pass.
__global__ void __launch_bounds__(SEG_WARPS * 32) k_segment_sum_bf16(
    const __nv_bfloat16* __restrict__ src, int ld, const int32_t* __restrict__ perm,
    const int32_t* __restrict__ rowptr, long long N, int C, float* __restrict__ out, int ldo, int accumulate,
    int out_bf16) {
  const int lane = threadIdx.x & 31;
  const long long node = (long long)blockIdx.x * SEG_WARPS + (threadIdx.x >> 5);
  if (node >= N) return;
  const int beg = __ldg(rowptr + node), end = __ldg(rowptr + node + 1);
  for (int c = lane * 8; c < C; c += 256) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int k = beg; k < end; ++k) {
      const long long e = perm ? __ldg(perm + k) : k;
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(src + e * ld + c));
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[2 * j] += __uint_as_float(w[j] << 16);
        acc[2 * j + 1] += __uint_as_float(w[j] & 0xFFFF0000u);
      }
    }
    if (out_bf16) {   // the consumer is a bf16 tensor-core tile: round once here instead of in its staging
      uint4 q;
      __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&q);
#pragma unroll
      for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(acc[2 * j], acc[2 * j + 1]);
      *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(out) + node * ldo + c) = q;
    } else {
      float* o = out + node * ldo + c;
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = accumulate ? o[j] + acc[j] : acc[j];
    }
  }
}

extern "C" int b3d_segment_sum(const void* src_v, int32_t src_dtype, int32_t ld_src, const int32_t* perm,
                               const int32_t* rowptr, int64_t N, int32_t C, float* out, int32_t ld_out,
                               int32_t flags, void* stream) {
  if (!rowptr || !out || C <= 0 || N < 0) return bad_arg("b3d_segment_sum");  // src may be NULL when E == 0
  if (N == 0) return 0;
  if (src_dtype == B3D_BF16) {
    if ((C & 7) || (ld_src & 7) || (reinterpret_cast<uintptr_t>(src_v) & 15))
      return bad_arg("b3d_segment_sum: bf16 rows need C % 8 == 0 and 16-byte alignment");
    const int obf = (flags & B3D_FLAG_OUT_BF16) ? 1 : 0;
    if (obf && ((flags & B3D_FLAG_ACCUMULATE) || (ld_out & 7) || (reinterpret_cast<uintptr_t>(out) & 15)))
      return bad_arg("b3d_segment_sum: bf16 output needs ld % 8 == 0, 16-byte alignment, no accumulate");
    k_segment_sum_bf16<<<(unsigned)ceil_div(N, SEG_WARPS), SEG_WARPS * 32, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(src_v), ld_src, perm, rowptr, N, C, out, ld_out,
        (flags & B3D_FLAG_ACCUMULATE) ? 1 : 0, obf);
    B3D_LAUNCH_CHECK("k_segment_sum_bf16");
    return 0;
  }
  if (flags & B3D_FLAG_OUT_BF16) return bad_arg("b3d_segment_sum: bf16 output needs a bf16 source");
  const float* src = reinterpret_cast<const float*>(src_v);
  bool vec = (C % 4 == 0) && (ld_src % 4 == 0) && (ld_out % 4 == 0) && al16(src) && al16(out);
  unsigned grid = (unsigned)ceil_div(N, SEG_WARPS);
  int acc = (flags & B3D_FLAG_ACCUMULATE) ? 1 : 0;
  if (vec) k_segment_sum<true><<<grid, SEG_WARPS * 32, 0, (cudaStream_t)stream>>>(src, ld_src, perm, rowptr, N, C, out, ld_out, acc);
  else k_segment_sum<false><<<grid, SEG_WARPS * 32, 0, (cudaStream_t)stream>>>(src, ld_src, perm, rowptr, N, C, out, ld_out, acc);
  B3D_LAUNCH_CHECK("k_segment_sum");
  return 0;
}

// Rows gathered into a bf16 matrix (gradient of a bf16 segment-sum input): out[r] = src[idx[r]], fp32 sources
// rounded on the way, optionally zeroed where the ReLU mask (bf16 activation or sign bits) says the summed
// output was <= 0. One warp handles GR_ROWS consecutive output rows, lanes cover 8-column chunks, and the
// index / row / mask loads of the GR_ROWS rows are issued together (one row at a time left a single
// dependent idx -> row -> store chain per warp: 1.9 TB/s).
constexpr int GR_ROWS = 8;
// SRC_BF16 / MASK (0 none, 1 bf16 activation, 2 sign bits) are compile-time so that only the registers of the
// path in use are allocated (the all-paths kernel needed ~90 registers: 16 warps per SM, latency-bound again).
// The 8-column chunks of the warp's GR_ROWS rows are numbered row after row and dealt to the lanes round-robin, so
// that all 32 lanes move data whatever C is (C = 192: 24 chunks per row left a quarter of the lanes idle), and every
// lane has its GR_ROWS * C / 256 independent gathers in flight together.
template <bool SRC_BF16, int MASK, int NIT>
__global__ void __launch_bounds__(SEG_WARPS * 32) k_gather_rows_bf16(
    const float* __restrict__ src, int ld, const int32_t* __restrict__ idx, long long M, int C,
    __nv_bfloat16* __restrict__ out, int ldo, const __nv_bfloat16* __restrict__ relu_mask, int ldm,
    const uint32_t* __restrict__ relu_bits) {
  const int lane = threadIdx.x & 31;
  const long long r0 = ((long long)blockIdx.x * SEG_WARPS + (threadIdx.x >> 5)) * GR_ROWS;
  if (r0 >= M) return;
  const int cpr = C >> 3;                                  // chunks per row
  const int gl = (lane < GR_ROWS && r0 + lane < M) ? __ldg(idx + r0 + lane) : -1;
  uint4 v[NIT], mk[NIT];
  float4 fa[NIT], fb[NIT];
  uint32_t word[NIT];
  int u_[NIT], c_[NIT];
#pragma unroll
  for (int it = 0; it < NIT; ++it) {
    const int q = it * 32 + lane;
    const int u = q / cpr, c = (q - u * cpr) * 8;
    const int g = __shfl_sync(0xffffffffu, gl, u & (GR_ROWS - 1));
    u_[it] = (u < GR_ROWS && g >= 0) ? u : -1;
    c_[it] = c;
    if (u_[it] < 0) continue;
    if (SRC_BF16) {   // already rounded by the producing input-gradient tile: a pure 16-byte copy
      v[it] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(src) + (long long)g * ld + c));
    } else {
      fa[it] = __ldg(reinterpret_cast<const float4*>(src + (long long)g * ld + c));
      fb[it] = __ldg(reinterpret_cast<const float4*>(src + (long long)g * ld + c) + 1);
    }
    if (MASK == 2) word[it] = __ldg(relu_bits + (long long)(c >> 5) * M + r0 + u) >> (c & 31);
    if (MASK == 1) mk[it] = __ldg(reinterpret_cast<const uint4*>(relu_mask + (r0 + u) * ldm + c));
  }
#pragma unroll
  for (int it = 0; it < NIT; ++it) {
    if (u_[it] < 0) continue;
    uint4 q;
    if (SRC_BF16) {
      q = v[it];
    } else {
      __nv_bfloat162 p0 = __floats2bfloat162_rn(fa[it].x, fa[it].y), p1 = __floats2bfloat162_rn(fa[it].z, fa[it].w);
      __nv_bfloat162 p2 = __floats2bfloat162_rn(fb[it].x, fb[it].y), p3 = __floats2bfloat162_rn(fb[it].z, fb[it].w);
      q = make_uint4(*reinterpret_cast<uint32_t*>(&p0), *reinterpret_cast<uint32_t*>(&p1),
                     *reinterpret_cast<uint32_t*>(&p2), *reinterpret_cast<uint32_t*>(&p3));
    }
    uint32_t* w = reinterpret_cast<uint32_t*>(&q);
    if (MASK == 1) {   // keep the gradient where the ReLU output was > 0 (bf16 > 0 <=> int16 bits > 0)
      const uint32_t mw[4] = {mk[it].x, mk[it].y, mk[it].z, mk[it].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if ((int16_t)(mw[j] & 0xFFFFu) <= 0) w[j] &= 0xFFFF0000u;
        if ((int16_t)(mw[j] >> 16) <= 0) w[j] &= 0x0000FFFFu;
      }
    }
    if (MASK == 2) {   // the same mask as sign bits: word [(c / 32) * M + r], bit c % 32
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (!((word[it] >> (2 * j)) & 1u)) w[j] &= 0xFFFF0000u;
        if (!((word[it] >> (2 * j + 1)) & 1u)) w[j] &= 0x0000FFFFu;
      }
    }
    *reinterpret_cast<uint4*>(out + (r0 + u_[it]) * ldo + c_[it]) = q;
  }
}

extern "C" int b3d_gather_rows(const float* src, int32_t ld_src, const int32_t* idx, int64_t M, int32_t C,
                               void* out_v, int32_t out_dtype, int32_t ld_out, const void* relu_mask,
                               int32_t ld_mask, int32_t mask_dtype, int32_t src_dtype, void* stream) {
  if (M == 0) return 0;
  if (!src || !idx || !out_v || C <= 0 || M < 0) return bad_arg("b3d_gather_rows");
  if (out_dtype == B3D_BF16) {
    if ((C & 7) || (ld_src & (src_dtype == B3D_BF16 ? 7 : 3)) || (ld_out & 7) || !al16(src) || !al16(out_v))
      return bad_arg("b3d_gather_rows: bf16 output needs C % 8 == 0 and 16-byte aligned rows");
    if (relu_mask && ((ld_mask & 7) || !al16(relu_mask))) return bad_arg("b3d_gather_rows: relu_mask alignment");
    const unsigned grid = (unsigned)ceil_div(M, (long long)SEG_WARPS * GR_ROWS);
    const int mk = !relu_mask ? 0 : (mask_dtype == B3D_BITS ? 2 : 1);
    const __nv_bfloat16* m16 = mk == 1 ? reinterpret_cast<const __nv_bfloat16*>(relu_mask) : nullptr;
    const uint32_t* mb = mk == 2 ? reinterpret_cast<const uint32_t*>(relu_mask) : nullptr;
    __nv_bfloat16* o16 = reinterpret_cast<__nv_bfloat16*>(out_v);
    cudaStream_t st = (cudaStream_t)stream;
    // NIT = chunk rounds per lane: GR_ROWS * (C / 8) chunks over 32 lanes (C <= 512 in one launch shape per NIT)
    const int nit = (GR_ROWS * (C >> 3) + 31) / 32;
    if (nit > 16) return bad_arg("b3d_gather_rows: C <= 512 for bf16 output");
#define B3D_GR3(SB, MK, NI) k_gather_rows_bf16<SB, MK, NI><<<grid, SEG_WARPS * 32, 0, st>>>(src, ld_src, idx, M, C, o16, ld_out, m16, ld_mask, mb)
#define B3D_GR(SB, MK) do { if (nit <= 2) B3D_GR3(SB, MK, 2); else if (nit <= 4) B3D_GR3(SB, MK, 4); else if (nit <= 6) B3D_GR3(SB, MK, 6); \
                            else if (nit <= 8) B3D_GR3(SB, MK, 8); else B3D_GR3(SB, MK, 16); } while (0)
    if (src_dtype == B3D_BF16) { if (mk == 0) B3D_GR(true, 0); else if (mk == 1) B3D_GR(true, 1); else B3D_GR(true, 2); }
    else { if (mk == 0) B3D_GR(false, 0); else if (mk == 1) B3D_GR(false, 1); else B3D_GR(false, 2); }
#undef B3D_GR
#undef B3D_GR3
    B3D_LAUNCH_CHECK("k_gather_rows_bf16");
    return 0;
  }
  if (src_dtype == B3D_BF16) return bad_arg("b3d_gather_rows: a bf16 source needs bf16 output");
  if (relu_mask && mask_dtype != B3D_F32) return bad_arg("b3d_gather_rows: fp32 output takes an fp32 relu_mask");
  float* out = reinterpret_cast<float*>(out_v);
  const float* mk = reinterpret_cast<const float*>(relu_mask);
  bool vec = (C % 4 == 0) && (ld_src % 4 == 0) && (ld_out % 4 == 0) && al16(src) && al16(out) &&
             (!mk || ((ld_mask % 4 == 0) && al16(mk)));
  unsigned grid = (unsigned)ceil_div(M, SEG_WARPS);
  cudaStream_t st = (cudaStream_t)stream;
  if (vec && mk) k_gather_rows<true, true><<<grid, SEG_WARPS * 32, 0, st>>>(src, ld_src, idx, M, C, out, ld_out, mk, ld_mask);
  else if (vec) k_gather_rows<true, false><<<grid, SEG_WARPS * 32, 0, st>>>(src, ld_src, idx, M, C, out, ld_out, nullptr, 0);
  else if (mk) k_gather_rows<false, true><<<grid, SEG_WARPS * 32, 0, st>>>(src, ld_src, idx, M, C, out, ld_out, mk, ld_mask);
  else k_gather_rows<false, false><<<grid, SEG_WARPS * 32, 0, st>>>(src, ld_src, idx, M, C, out, ld_out, nullptr, 0);
  B3D_LAUNCH_CHECK("k_gather_rows");
  return 0;
}

extern "C" int b3d_row_nonzero(const float* feats, int64_t row_len, int64_t N, uint8_t* mask, void* stream) {
  if (!feats || !mask || row_len <= 0 || N < 0) return bad_arg("b3d_row_nonzero");
  if (N == 0) return 0;
  k_row_nonzero<<<(unsigned)ceil_div(N, SEG_WARPS), SEG_WARPS * 32, 0, (cudaStream_t)stream>>>(feats, row_len, N, mask);
  B3D_LAUNCH_CHECK("k_row_nonzero");
  return 0;
}

extern "C" int b3d_add_n(const b3d_seg_t* ins, int32_t n, int64_t M, void* out, int32_t out_dtype, int32_t ldo,
                         void* stream) {
  AddNArgs a;
  if (!out || M < 0 || to_dev(ins, n, a.in)) return bad_arg("b3d_add_n");
  if (M == 0) return 0;
  a.n = n; a.M = M; a.C = a.in[0].width; a.out = out; a.out_bf16 = (out_dtype == B3D_BF16); a.ldo = ldo;
  if (a.C % 8 || !al16(out) || (ldo % (a.out_bf16 ? 8 : 4))) return bad_arg("b3d_add_n: width % 8, 16-byte aligned rows");
  for (int k = 0; k < n; ++k)
    if (a.in[k].width != a.C || a.in[k].idx || a.in[k].mask_mode != B3D_MASK_NONE || !al16(a.in[k].ptr) ||
        (a.in[k].ld % (a.in[k].dtype == B3D_BF16 ? 8 : 4)))
      return bad_arg("b3d_add_n: inputs must be dense, equally wide, 16-byte aligned rows");
  const long long total = M * (a.C / 8);
  long long blocks = ceil_div(total, 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  k_add_n<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a);
  B3D_LAUNCH_CHECK("k_add_n");
  return 0;
}
