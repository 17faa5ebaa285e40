// bf16 tensor-core ("2e-2 mode") dense layers for sm_100a: tcgen05.mma with fp32 accumulators
// in TMEM, operands staged in shared memory in the no-swizzle canonical UMMA layouts.
//
//   b3d_linear_tc : Y = act(cat_s(gather(A_s)) W^T + b) [* (out_mask > 0)]  (forward and, with a
//                   transposed pack, the input gradient). A segments: fp32 (converted to bf16
//                   while staging) or bf16 (cp.async), K-major canonical layout; W: pre-packed
//                   bf16 slabs (b3d_tc_pack_weights). Y: fp32 or bf16.
//   b3d_wgrad_tc  : dW = dY^T cat_s(A_s), db = colsum(dY). Both operands staged MN-major (rows are
//                   the reduction dimension), rows split over CTAs, fixed-order second pass.
//
// Pipeline (both kernels): 256 threads stage chunk c+1 while the tensor core works on chunk c
// (tcgen05.mma is asynchronous; tcgen05.commit -> mbarrier frees the stage). Two CTAs per SM
// overlap one CTA's epilogue with the other's main loop.
//
//   b3d_linear_tma / b3d_wgrad_tma (second half of the file): the TMA-fed persistent kernels that run the dense bf16
//                   layers of the bf16 mode — operand tiles by cp.async.bulk.tensor (128-byte swizzle), the weight
//                   block resident in shared memory, double-buffered TMEM accumulators; row-gathered bf16 addends are
//                   staged by producer warps (cp.async) into operand-shaped tiles and added to the accumulator by
//                   identity MMAs; lean epilogue block for bf16 outputs; see the comment above k_linear_tma.
//   SPLIT variants of the thread-staged kernels: the 1e-4 mode (tf32 x3 layers, split-bf16 weight gradients).
#include <cuda.h>
#include <stdlib.h>

#include "b3d_common.cuh"
#include "tc_common.cuh"

namespace b3d {

constexpr int TC_BM = 128;      // rows per CTA (UMMA M)
constexpr int TC_BK = 64;       // K elements staged per chunk (4 MMAs of K=16)
constexpr int TC_NMAX = 256;    // max UMMA N per CTA
constexpr int TC_THREADS = 256;
constexpr int TC_A_STAGE = TC_BM * TC_BK * 2;  // 16 KB
constexpr int TC_TAB = 256;     // max 8-column groups of a concatenated operand (K <= 2048)

__host__ __device__ inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
__host__ __device__ inline uint32_t tmem_cols_for(int n) { return n <= 32 ? 32u : n <= 64 ? 64u : n <= 128 ? 128u : 256u; }

// ------------------------------------------------------------------ weight packing
// Wp[((k/8) * Npad + n) * 8 + k%8] = bf16(B[n][k]),  B[n][k] = transpose ? W[k*ldw+n] : W[n*ldw+k],
// zero padded to Npad = round_up(n_logical,16), Kpad = round_up(k_logical,64): every (chunk, k-group)
// slab is a contiguous run of 16-byte rows == the shared-memory image of a K-major operand.
// split != 0: a second image of the same shape follows the first and holds the bf16 residual
// lo = bf16(w - float(hi)) — operands of the 3-MMA "split bf16" mode (hi*hi + lo*hi + hi*lo ~ fp32 products).
__global__ void k_pack_weights(const float* __restrict__ W, int ldw, int n_log, int k_log, int transpose,
                               __nv_bfloat16* __restrict__ Wp, int Npad, int Kpad, int split = 0) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)Npad * Kpad;
  if (i >= total) return;
  int j = (int)(i & 7);
  long long q = i >> 3;
  int n = (int)(q % Npad);
  int kg = (int)(q / Npad);
  int k = kg * 8 + j;
  float v = 0.f;
  if (n < n_log && k < k_log) v = transpose ? W[(long long)k * ldw + n] : W[(long long)n * ldw + k];
  const __nv_bfloat16 hi = __float2bfloat16_rn(v);
  Wp[i] = hi;
  if (split) Wp[total + i] = __float2bfloat16_rn(v - __bfloat162float(hi));
}

// tf32 "x3" pack for the 1e-4 parity mode: 4-element (16-byte) K groups of fp32 words,
//   Wt[((k/4) * Npad + n) * 4 + k%4] = hi(B[n][k]) (low 13 mantissa bits cleared = what the tensor core reads),
// followed by a second image with the exact fp32 residuals lo = w - hi. Kpad = round_up(k_logical, 32).
__global__ void k_pack_weights_tf32(const float* __restrict__ W, int ldw, int n_log, int k_log, int transpose,
                                    float* __restrict__ Wp, int Npad, int Kpad) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)Npad * Kpad;
  if (i >= total) return;
  int j = (int)(i & 3);
  long long q = i >> 2;
  int n = (int)(q % Npad);
  int k = (int)(q / Npad) * 4 + j;
  float v = 0.f;
  if (n < n_log && k < k_log) v = transpose ? W[(long long)k * ldw + n] : W[(long long)n * ldw + k];
  const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
  Wp[i] = hi;
  Wp[total + i] = v - hi;
}

// (segment, offset) of every 8-column group of the concatenated operand; -1 past the end.
__device__ __forceinline__ void build_group_table(const SegDev* seg, int nseg, int ngroups, int32_t* tab, int tid,
                                                  int nthreads, int gw = 8) {
  for (int g = tid; g < ngroups; g += nthreads) {
    int off = g * gw, sg = 0;
    while (sg < nseg && off >= seg[sg].width) { off -= seg[sg].width; ++sg; }
    tab[g] = (sg < nseg) ? ((sg << 24) | off) : -1;
  }
}

__device__ __forceinline__ const void* seg_addr(const SegDev& S, long long row, int off) {
  return S.dtype == B3D_BF16
             ? static_cast<const void*>(reinterpret_cast<const __nv_bfloat16*>(S.ptr) + row * S.ld + off)
             : static_cast<const void*>(S.ptr + row * S.ld + off);
}

// ------------------------------------------------------------------ forward / dgrad
struct TcArgs {
  SegDev seg[B3D_MAX_SEGS];
  int nseg, Ktot;
  const __nv_bfloat16* Wp;
  int Npad, Kpad;
  const float* bias;
  void* Y;
  int ldy, y_bf16;
  long long M;
  int Nout, act, flags;
  const void* out_mask;
  int ldm, mask_bf16;
  const uint8_t* row_mask;
  SegDev add[2];   // row-gathered fp32 addends of the pre-activation
  int nadd;
  const uint32_t* mask_bits;   // ReLU mask as sign bits: word [(col / 32) * M + row], bit col % 32
  uint32_t* bits_out;          // same layout, written for this layer's (post-activation) output
  int stage_mask = 0;          // (TMA kernel only) addends staged through shared memory
};

template <int ACT>
__device__ __forceinline__ float activate(float v) {
  if (ACT == B3D_ACT_RELU) return fmaxf(v, 0.f);
  if (ACT == B3D_ACT_SIGMOID) return 1.f / (1.f + expf(-v));
  return v;
}

// One 32-column block of one output row: bias + row-gathered addends + activation, then the ReLU
// mask of the producing layer / row mask / accumulate, then a bf16 or fp32 store.
// EP is any struct with the TcArgs epilogue fields; sb = bias of this CTA's column block (smem).
// Global operands of one epilogue block (ReLU mask, gathered addends; bf16 fast paths only), loaded
// BEFORE the TMEM load is waited on so that the two latencies overlap.
// 256-bit global accesses (sm_100: LDG.256 / STG.256). The epilogue is row-per-lane, so every lane of
// a warp-wide access lands in a different 128-byte line and costs its own L1 wavefront whatever its
// width: moving 32 bytes per lane instead of 16 halves the L1 data-pipe load of the epilogue, which
// ncu shows as the top limiter of the mask / addend variants (profiles/r1_epilogue_l1.md).
struct EpiPrefetch {
  uint4 u[4], v[4];   // u: ReLU mask if bit0, else addend 0;  v: addend 0 if bit0, else addend 1
  uint32_t bits;      // sign-bit mask word of the block (bit3)
  int flags;          // bit0: mask prefetched, bit1: addend 0 prefetched, bit2: addend 1 prefetched, bit3: bits
};
// g0 / g1: source rows of the two addends for output row `row` (already gathered through add[t].idx).
template <class EP>
__device__ __forceinline__ void epilogue_prefetch(const EP& a, long long row, int g0, int g1, int cbase, int nlim,
                                                  EpiPrefetch& pf) {
  pf.flags = 0;
  if (a.mask_bits && (cbase & 31) == 0 && cbase < nlim) {   // one coalesced word per row
    pf.bits = __ldg(a.mask_bits + (long long)(cbase >> 5) * a.M + row);
    pf.flags |= 8;
  }
  if (cbase + 31 >= nlim) return;
  if (a.out_mask && a.mask_bf16 && (a.ldm & 7) == 0) {
    ld64B(reinterpret_cast<const __nv_bfloat16*>(a.out_mask) + row * a.ldm + cbase, pf.u);
    pf.flags |= 1;
  }
  if (a.nadd > 0 && a.add[0].dtype == B3D_BF16 && !(a.stage_mask & 1)) {
    const SegDev& S = a.add[0];
    const __nv_bfloat16* ap = reinterpret_cast<const __nv_bfloat16*>(S.ptr) + (long long)g0 * S.ld + cbase;
    if (pf.flags & 1) ld64B(ap, pf.v); else ld64B(ap, pf.u);
    pf.flags |= 2;
  }
  if (a.nadd > 1 && a.add[1].dtype == B3D_BF16 && !(pf.flags & 1) && !(a.stage_mask & 2)) {
    const SegDev& S = a.add[1];
    ld64B(reinterpret_cast<const __nv_bfloat16*>(S.ptr) + (long long)g1 * S.ld + cbase, pf.v);
    pf.flags |= 4;
  }
}

// Addends whose bit is set in a.stage_mask never reach the epilogue: the MMA warp adds their staged tiles to the
// accumulator (k_linear_tma).
template <int ACT, class EP>
__device__ __forceinline__ void epilogue_block32(const EP& a, long long row, int n0, int col0, const uint32_t (&r)[32],
                                                 const float* s_bias, bool plain, bool rz, int nlim,
                                                 const EpiPrefetch* pf = nullptr) {
  // nlim: first global column this CTA must NOT write (min(Nout, end of its column block))
  using namespace tc;
    float o[32];
  const int cbase = n0 + col0;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 b4 = *reinterpret_cast<const float4*>(s_bias + col0 + 4 * q);
    o[4 * q + 0] = __uint_as_float(r[4 * q + 0]) + b4.x;
    o[4 * q + 1] = __uint_as_float(r[4 * q + 1]) + b4.y;
    o[4 * q + 2] = __uint_as_float(r[4 * q + 2]) + b4.z;
    o[4 * q + 3] = __uint_as_float(r[4 * q + 3]) + b4.w;
  }
  for (int t = 0; t < a.nadd; ++t) {   // node-side first-layer blocks, pre-projected per node
    const SegDev& S = a.add[t];
    if ((a.stage_mask >> t) & 1) continue;   // already in the accumulator
    if (pf && (pf->flags & (2 << t))) {   // already in registers
      const uint4* src = (t == 0 && !(pf->flags & 1)) ? pf->u : pf->v;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t w[4] = {src[q].x, src[q].y, src[q].z, src[q].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          o[8 * q + 2 * j] += __uint_as_float(w[j] << 16);
          o[8 * q + 2 * j + 1] += __uint_as_float(w[j] & 0xFFFF0000u);
        }
      }
      continue;
    }
    const long long arow = (long long)(S.idx ? __ldg(S.idx + row) : (int32_t)row) * S.ld + cbase;
    if (S.dtype == B3D_BF16) {
      const __nv_bfloat16* ap = reinterpret_cast<const __nv_bfloat16*>(S.ptr) + arow;
      if (cbase + 31 < nlim) {
        uint4 av[4];
        ld64B(ap, av);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint4 v = av[q];
          const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            o[8 * q + 2 * j] += __uint_as_float(w[j] << 16);
            o[8 * q + 2 * j + 1] += __uint_as_float(w[j] & 0xFFFF0000u);
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (cbase + j < nlim) o[j] += __bfloat162float(ap[j]);
      }
    } else {
      const float* ap = S.ptr + arow;
      if (cbase + 31 < nlim) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(ap) + q);
          o[4 * q] += v.x; o[4 * q + 1] += v.y; o[4 * q + 2] += v.z; o[4 * q + 3] += v.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (cbase + j < nlim) o[j] += __ldg(ap + j);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 32; ++j) o[j] = activate<ACT>(o[j]);
  if (a.bits_out && (cbase & 31) == 0 && cbase < nlim) {   // sign bits of this output block for the backward pass
    uint32_t word = 0u;
#pragma unroll
    for (int j = 0; j < 32; ++j) word |= (o[j] > 0.f && cbase + j < nlim) ? (1u << j) : 0u;
    a.bits_out[(long long)(cbase >> 5) * a.M + row] = word;
  }
  if (!plain) {
    if (a.mask_bits && (cbase & 31) == 0 && cbase < nlim) {   // ReLU backward from the producing layer's sign bits
      const uint32_t word = (pf && (pf->flags & 8)) ? pf->bits : __ldg(a.mask_bits + (long long)(cbase >> 5) * a.M + row);
#pragma unroll
      for (int j = 0; j < 32; ++j) o[j] = ((word >> j) & 1u) ? o[j] : 0.f;
    }
    if (a.out_mask) {   // ReLU backward of the producing layer: keep the gradient where its output was > 0
      if (a.mask_bf16 && cbase + 31 < nlim && (a.ldm & 7) == 0) {
        const bool pre = pf && (pf->flags & 1);
        uint4 mv4[4];
        if (!pre) ld64B(reinterpret_cast<const __nv_bfloat16*>(a.out_mask) + row * a.ldm + cbase, mv4);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint4 m = pre ? pf->u[q] : mv4[q];
          const uint32_t w[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if ((int16_t)(w[j] & 0xFFFFu) <= 0) o[8 * q + 2 * j] = 0.f;      // bf16 > 0  <=>  int16 bits > 0
            if ((int16_t)(w[j] >> 16) <= 0) o[8 * q + 2 * j + 1] = 0.f;
          }
        }
      } else if (!a.mask_bf16 && cbase + 31 < nlim && (a.ldm & 3) == 0) {
        const float4* mp = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(a.out_mask) + row * a.ldm + cbase);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 m = __ldg(mp + q);
          if (!(m.x > 0.f)) o[4 * q] = 0.f;
          if (!(m.y > 0.f)) o[4 * q + 1] = 0.f;
          if (!(m.z > 0.f)) o[4 * q + 2] = 0.f;
          if (!(m.w > 0.f)) o[4 * q + 3] = 0.f;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int cc = cbase + j;
          if (cc < nlim) {
            const float mv = a.mask_bf16
                ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(a.out_mask)[row * a.ldm + cc])
                : reinterpret_cast<const float*>(a.out_mask)[row * a.ldm + cc];
            o[j] = mv > 0.f ? o[j] : 0.f;
          }
        }
      }
    }
    if (rz) {
#pragma unroll
      for (int j = 0; j < 32; ++j) o[j] = 0.f;
    }
    if ((a.flags & B3D_FLAG_ACCUMULATE) && !a.y_bf16) {
      const float* yrow = reinterpret_cast<const float*>(a.Y) + row * a.ldy;
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (cbase + j < nlim) o[j] += yrow[cbase + j];
    }
  }
  if (a.y_bf16) {
    __nv_bfloat16* yrow = reinterpret_cast<__nv_bfloat16*>(a.Y) + row * a.ldy + cbase;
    if (cbase + 31 < nlim && al32(yrow)) {
      uint4 pk[4];
#pragma unroll
      for (int q = 0; q < 4; ++q)
        pk[q] = make_uint4(pack_bf16x2(o[8 * q], o[8 * q + 1]), pack_bf16x2(o[8 * q + 2], o[8 * q + 3]),
                           pack_bf16x2(o[8 * q + 4], o[8 * q + 5]), pack_bf16x2(o[8 * q + 6], o[8 * q + 7]));
      stg256(yrow, pk[0], pk[1]);
      stg256(yrow + 16, pk[2], pk[3]);
      return;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (cbase + 8 * q + 7 < nlim) {
        uint4 pk = make_uint4(pack_bf16x2(o[8 * q], o[8 * q + 1]), pack_bf16x2(o[8 * q + 2], o[8 * q + 3]),
                              pack_bf16x2(o[8 * q + 4], o[8 * q + 5]), pack_bf16x2(o[8 * q + 6], o[8 * q + 7]));
        *reinterpret_cast<uint4*>(yrow + 8 * q) = pk;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (cbase + 8 * q + j < nlim) yrow[8 * q + j] = __float2bfloat16_rn(o[8 * q + j]);
      }
    }
  } else {
    float* yrow = reinterpret_cast<float*>(a.Y) + row * a.ldy + cbase;
    if (cbase + 31 < nlim && al32(yrow)) {
#pragma unroll
      for (int q = 0; q < 4; ++q)
        stg256(yrow + 8 * q,
               make_uint4(__float_as_uint(o[8 * q]), __float_as_uint(o[8 * q + 1]), __float_as_uint(o[8 * q + 2]), __float_as_uint(o[8 * q + 3])),
               make_uint4(__float_as_uint(o[8 * q + 4]), __float_as_uint(o[8 * q + 5]), __float_as_uint(o[8 * q + 6]), __float_as_uint(o[8 * q + 7])));
      return;
    }
    const bool vec_ok = ((a.ldy & 3) == 0) && ((reinterpret_cast<uintptr_t>(a.Y) & 15) == 0);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      if (vec_ok && cbase + 4 * q + 3 < nlim) {
        *reinterpret_cast<float4*>(yrow + 4 * q) = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (cbase + 4 * q + j < nlim) yrow[4 * q + j] = o[4 * q + j];
      }
    }
  }
}

// The common bf16 block of the persistent TMA kernel, with everything the generic block decides per element hoisted
// out: a full, 32-byte aligned block of a bf16 output whose addends (if any) are already in the accumulator, no row
// mask, no bf16 activation mask, no accumulate. ~4 instructions per element instead of ~11: integer max as ReLU
// (exact for every non-NaN input, -0 included), the sign-bit word built with one subtract + one funnel shift per
// element (x > 0 <=> the integer 0 - bits(x) is negative, for x >= 0), bias only when the layer has one.
template <int ACT, bool BIAS, bool MASK, bool BITS>
__device__ __forceinline__ void epilogue_fast32(const uint32_t (&r)[32], const float* sb, uint32_t* bits_out_word,
                                                uint32_t mword, __nv_bfloat16* yrow) {
  using namespace tc;
  uint32_t o[32];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    if (BIAS) {
      const float4 b4 = *reinterpret_cast<const float4*>(sb + 4 * q);
      o[4 * q + 0] = __float_as_uint(__uint_as_float(r[4 * q + 0]) + b4.x);
      o[4 * q + 1] = __float_as_uint(__uint_as_float(r[4 * q + 1]) + b4.y);
      o[4 * q + 2] = __float_as_uint(__uint_as_float(r[4 * q + 2]) + b4.z);
      o[4 * q + 3] = __float_as_uint(__uint_as_float(r[4 * q + 3]) + b4.w);
    } else {
      o[4 * q + 0] = r[4 * q + 0]; o[4 * q + 1] = r[4 * q + 1]; o[4 * q + 2] = r[4 * q + 2]; o[4 * q + 3] = r[4 * q + 3];
    }
  }
  if (ACT == B3D_ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) o[j] = (uint32_t)max((int)o[j], 0);
  }
  if (BITS) {
    uint32_t word = 0u;
    if (ACT == B3D_ACT_RELU) {
#pragma unroll
      for (int j = 31; j >= 0; --j) word = __funnelshift_l(0u - o[j], word, 1);   // (word << 1) | (o[j] != 0)
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) word |= (__uint_as_float(o[j]) > 0.f) ? (1u << j) : 0u;
    }
    *bits_out_word = word;
  }
  if (MASK) {
#pragma unroll
    for (int j = 0; j < 32; ++j)   // test-bit-into-predicate + select: 2 instructions per element
      asm("{\n\t.reg .pred p;\n\t.reg .b32 t;\n\tand.b32 t, %1, %2;\n\tsetp.ne.b32 p, t, 0;\n\tselp.b32 %0, %0, 0, p;\n\t}"
          : "+r"(o[j]) : "r"(mword), "r"(1u << j));
  }
  uint4 pk[4];
#pragma unroll
  for (int q = 0; q < 4; ++q)
    pk[q] = make_uint4(pack_bf16x2(__uint_as_float(o[8 * q]), __uint_as_float(o[8 * q + 1])),
                       pack_bf16x2(__uint_as_float(o[8 * q + 2]), __uint_as_float(o[8 * q + 3])),
                       pack_bf16x2(__uint_as_float(o[8 * q + 4]), __uint_as_float(o[8 * q + 5])),
                       pack_bf16x2(__uint_as_float(o[8 * q + 6]), __uint_as_float(o[8 * q + 7])));
  stg256(yrow, pk[0], pk[1]);
  stg256(yrow + 16, pk[2], pk[3]);
}

// split helpers. bf16 pair (weight-gradient kernel): v = hi + lo with hi = bf16(v), lo = bf16(v - hi).
__device__ __forceinline__ void split_pack2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat16 ha = __float2bfloat16_rn(a), hb = __float2bfloat16_rn(b);
  hi = (uint32_t)__bfloat16_as_ushort(ha) | ((uint32_t)__bfloat16_as_ushort(hb) << 16);
  lo = tc::pack_bf16x2(a - __bfloat162float(ha), b - __bfloat162float(hb));
}
// tf32 pair (layer kernel): hi = v with the 13 low mantissa bits cleared (exactly what kind::tf32 reads), lo = v - hi
// (exact in fp32; the tensor core again reads its top 11 significant bits): 22 significant bits per operand, so
// hi*hi + lo*hi + hi*lo reproduces the fp32 product to ~2^-21 relative. Accumulation is fp32 in TMEM.
__device__ __forceinline__ void split_tf32(const float4& v, uint4& hi, uint4& lo) {
  hi = make_uint4(__float_as_uint(v.x) & 0xFFFFE000u, __float_as_uint(v.y) & 0xFFFFE000u,
                  __float_as_uint(v.z) & 0xFFFFE000u, __float_as_uint(v.w) & 0xFFFFE000u);
  lo = make_uint4(__float_as_uint(v.x - __uint_as_float(hi.x)), __float_as_uint(v.y - __uint_as_float(hi.y)),
                  __float_as_uint(v.z - __uint_as_float(hi.z)), __float_as_uint(v.w - __uint_as_float(hi.w)));
}

// SPLIT ("tf32 x3"): operands are 32-bit words; a K chunk is 32 elements (the same 128 bytes per row, 8 groups of
// 16 bytes = 4 elements), every operand stage holds a hi image followed by a lo image (A: +TC_A_STAGE, B: +b_stage)
// and each K step (8 elements = 2 groups) issues three kind::tf32 MMAs. This is the tensor-core path of the 1e-4
// ("fp32") parity mode.
template <int ACT, bool SPLIT>
__global__ void __launch_bounds__(TC_THREADS) k_linear_tc(const TcArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  using namespace tc;
  constexpr int NI = SPLIT ? 2 : 1;                      // images per operand stage
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long m0 = (long long)blockIdx.x * TC_BM;
  const int n0 = blockIdx.y * TC_NMAX;
  const int Nb = min(TC_NMAX, a.Npad - n0);
  const uint32_t b_stage = (uint32_t)Nb * (TC_BK * 2);   // Nb rows x 128 B
  const uint32_t base_off = NI * (2 * TC_A_STAGE + 2 * b_stage);
  const uint32_t sA = smem_u32(smem);
  const uint32_t sB = sA + NI * 2 * TC_A_STAGE;
  const uint32_t sBar = sA + base_off;                    // free[0] @0, free[1] @8, done @16, tmem ptr @24
  volatile uint32_t* s_tmem = reinterpret_cast<volatile uint32_t*>(smem + base_off + 24);
  float* s_bias = reinterpret_cast<float*>(smem + base_off + 64);                       // [256]
  int32_t* s_tab = reinterpret_cast<int32_t*>(smem + base_off + 64 + 1024);            // [TC_TAB]
  int32_t* s_grow = reinterpret_cast<int32_t*>(smem + base_off + 64 + 1024 + 4 * TC_TAB);  // [nseg][128]
  const uint32_t ncols = tmem_cols_for(Nb);

  if (warp == 0) tmem_alloc(sBar + 24, ncols);
  if (tid == 0) {
    mbar_init(sBar + 0, 1);
    mbar_init(sBar + 8, 1);
    mbar_init(sBar + 16, 1);
    fence_mbar_init();
  }
  const int arow = tid & (TC_BM - 1), ahalf = tid >> 7;   // A staging role: row, which 4 of the 8 groups
  const long long grow_row = m0 + arow;
  const bool arow_ok = grow_row < a.M;
  if (tid < TC_BM)
    for (int s = 0; s < a.nseg; ++s)
      s_grow[s * TC_BM + tid] = arow_ok ? (a.seg[s].idx ? __ldg(a.seg[s].idx + grow_row) : (int32_t)grow_row) : 0;
  if (tid < Nb) s_bias[tid] = (a.bias && n0 + tid < a.Nout) ? __ldg(a.bias + n0 + tid) : 0.f;
  build_group_table(a.seg, a.nseg, a.Kpad / (SPLIT ? 4 : 8), s_tab, tid, TC_THREADS, SPLIT ? 4 : 8);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *s_tmem;

  const int nchunks = a.Kpad / (SPLIT ? 32 : TC_BK);
  const uint32_t idesc = SPLIT ? make_idesc_tf32(TC_BM, (uint32_t)Nb, 0, 0) : make_idesc_bf16(TC_BM, (uint32_t)Nb, 0, 0);
  const uint32_t lbo_a = TC_BM * 16, lbo_b = (uint32_t)Nb * 16, sbo = 128;

  for (int c = 0; c < nchunks; ++c) {
    const int s = c & 1;
    if (c >= 2) mbar_wait(sBar + 8 * s, ((c >> 1) - 1) & 1);   // MMAs of chunk c-2 released this stage
    // ---- B: contiguous pre-packed slabs -> cp.async (SPLIT: the lo image lies Npad*Kpad elements further)
    {
      const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(a.Wp) + ((size_t)(c * 8) * a.Npad + n0) * 16;
      const uint32_t dst = sB + s * NI * b_stage;
      if (tid < Nb) {
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          cp_async16(dst + (uint32_t)(g * Nb + tid) * 16, wsrc + ((size_t)g * a.Npad + tid) * 16);
          if (SPLIT)
            cp_async16(dst + b_stage + (uint32_t)(g * Nb + tid) * 16,
                       wsrc + (size_t)a.Npad * a.Kpad * 4 + ((size_t)g * a.Npad + tid) * 16);
        }
      }
    }
    // ---- A: thread = (row, half): 4 groups of 8 columns; bf16 sources by cp.async, fp32 via registers
    {
      float4 v[8];
      int ent[4];
      const uint32_t dst0 = sA + s * NI * TC_A_STAGE + arow * 16;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int g = ahalf * 4 + j;
        ent[j] = arow_ok ? s_tab[c * 8 + g] : -1;
        if (ent[j] >= 0) {
          const SegDev& S = a.seg[ent[j] >> 24];
          const long long gr = s_grow[(ent[j] >> 24) * TC_BM + arow];
          const void* p = seg_addr(S, gr, ent[j] & 0xFFFFFF);
          if (SPLIT) {
            if (S.dtype == B3D_BF16) {      // 4 bf16 -> 4 exact tf32 words (8 mantissa bits), zero residual
              const uint2 q = __ldg(reinterpret_cast<const uint2*>(p));
              v[2 * j] = make_float4(__uint_as_float(q.x << 16), __uint_as_float(q.x & 0xFFFF0000u),
                                     __uint_as_float(q.y << 16), __uint_as_float(q.y & 0xFFFF0000u));
            } else {
              v[2 * j] = __ldg(reinterpret_cast<const float4*>(p));
            }
          } else if (S.dtype == B3D_BF16) {
            cp_async16(dst0 + g * (TC_BM * 16), p);
            ent[j] = -2;   // done
          } else {
            v[2 * j] = __ldg(reinterpret_cast<const float4*>(p));
            v[2 * j + 1] = __ldg(reinterpret_cast<const float4*>(p) + 1);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int g = ahalf * 4 + j;
        const uint32_t d = dst0 + g * (TC_BM * 16);
        if (ent[j] >= 0) {
          if (SPLIT) {
            uint4 h, l;
            split_tf32(v[2 * j], h, l);
            st_shared_v4(d, h.x, h.y, h.z, h.w);
            st_shared_v4(d + TC_A_STAGE, l.x, l.y, l.z, l.w);
          } else {
            st_shared_v4(d, pack_bf16x2(v[2 * j].x, v[2 * j].y), pack_bf16x2(v[2 * j].z, v[2 * j].w),
                         pack_bf16x2(v[2 * j + 1].x, v[2 * j + 1].y), pack_bf16x2(v[2 * j + 1].z, v[2 * j + 1].w));
          }
        } else {
          if (ent[j] == -1) st_shared_v4(d, 0u, 0u, 0u, 0u);
          if (SPLIT && ent[j] == -1) st_shared_v4(d + TC_A_STAGE, 0u, 0u, 0u, 0u);
        }
      }
    }
    cp_async_wait_all();
    fence_proxy_async_smem();      // generic-proxy smem writes -> visible to the tensor core (async proxy)
    __syncthreads();
    if (tid == 0) {
      tc_fence_after_sync();
#pragma unroll
      for (int j = 0; j < TC_BK / 16; ++j) {        // 4 K steps per chunk: 16 bf16 or 8 tf32 elements = 2 groups each
        const uint64_t ad = make_smem_desc(sA + s * NI * TC_A_STAGE + j * 2 * lbo_a, lbo_a, sbo);
        const uint64_t bd = make_smem_desc(sB + s * NI * b_stage + j * 2 * lbo_b, lbo_b, sbo);
        if (SPLIT) {
          const uint64_t al = make_smem_desc(sA + s * NI * TC_A_STAGE + TC_A_STAGE + j * 2 * lbo_a, lbo_a, sbo);
          const uint64_t bl = make_smem_desc(sB + s * NI * b_stage + b_stage + j * 2 * lbo_b, lbo_b, sbo);
          mma_tf32_ss(tmem, ad, bd, idesc, (c | j) != 0);
          mma_tf32_ss(tmem, al, bd, idesc, 1);
          mma_tf32_ss(tmem, ad, bl, idesc, 1);
        } else {
          mma_bf16_ss(tmem, ad, bd, idesc, (c | j) != 0);
        }
      }
      mma_commit(sBar + 8 * s);
      if (c == nchunks - 1) mma_commit(sBar + 16);
    }
  }
  mbar_wait(sBar + 16, 0);
  tc_fence_after_sync();

  // ---- epilogue: warp w reads TMEM lanes 32*(w%4).., column blocks (w/4), (w/4)+2, ...
  const int lq = warp & 3;
  const long long row = m0 + lq * 32 + lane;
  const bool row_ok = row < a.M;
  const bool plain = !a.out_mask && !a.mask_bits && !a.row_mask && !(a.flags & B3D_FLAG_ACCUMULATE);
  const bool rz = row_ok && a.row_mask && a.row_mask[row] == 0;
  for (int col0 = (warp >> 2) * 32; col0 < Nb; col0 += 64) {
    uint32_t r[32];
    tmem_ld32(tmem + ((uint32_t)(lq * 32) << 16) + (uint32_t)col0, r);
    tmem_ld_wait();
    if (!row_ok) continue;
    epilogue_block32<ACT>(a, row, n0, col0, r, s_bias, plain, rz, a.Nout);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, ncols);
}

// ------------------------------------------------------------------ weight gradient
// D[n (<=128 per CTA), k (<=256 per CTA)] += sum_r dY[r,n] * Acat[r,k]; column Ktot of Acat is a
// virtual all-ones column so that D[:, Ktot] = db. Operands staged MN-major:
//   offset(mn, r) = (mn/8)*1024 + (r/8)*128 + (r%8)*16 + (mn%8)*2   (64-row chunk; SBO=1024, LBO=128)
struct WgTcArgs {
  SegDev dy;
  SegDev seg[B3D_MAX_SEGS];
  int nseg, Ktot, Kp, Nout, ktiles;   // Kp = round_up(Ktot+1, 16)
  long long M, rows_per_split;
  float* part;                         // [S][Nout][Ktot+1]
};

// Stage one 8-element piece (row r, 8 consecutive columns at `p`) of an fp32/bf16 source.
// lo_off != 0 (split-bf16 mode): the bf16 residual goes to dst + lo_off.
__device__ __forceinline__ void stage_piece(uint32_t dst, const void* p, int dtype, uint32_t lo_off = 0) {
  using namespace tc;
  if (dtype == B3D_BF16) {
    cp_async16(dst, p);
    if (lo_off) st_shared_v4(dst + lo_off, 0u, 0u, 0u, 0u);
  } else {
    const float4 v0 = __ldg(reinterpret_cast<const float4*>(p));
    const float4 v1 = __ldg(reinterpret_cast<const float4*>(p) + 1);
    if (lo_off) {
      uint32_t h[4], l[4];
      split_pack2(v0.x, v0.y, h[0], l[0]); split_pack2(v0.z, v0.w, h[1], l[1]);
      split_pack2(v1.x, v1.y, h[2], l[2]); split_pack2(v1.z, v1.w, h[3], l[3]);
      st_shared_v4(dst, h[0], h[1], h[2], h[3]);
      st_shared_v4(dst + lo_off, l[0], l[1], l[2], l[3]);
    } else {
      st_shared_v4(dst, pack_bf16x2(v0.x, v0.y), pack_bf16x2(v0.z, v0.w), pack_bf16x2(v1.x, v1.y), pack_bf16x2(v1.z, v1.w));
    }
  }
}

// The register half of stage_piece for fp32 sources: the stores (asm volatile, memory clobber) keep the compiler from
// moving a later piece's loads above them, so a loop of stage_piece calls pays one global-load latency PER PIECE
// (12 per 64-row chunk: ~14 us per chunk, 1 TB/s). The staging loops below load all pieces of a batch first.
__device__ __forceinline__ void store_piece_f32(uint32_t dst, const float4& v0, const float4& v1, uint32_t lo_off) {
  using namespace tc;
  if (lo_off) {
    uint32_t h[4], l[4];
    split_pack2(v0.x, v0.y, h[0], l[0]); split_pack2(v0.z, v0.w, h[1], l[1]);
    split_pack2(v1.x, v1.y, h[2], l[2]); split_pack2(v1.z, v1.w, h[3], l[3]);
    st_shared_v4(dst, h[0], h[1], h[2], h[3]);
    st_shared_v4(dst + lo_off, l[0], l[1], l[2], l[3]);
  } else {
    st_shared_v4(dst, pack_bf16x2(v0.x, v0.y), pack_bf16x2(v0.z, v0.w), pack_bf16x2(v1.x, v1.y), pack_bf16x2(v1.z, v1.w));
  }
}

template <bool SPLIT>
__global__ void __launch_bounds__(TC_THREADS, SPLIT ? 1 : 2) k_wgrad_tc(const WgTcArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  using namespace tc;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int split = blockIdx.x;
  const int nt = blockIdx.y / a.ktiles, kt = blockIdx.y % a.ktiles;
  const int n0 = nt * TC_BM, k0 = kt * TC_NMAX;
  const int Nk = min(TC_NMAX, a.Kp - k0);
  constexpr int NI = SPLIT ? 2 : 1;      // SPLIT: hi image followed by the bf16 residual image in every stage
  const uint32_t b_stage = (uint32_t)Nk * 128;
  const uint32_t base_off = NI * (2 * TC_A_STAGE + 2 * b_stage);
  const uint32_t sA = smem_u32(smem);
  const uint32_t sB = sA + NI * 2 * TC_A_STAGE;
  const uint32_t a_lo = SPLIT ? TC_A_STAGE : 0u, b_lo = SPLIT ? b_stage : 0u;
  const uint32_t sBar = sA + base_off;
  volatile uint32_t* s_tmem = reinterpret_cast<volatile uint32_t*>(smem + base_off + 24);
  int32_t* s_tab = reinterpret_cast<int32_t*>(smem + base_off + 64);   // [TC_TAB]
  const uint32_t ncols = tmem_cols_for(Nk);
  if (warp == 0) tmem_alloc(sBar + 24, ncols);
  if (tid == 0) {
    mbar_init(sBar + 0, 1);
    mbar_init(sBar + 8, 1);
    mbar_init(sBar + 16, 1);
    fence_mbar_init();
  }
  build_group_table(a.seg, a.nseg, a.Kp / 8 + 1, s_tab, tid, TC_THREADS);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *s_tmem;

  const long long r0 = (long long)split * a.rows_per_split;
  const long long r1 = min(a.M, r0 + a.rows_per_split);
  const int nchunks = (int)((r1 - r0 + TC_BK - 1) / TC_BK);
  const uint32_t idesc = make_idesc_bf16(TC_BM, (uint32_t)Nk, 1, 1);
  const int rl = tid & 7;        // row within an 8-row group
  const int gq = tid >> 3;       // 0..31

  for (int c = 0; c < nchunks; ++c) {
    const int s = c & 1;
    const long long rbase = r0 + (long long)c * TC_BK;
    if (c >= 2) mbar_wait(sBar + 8 * s, ((c >> 1) - 1) & 1);
    // ---- A' = dY^T chunk: 16 n-groups x 64 rows; thread: n-group gq&15, r-groups (gq>>4)*4 .. +3
    {
      const int ng = gq & 15;
      const int n = n0 + ng * 8;
      const bool f32 = a.dy.dtype != B3D_BF16;
      float4 va[4][2];
      int kind[4];      // 0: zeros, 1: fp32 piece in registers, 2: done (bf16 cp.async / ragged tail)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int i = (gq >> 4) * 4 + j;
        const long long r = rbase + i * 8 + rl;
        const uint32_t dst = sA + s * NI * TC_A_STAGE + ng * 1024 + i * 128 + rl * 16;
        kind[j] = 0;
        if (r < r1 && n + 7 < a.Nout) {
          if (f32) {
            const float4* p = reinterpret_cast<const float4*>(seg_addr(a.dy, r, n));
            va[j][0] = __ldg(p);
            va[j][1] = __ldg(p + 1);
            kind[j] = 1;
          } else {
            stage_piece(dst, seg_addr(a.dy, r, n), a.dy.dtype, a_lo);
            kind[j] = 2;
          }
        } else if (r < r1 && n < a.Nout) {   // ragged tail of Nout (not a multiple of 8): scalar
          float t[8];
#pragma unroll
          for (int q = 0; q < 8; ++q)
            t[q] = (n + q < a.Nout)
                       ? (a.dy.dtype == B3D_BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(a.dy.ptr)[r * a.dy.ld + n + q])
                                                 : a.dy.ptr[r * a.dy.ld + n + q])
                       : 0.f;
          va[j][0] = make_float4(t[0], t[1], t[2], t[3]);
          va[j][1] = make_float4(t[4], t[5], t[6], t[7]);
          kind[j] = 1;
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int i = (gq >> 4) * 4 + j;
        const uint32_t dst = sA + s * NI * TC_A_STAGE + ng * 1024 + i * 128 + rl * 16;
        if (kind[j] == 1) {
          store_piece_f32(dst, va[j][0], va[j][1], a_lo);
        } else if (kind[j] == 0) {
          st_shared_v4(dst, 0u, 0u, 0u, 0u);
          if (SPLIT) st_shared_v4(dst + a_lo, 0u, 0u, 0u, 0u);
        }
      }
    }
    // ---- B' = Acat^T chunk: Nk/8 k-groups x 64 rows; thread: k-group gq, all 8 r-groups
    if (gq < Nk / 8) {
      const int ent = s_tab[k0 / 8 + gq];
      const bool ones = (ent < 0) && (k0 + gq * 8 == a.Ktot);
      const bool f32 = ent >= 0 && a.seg[ent >> 24].dtype != B3D_BF16;
      float4 vb[8][2];
      long long gr[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {       // row indices of a gathered segment first (their own latency)
        const long long r = rbase + i * 8 + rl;
        gr[i] = r;
        if (ent >= 0 && r < r1 && a.seg[ent >> 24].idx) gr[i] = (long long)__ldg(a.seg[ent >> 24].idx + r);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const long long r = rbase + i * 8 + rl;
        const uint32_t dst = sB + s * NI * b_stage + gq * 1024 + i * 128 + rl * 16;
        if (ent >= 0 && r < r1) {
          const SegDev& S = a.seg[ent >> 24];
          if (f32) {
            const float4* p = reinterpret_cast<const float4*>(seg_addr(S, gr[i], ent & 0xFFFFFF));
            vb[i][0] = __ldg(p);
            vb[i][1] = __ldg(p + 1);
          } else {
            stage_piece(dst, seg_addr(S, gr[i], ent & 0xFFFFFF), S.dtype, b_lo);
          }
        } else {
          st_shared_v4(dst, (ones && r < r1) ? 0x00003F80u : 0u, 0u, 0u, 0u);   // bf16(1.0) in element 0
          if (SPLIT) st_shared_v4(dst + b_lo, 0u, 0u, 0u, 0u);
        }
      }
      if (f32) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const long long r = rbase + i * 8 + rl;
          const uint32_t dst = sB + s * NI * b_stage + gq * 1024 + i * 128 + rl * 16;
          if (r < r1) store_piece_f32(dst, vb[i][0], vb[i][1], b_lo);
        }
      }
    }
    cp_async_wait_all();
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after_sync();
#pragma unroll
      for (int j = 0; j < TC_BK / 16; ++j) {
        const uint64_t ad = make_smem_desc(sA + s * NI * TC_A_STAGE + j * 256, 128, 1024);   // LBO: next 8-row K group
        const uint64_t bd = make_smem_desc(sB + s * NI * b_stage + j * 256, 128, 1024);      // SBO: next 8-wide MN group
        mma_bf16_ss(tmem, ad, bd, idesc, (c | j) != 0);
        if (SPLIT) {
          mma_bf16_ss(tmem, make_smem_desc(sA + s * NI * TC_A_STAGE + a_lo + j * 256, 128, 1024), bd, idesc, 1);
          mma_bf16_ss(tmem, ad, make_smem_desc(sB + s * NI * b_stage + b_lo + j * 256, 128, 1024), idesc, 1);
        }
      }
      mma_commit(sBar + 8 * s);
      if (c == nchunks - 1) mma_commit(sBar + 16);
    }
  }
  const int Kw = a.Ktot + 1;
  if (nchunks > 0) {
    mbar_wait(sBar + 16, 0);
    tc_fence_after_sync();
  }
  const int lq = warp & 3;
  const int n = n0 + lq * 32 + lane;   // TMEM lane == output row n
  for (int col0 = (warp >> 2) * 32; col0 < Nk; col0 += 64) {
    uint32_t r[32];
    if (nchunks > 0) {
      tmem_ld32(tmem + ((uint32_t)(lq * 32) << 16) + (uint32_t)col0, r);
      tmem_ld_wait();
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) r[j] = 0u;
    }
    if (n < a.Nout) {
      float* prow = a.part + ((long long)split * a.Nout + n) * Kw;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int k = k0 + col0 + j;
        if (k < Kw) prow[k] = __uint_as_float(r[j]);
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, ncols);
}

// Fixed-order sum of the S split partials. A block owns 32 consecutive output elements; its 8 warps take the partials
// p = w, w + 8, ... (each load a coalesced 128-byte line), the eight sums meet in shared memory and are added in warp
// order: deterministic, and S / 8 dependent loads per thread instead of S (the launches were latency-bound: a
// [128, 201] gradient with S = 125 is 100 blocks of threads walking 125 strided loads each).
constexpr int WG_RED_WARPS = 8;
__global__ void __launch_bounds__(32 * WG_RED_WARPS)
k_wgrad_tc_reduce(const float* __restrict__ part, int S, int Nout, int Ktot,
                  float* __restrict__ dW, int lddw, float* __restrict__ db, int accumulate) {
  __shared__ float sm[WG_RED_WARPS][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long i = (long long)blockIdx.x * 32 + lane;
  const int Kw = Ktot + 1;
  const long long tot = (long long)Nout * Kw;
  float s = 0.f;
  if (i < tot)
    for (int p = w; p < S; p += WG_RED_WARPS) s += __ldg(part + (long long)p * tot + i);
  sm[w][lane] = s;
  __syncthreads();
  if (w != 0 || i >= tot) return;
#pragma unroll
  for (int q = 1; q < WG_RED_WARPS; ++q) s += sm[q][lane];
  const int n = (int)(i / Kw), k = (int)(i % Kw);
  if (k < Ktot) {
    float* o = dW + (long long)n * lddw + k;
    *o = accumulate ? *o + s : s;
  } else if (db) {
    db[n] = accumulate ? db[n] + s : s;
  }
}

static void wgrad_tc_plan(long long M, int Nout, int Ktot, int* S, long long* rps, int* ktiles, int* Kp) {
  *Kp = round_up(Ktot + 1, 16);
  *ktiles = (*Kp + TC_NMAX - 1) / TC_NMAX;
  long long tiles = (long long)((Nout + TC_BM - 1) / TC_BM) * *ktiles;
  long long s = (2 * 148 + tiles - 1) / tiles;
  long long smax = (M + 127) / 128;
  if (s > smax) s = smax;
  if (s < 1) s = 1;
  long long r = ((M + s - 1) / s + TC_BK - 1) / TC_BK * TC_BK;
  *S = (int)((M + r - 1) / r);
  *rps = r;
}

static bool seg_tc_ok(const SegDev& S) {
  if (S.mask_mode != B3D_MASK_NONE || (S.width & 7)) return false;
  if (reinterpret_cast<uintptr_t>(S.ptr) & 15) return false;
  return S.dtype == B3D_BF16 ? (S.ld & 7) == 0 : (S.dtype == B3D_F32 && (S.ld & 3) == 0);
}

// ------------------------------------------------------------------ TMA-fed persistent forward / dgrad
// Dense bf16 operands (1-6 row-major segments, each padded to whole 64-column chunks in the packed weights; row-gathered
// segments by TMA tile::gather4 are supported but off in the model path):
//   warp 0 : TMA producer  (cp.async.bulk.tensor 2D boxes {64 cols, 128 rows}, 128B swizzle, 4-8 stage ring)
//   warp 1 : MMA issuer    (tcgen05.mma, operands straight from the TMA-written tiles; identity MMAs for staged addends)
//   warps 2-9: epilogue    (TMEM -> registers -> [bias / un-staged addends] / act / masks -> global; lean or generic block)
//   warps 10..: addend producers (cp.async of whole addend rows into operand-shaped tiles)
// The weight block [Nb, K] stays resident in shared memory for the CTA's lifetime; CTAs are
// persistent over row tiles; two TMEM accumulators let the epilogue of tile t overlap the MMAs of
// tile t+1. No thread touches the operands: the staging cost of k_linear_tc disappears.
constexpr int TMA_STAGES = 4;
constexpr uint32_t TMA_ID_BYTES = 512;
constexpr int TMA_BAR_BYTES = 208;     // mbarriers + TMEM pointer of k_linear_tma (the bias follows)
// warp 0 producer, warp 1 MMA, warps 2-9 epilogue (2 per TMEM lane quarter), warps 10.. addend rows. Two builds of the
// kernel: the general one (2 addend warps, 384 threads x 168 registers: generic epilogue with its look-ahead operand
// registers) and the LEAN one for layers the host knows to need only the lean epilogue block (bf16 output, addends all
// staged, sign-bit masks, no ragged column block): 4 addend warps, 448 threads x 128 registers (a scheduler's 16 K
// registers hold 4 warps of 128).
constexpr int TMA_FIRST_ADD_WARP = 10;
template <bool LEAN> struct TmaCfg {
  static constexpr int ADD_WARPS = LEAN ? 4 : 2;
  static constexpr int THREADS = 32 * (TMA_FIRST_ADD_WARP + ADD_WARPS);
};
constexpr int WG_THREADS = 192;
constexpr int WG_MIN_ROWS = 256;       // fewest rows a split of k_wgrad_tma takes (bounds the number of partials)

// TMA tile::gather4: four rows r0..r3 of a 2D tensor (tensor map encoded with box {64 columns, 1 row}), 64 columns
// from column c0, land in four consecutive 128-byte rows at dst with the same address-based 128B swizzle as a
// tile-mode box (verified on B200: scripts/exp/gather4_test.cu); 512 bytes are signalled on the mbarrier.
__device__ __forceinline__ void tma_gather4(uint32_t dst, const void* tmap, int c0, int r0, int r1, int r2, int r3,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar)
      : "memory");
}

constexpr int TMA_MAX_SEGS = 6;
struct TmaMaps {
  CUtensorMap m[TMA_MAX_SEGS];
};
struct TmaArgs {
  int nseg, seg_chunks[TMA_MAX_SEGS], seg_gsel[TMA_MAX_SEGS];   // 64-column chunks per segment; -1 dense, 0/1 = gathered by gidx[.]
  const int32_t* gidx[2];                                       // row indices of the gathered segments
  int nchunks, Nb, acc_stride;
  long long ntiles;
  const float* bias;
  void* Y;
  int ldy, y_bf16;
  long long M;
  int Nout, act, flags;
  const void* out_mask;
  int ldm, mask_bf16;
  const uint8_t* row_mask;
  SegDev add[2];
  int nadd;
  const uint32_t* mask_bits;
  uint32_t* bits_out;
  // Row-gathered bf16 addends staged through shared memory by two producer warps (cp.async of whole addend rows:
  // coalesced 16-byte chunks, no registers held across the latency) instead of one 32-byte request per epilogue
  // lane: bit t of stage_mask = addend t is staged; add_bufs buffers of add_slots tiles of add_tile_bytes each.
  int stage_mask, add_bufs, add_slots, add_tile_bytes, stages;
  int fast_epi;   // lean epilogue block for plain bf16 layers (B3D_FAST_EPI=0 turns it off for A/B runs)
  int add_ca;     // addend row copies through L1 (cp.async.ca) instead of L2 only (.cg)
  int cluster;    // > 1: the column-block CTAs of a row tile form a thread-block cluster and share its operand loads
};

// One staged addend of one tile, this warp's RPW rows: every instruction copies R rows (lane = chunk c of row
// j + sub); g0 / g1 hold the source rows of tile rows lane and lane + 32 of the warp's share. Fully unrolled, so that
// the choice between g0 and g1, the row offsets and the row part of the swizzle are compile-time.
template <int R, int RPW, bool CA>
__device__ __forceinline__ void add_rows(uint32_t dst, const uint8_t* base, uint32_t ldb, uint32_t g0, uint32_t g1, int sub,
                                         int c, bool on) {
  const uint32_t x0 = (uint32_t)((c ^ sub) & 7);
#pragma unroll
  for (int j = 0; j < RPW; j += R) {
    const uint32_t gr = __shfl_sync(0xffffffffu, j < 32 ? g0 : g1, (j & 31) + sub);
    const uint32_t swz = (x0 ^ (uint32_t)(j & 7)) << 4;     // (c ^ (j + sub)) & 7: j is a multiple of R > sub
    // .ca (opt-in): target-sorted edges repeat the same addend row ~60 times in a row (one node's in-edges), so those
    // copies could hit L1 instead of crossing the L2 -> SM fabric again — measured 2-6 % SLOWER than .cg
    if (on) {
      if (CA)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;"
                     ::"r"(dst + (uint32_t)j * 128u + swz), "l"(base + (unsigned long long)gr * ldb)
                     : "memory");
      else
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;"
                     ::"r"(dst + (uint32_t)j * 128u + swz), "l"(base + (unsigned long long)gr * ldb)
                     : "memory");
    }
  }
}

// Ragged last block of a lean-epilogue layer (fewer than 32 valid columns): the generic block, kept out of line (it
// is ~2,500 instructions) and loading its own accumulator columns so that no register array crosses the call.
template <int ACT>
__device__ __noinline__ void ragged_block(const TmaArgs& a, uint32_t taddr, long long row, int n0, int col0,
                                          const float* s_bias, bool plain, int nlim) {
  uint32_t r[32];
  tc::tmem_ld32(taddr, r);
  tc::tmem_ld_wait();
  if (row < a.M) epilogue_block32<ACT>(a, row, n0, col0, r, s_bias, plain, false, nlim, nullptr);
}

// Epilogue warp of k_linear_tma over this CTA's tiles, lean form (see epilogue_fast32). tm = TMEM address of this
// warp's lane quarter.
template <int ACT, bool BIAS, bool MASK, bool BITS, bool RAGGED>
__device__ __forceinline__ void fast_tiles(const TmaArgs& a, uint32_t tm, uint32_t sBar, const float* s_bias, int n0, int nlim,
                                           int cb0, long long lrow, int lane, bool plain) {
  using namespace tc;
  const long long M = a.M, ntiles = a.ntiles;
  const int Nb = a.Nb, acc_stride = a.acc_stride, ldy = a.ldy;
  const uint32_t* mask_bits = a.mask_bits;
  uint32_t* bits_out = a.bits_out;
  // sign-bit mask words of this lane's row, one per 32-column block of this warp (at most 4: Nb <= 256), loaded ONE
  // TILE AHEAD: the words stream from DRAM (each is used once), and a block that waits for its own word pays a full
  // DRAM latency with only two warps per scheduler to hide it
  uint32_t mw[4] = {0u, 0u, 0u, 0u}, mwn[4] = {0u, 0u, 0u, 0u};
  auto load_masks = [&](long long tile, uint32_t (&w)[4]) {
    const long long row = tile * TC_BM + lrow;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int cbase = n0 + cb0 + 64 * b;
      w[b] = (tile < ntiles && row < M && cb0 + 64 * b < Nb && cbase + 31 < nlim)
                 ? __ldg(mask_bits + (long long)(cbase >> 5) * M + row) : 0u;
    }
  };
  if (MASK) load_masks(blockIdx.x, mw);
  int tcount = 0;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tcount) {
    const int acc = tcount & 1;
    const long long row = tile * TC_BM + lrow;
    const bool row_ok = row < M;
    __nv_bfloat16* yrow = reinterpret_cast<__nv_bfloat16*>(a.Y) + row * ldy + n0;
    if (MASK) load_masks(tile + gridDim.x, mwn);
    mbar_wait(sBar + 64 + 8 * acc, (tcount >> 1) & 1);
    tc_fence_after_sync();
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int col0 = cb0 + 64 * b;
      if (col0 >= Nb) break;
      const int cbase = n0 + col0;
      const uint32_t taddr = tm + (uint32_t)(acc * acc_stride + col0);
      if (!RAGGED && cbase >= nlim) break;      // column block past the last output column (N % Nb != 0)
      if (!RAGGED || cbase + 31 < nlim) {       // warp-uniform
        uint32_t r[32];
        tmem_ld32(taddr, r);
        tmem_ld_wait();
        if (row_ok)
          epilogue_fast32<ACT, BIAS, MASK, BITS>(r, s_bias + col0, BITS ? bits_out + (long long)(cbase >> 5) * M + row : nullptr,
                                                 mw[b], yrow + col0);
      } else if (RAGGED) {
        ragged_block<ACT>(a, taddr, row, n0, col0, s_bias, plain, nlim);
      }
    }
    tc_fence_before_sync();
    __syncwarp();
    if (lane == 0) mbar_arrive(sBar + 80 + 8 * acc);
    if (MASK) {
#pragma unroll
      for (int b = 0; b < 4; ++b) mw[b] = mwn[b];
    }
  }
}

template <int ACT, bool LEAN>
__global__ void __launch_bounds__(TmaCfg<LEAN>::THREADS, 1)
k_linear_tma(const __grid_constant__ TmaMaps maps, const __grid_constant__ CUtensorMap mapW,
             const __grid_constant__ TmaArgs a) {
  constexpr int TMA_THREADS = TmaCfg<LEAN>::THREADS, TMA_ADD_WARPS = TmaCfg<LEAN>::ADD_WARPS;
  extern __shared__ __align__(1024) uint8_t smem[];
  using namespace tc;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n0 = blockIdx.y * a.Nb;
  const uint32_t w_chunk = (uint32_t)a.Nb * 128;
  const uint32_t pad = (1024u - (smem_u32(smem) & 1023u)) & 1023u;   // 128B-swizzled tiles need 1024-byte alignment
  const uint32_t sW = smem_u32(smem) + pad;
  const uint32_t sA = sW + a.nchunks * w_chunk;                 // both multiples of 1024
  const int STG = a.stages;
  const uint32_t add_buf_bytes = (uint32_t)(a.add_slots * a.add_tile_bytes);
  // staged addend tiles: [add_bufs][add_slots][Nb / 64 column groups][128 rows][64 bf16], every group a 128B-swizzled
  // K-major operand tile (what TMA would write), followed by a 16 x 16 identity (the B operand of the addend MMAs)
  const uint32_t sAdd = sA + STG * TC_A_STAGE;
  const uint32_t sId = sAdd + a.add_bufs * add_buf_bytes;
  const uint32_t id_bytes = a.stage_mask ? TMA_ID_BYTES : 0u;
  const uint32_t misc = a.nchunks * w_chunk + STG * TC_A_STAGE + a.add_bufs * add_buf_bytes + id_bytes;
  // full[0..3] @0, empty[0..3] @32, accf[2] @64, acce[2] @80, wfull @96, tmem ptr @104, addf[2] @112, adde[2] @128,
  // full[4..7] @144, empty[4..7] @176 (operand rings of up to 8 stages), bias @TMA_BAR_BYTES
  const uint32_t sBar = sW + misc;
  volatile uint32_t* s_tmem = reinterpret_cast<volatile uint32_t*>(smem + pad + misc + 104);
  float* s_bias = reinterpret_cast<float*>(smem + pad + misc + TMA_BAR_BYTES);  // [256]
  auto bar_full = [&](int s) { return sBar + (uint32_t)(s < 4 ? 8 * s : 144 + 8 * (s - 4)); };
  auto bar_empty = [&](int s) { return sBar + (uint32_t)(s < 4 ? 32 + 8 * s : 176 + 8 * (s - 4)); };
  const uint32_t ncols = 2 * (uint32_t)a.acc_stride;

  if (warp == 1) tmem_alloc(sBar + 104, ncols);
  if (tid == 0) {
    // cluster mode: a stage is free once the MMAs of EVERY CTA of the cluster have read it (multicast commits)
    for (int s = 0; s < 8; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_empty(s), (uint32_t)a.cluster); }
    mbar_init(sBar + 64, 1); mbar_init(sBar + 72, 1);      // accumulator full (tcgen05.commit)
    mbar_init(sBar + 80, 8); mbar_init(sBar + 88, 8);      // accumulator empty (8 epilogue warps)
    mbar_init(sBar + 96, 1);                                // weights resident
    mbar_init(sBar + 112, 32 * TMA_ADD_WARPS); mbar_init(sBar + 120, 32 * TMA_ADD_WARPS);  // staged addends landed (all producer threads)
    mbar_init(sBar + 128, 1); mbar_init(sBar + 136, 1);    // staged addends consumed (tcgen05.commit)
    fence_mbar_init();
  }
  if (a.stage_mask) {
    // B operand of the addend MMAs: a 16 x 16 bf16 identity (row n = output column n of a 16-column group, K index k),
    // K-major without swizzle: offset(n, k) = (k / 8) * 256 + n * 16 + (k % 8) * 2  (LBO 256, SBO 128; 512 bytes).
    for (int wi = tid; wi < (int)TMA_ID_BYTES / 4; wi += TMA_THREADS) {
      const int n = (wi & 63) >> 2, k0 = (wi >> 6) * 8 + (wi & 3) * 2;
      const uint32_t v = ((k0 == n) ? 0x3F80u : 0u) | ((k0 + 1 == n) ? 0x3F800000u : 0u);
      asm volatile("st.shared.b32 [%0], %1;" ::"r"(sId + 4u * (uint32_t)wi), "r"(v) : "memory");
    }
    fence_proxy_async_smem();
  }
  for (int i = tid; i < 256; i += TMA_THREADS) s_bias[i] = (a.bias && i < a.Nb && n0 + i < a.Nout) ? __ldg(a.bias + n0 + i) : 0.f;
  tc_fence_before_sync();
  if (a.cluster > 1) cluster_sync_all();      // peers' barriers are initialised before any remote arrival / multicast
  else __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *s_tmem;
  const uint16_t cmask = (uint16_t)((1u << a.cluster) - 1u);
  const int crank = a.cluster > 1 ? (int)cluster_ctarank() : 0;

  if (warp == 0) {
    // Producer warp. Dense segments: lane 0 issues one {64 x 128} box per chunk. Cluster mode (the column-block CTAs
    // of a row tile, which all read the same [128, K] operand): every CTA arms its own barrier for every chunk, the
    // chunks are issued round-robin by the CTAs of the cluster and MULTICAST into all of them — the operand crosses the
    // L2 -> SM fabric once per row tile instead of once per column block. Row-gathered segments (the
    // reference's x[edge_index] operands, pose_gnn.py:180 / clr_att_gnn.py:284): every lane owns 4 rows of the tile
    // and issues one tile::gather4 per chunk for them, so the gather runs in the TMA unit (no registers, no L1
    // wavefronts, hundreds of rows in flight) instead of in the epilogue.
    if (lane == 0) {
      mbar_expect_tx(sBar + 96, a.nchunks * w_chunk);
      for (int c = 0; c < a.nchunks; ++c) tma_load_2d(sW + c * w_chunk, &mapW, c * 64, n0, sBar + 96);
    }
    int it = 0;
    for (long long tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
      int g[2][4];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const long long r = tile * TC_BM + 4 * lane + j;
          g[q][j] = (a.gidx[q] && r < a.M) ? __ldg(a.gidx[q] + r) : 0;
        }
      }
      for (int sg = 0; sg < a.nseg; ++sg) {
        const int sel = a.seg_gsel[sg];
        for (int c = 0; c < a.seg_chunks[sg]; ++c, ++it) {
          const int s = it % STG;
          if (lane == 0) {
            if (it >= STG) mbar_wait(bar_empty(s), ((it / STG) - 1) & 1);
            mbar_expect_tx(bar_full(s), TC_A_STAGE);
            if (sel < 0) {
              if (a.cluster <= 1) tma_load_2d(sA + s * TC_A_STAGE, &maps.m[sg], c * 64, (int)(tile * TC_BM), bar_full(s));
              else if (it % a.cluster == crank)
                tma_load_2d_mc(sA + s * TC_A_STAGE, &maps.m[sg], c * 64, (int)(tile * TC_BM), bar_full(s), cmask);
            }
          }
          if (sel >= 0) {
            __syncwarp();      // the stage is free and its transaction count armed before any lane writes into it
            tma_gather4(sA + s * TC_A_STAGE + lane * 512, &maps.m[sg], c * 64, g[sel][0], g[sel][1], g[sel][2], g[sel][3],
                        bar_full(s));
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(TC_BM, (uint32_t)a.Nb, 0, 0);
      const uint32_t idesc16 = make_idesc_bf16(TC_BM, 16u, 0, 0);
      const uint64_t id_desc = make_smem_desc(sId, 256, 128);
      mbar_wait(sBar + 96, 0);
      int it = 0, tcount = 0;
      for (long long tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++tcount) {
        const int acc = tcount & 1;
        if (tcount >= 2) mbar_wait(sBar + 80 + 8 * acc, ((tcount >> 1) - 1) & 1);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem + (uint32_t)(acc * a.acc_stride);
        for (int c = 0; c < a.nchunks; ++c, ++it) {
          const int s = it % STG;
          mbar_wait(bar_full(s), (it / STG) & 1);
          tc_fence_after_sync();
#pragma unroll
          for (int j = 0; j < TC_BK / 16; ++j)
            mma_bf16_ss(d_tmem, make_smem_desc_sw128(sA + s * TC_A_STAGE + j * 32),
                        make_smem_desc_sw128(sW + c * w_chunk + j * 32), idesc, (c | j) != 0);
          if (a.cluster > 1) mma_commit_mc(bar_empty(s), cmask);
          else mma_commit(bar_empty(s));      // stage free once these MMAs have read it
        }
        if (a.stage_mask) {
          // Row-gathered addends: D[:, 16-column group] += staged tile[:, same group] * I. The rows were gathered by
          // the producer warps 10-11 into operand-shaped tiles, so the tensor core (idle > 85 % of the time in these
          // layers) does the additions and the epilogue never sees an addend: bf16 * 1.0 accumulated in fp32 is exact.
          const int buf = tcount % a.add_bufs;
          mbar_wait(sBar + 112 + 8 * buf, (tcount / a.add_bufs) & 1);
          fence_proxy_async_smem();
          tc_fence_after_sync();
          for (int slot = 0; slot < a.add_slots; ++slot) {
            const uint32_t base = sAdd + buf * add_buf_bytes + slot * a.add_tile_bytes;
            for (int q = 0; q < (a.Nb >> 6); ++q) {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                mma_bf16_ss(d_tmem + (uint32_t)(q * 64 + j * 16), make_smem_desc_sw128(base + q * TC_A_STAGE + j * 32),
                            id_desc, idesc16, 1u);
            }
          }
          mma_commit(sBar + 128 + 8 * buf);   // addend buffer free once these MMAs have read it
        }
        mma_commit(sBar + 64 + 8 * acc);      // accumulator ready for the epilogue
      }
    }
  } else if (warp >= TMA_FIRST_ADD_WARP) {
    // Addend producer warps: a warp instruction copies whole addend rows — R = 32 / chunks rows per instruction, the
    // lanes of a row = its consecutive 16-byte chunks (coalesced 256- or 512-byte requests, 2-4 L1 wavefronts per row,
    // no registers held across the latency) — into tiles shaped like the TMA-written operand tiles: column group
    // q = chunk / 8 of tile row rr at q * 16 KB + rr * 128, its chunk c at position (c ^ rr) & 7. The MMA warp consumes
    // them. Completion is signalled by the copies themselves (cp.async.mbarrier.arrive.noinc), so the producers run as
    // far ahead as the buffers allow. Rows past M copy source row 0 (never stored): the loop has no predicates, and
    // the row indices of the next tile are loaded while this tile's copies are issued.
    if (a.stage_mask) {
      const int wp = warp - TMA_FIRST_ADD_WARP;
      constexpr int RPW = TC_BM / TMA_ADD_WARPS;          // tile rows per producer warp (<= 64: two index registers)
      const int chunks = (a.Nb * 2) >> 4;                 // 16-byte chunks per staged row (Nb % 64 == 0, <= 32)
      const int R = chunks <= 8 ? 4 : chunks <= 16 ? 2 : 1;   // rows per warp instruction
      const int sub = lane / (32 / R);                    // this lane's row within the instruction
      const int c = lane % (32 / R);                      // its chunk
      const bool on = c < chunks;                         // (a 24-chunk row leaves 8 lanes idle)
      const uint32_t c_off = (uint32_t)(c >> 3) * TC_A_STAGE;
      auto load_idx = [&](long long tile, int t, uint32_t& g0, uint32_t& g1) {
        const SegDev& S = a.add[t];
        const long long r0 = tile * TC_BM + RPW * wp + lane, r1 = r0 + 32;
        g0 = (tile < a.ntiles && r0 < a.M && lane < RPW) ? (S.idx ? (uint32_t)__ldg(S.idx + r0) : (uint32_t)r0) : 0u;
        g1 = (tile < a.ntiles && r1 < a.M && lane + 32 < RPW) ? (S.idx ? (uint32_t)__ldg(S.idx + r1) : (uint32_t)r1) : 0u;
      };
      uint32_t g[2][2], gn[2][2];
      for (int t = 0; t < 2; ++t) {
        g[t][0] = g[t][1] = 0u;
        if (t < a.nadd && ((a.stage_mask >> t) & 1)) load_idx(blockIdx.x, t, g[t][0], g[t][1]);
      }
      int tcount = 0;
      for (long long tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++tcount) {
        const int buf = tcount % a.add_bufs;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          gn[t][0] = gn[t][1] = 0u;
          if (t < a.nadd && ((a.stage_mask >> t) & 1)) load_idx(tile + gridDim.x, t, gn[t][0], gn[t][1]);
        }
        if (tcount >= a.add_bufs) {
          if (lane == 0) mbar_wait(sBar + 128 + 8 * buf, ((tcount / a.add_bufs) - 1) & 1);
          __syncwarp();
        }
        int slot = 0;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          if (!(t < a.nadd && ((a.stage_mask >> t) & 1))) continue;
          const SegDev& S = a.add[t];
          const uint32_t dst_t = sAdd + buf * add_buf_bytes + slot * a.add_tile_bytes + c_off + (uint32_t)(RPW * wp + sub) * 128u;
          const uint8_t* base = reinterpret_cast<const uint8_t*>(reinterpret_cast<const __nv_bfloat16*>(S.ptr) + n0) + c * 16;
          const uint32_t ldb = (uint32_t)S.ld * 2u;
          const uint32_t g0 = g[t][0], g1 = g[t][1];
          if (a.add_ca) {
            if (R == 1) add_rows<1, RPW, true>(dst_t, base, ldb, g0, g1, sub, c, on);
            else if (R == 2) add_rows<2, RPW, true>(dst_t, base, ldb, g0, g1, sub, c, on);
            else add_rows<4, RPW, true>(dst_t, base, ldb, g0, g1, sub, c, on);
          } else {
            if (R == 1) add_rows<1, RPW, false>(dst_t, base, ldb, g0, g1, sub, c, on);
            else if (R == 2) add_rows<2, RPW, false>(dst_t, base, ldb, g0, g1, sub, c, on);
            else add_rows<4, RPW, false>(dst_t, base, ldb, g0, g1, sub, c, on);
          }
          ++slot;
        }
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(sBar + 112 + 8 * buf) : "memory");
#pragma unroll
        for (int t = 0; t < 2; ++t) { g[t][0] = gn[t][0]; g[t][1] = gn[t][1]; }
      }
    }
  } else {
    const int lq = warp & 3;                  // TMEM lane quarter this warp may access
    const bool plain = !a.out_mask && !a.mask_bits && !a.row_mask && !(a.flags & B3D_FLAG_ACCUMULATE);
    const int cb0 = ((warp - 2) >> 2) * 32;   // the two warps of a quarter interleave 32-column blocks
    const int nlim = min(a.Nout, n0 + a.Nb);
    const long long lrow = lq * 32 + lane;
    // The global operands of a block (bf16 ReLU mask, row-gathered addends) are requested ONE BLOCK
    // AHEAD — the next block of this tile, or the first block of this CTA's next tile — so their
    // latency overlaps the TMEM load, the arithmetic and the stores of the current block. Two register
    // sets alternate roles (no moves: a move would wait for the load it copies).
    auto gather_rows = [&](long long tile, int& g0, int& g1) {
      const long long row = tile * TC_BM + lrow;
      g0 = g1 = 0;
      if (tile < a.ntiles && row < a.M) {
        g0 = (a.nadd > 0 && a.add[0].idx) ? __ldg(a.add[0].idx + row) : (int)row;
        g1 = (a.nadd > 1 && a.add[1].idx) ? __ldg(a.add[1].idx + row) : (int)row;
      }
    };
    EpiPrefetch pfa, pfb;
    pfa.flags = pfb.flags = 0;
    // addend source rows of this tile / this CTA's next tile / the tile after that: the index loads run
    // TWO tiles ahead, because the look-ahead prefetch selects between this tile's and the next tile's rows
    // and a select waits for both operands (ncu: 20 % of the stall samples sat on that address computation)
    int cg0, cg1, ng0, ng1, fg0, fg1;
    gather_rows(blockIdx.x, cg0, cg1);
    // Measured (scripts/epi_probe.py): the look-ahead pays for the gathered addends (L2-resident node
    // tables: 437 -> 380 us on the 64->192 message layer) but not for the dense DRAM-streamed ReLU mask
    // (647 -> 700 us), which keeps the same-block prefetch.
    const bool all_staged = a.nadd > 0 && a.stage_mask == (1 << a.nadd) - 1;
    const bool fast = LEAN || (a.fast_epi && ACT != B3D_ACT_SIGMOID && a.y_bf16 && !a.out_mask && !a.row_mask &&
                      !(a.flags & B3D_FLAG_ACCUMULATE) && (a.nadd == 0 || all_staged) && (a.ldy & 15) == 0 && al32(a.Y) &&
                      (n0 & 31) == 0);
    if (fast) {
      // bf16 layer whose addends (if any) are already in the accumulator: the lean block (epilogue_fast32),
      // instantiated per (bias, sign-bit mask, sign-bit output) combination so that no per-block branch is left
      const int sel = (a.bias ? 1 : 0) | (a.mask_bits ? 2 : 0) | (a.bits_out ? 4 : 0);
      const uint32_t tm = tmem + ((uint32_t)(lq * 32) << 16);
      switch (sel) {
        case 0: fast_tiles<ACT, false, false, false, !LEAN>(a, tm, sBar, s_bias, n0, nlim, cb0, lrow, lane, plain); break;
        case 1: fast_tiles<ACT, true, false, false, !LEAN>(a, tm, sBar, s_bias, n0, nlim, cb0, lrow, lane, plain); break;
        case 2: fast_tiles<ACT, false, true, false, !LEAN>(a, tm, sBar, s_bias, n0, nlim, cb0, lrow, lane, plain); break;
        case 3: fast_tiles<ACT, true, true, false, !LEAN>(a, tm, sBar, s_bias, n0, nlim, cb0, lrow, lane, plain); break;
        case 4: fast_tiles<ACT, false, false, true, !LEAN>(a, tm, sBar, s_bias, n0, nlim, cb0, lrow, lane, plain); break;
        case 5: fast_tiles<ACT, true, false, true, !LEAN>(a, tm, sBar, s_bias, n0, nlim, cb0, lrow, lane, plain); break;
        case 6: fast_tiles<ACT, false, true, true, !LEAN>(a, tm, sBar, s_bias, n0, nlim, cb0, lrow, lane, plain); break;
        default: fast_tiles<ACT, true, true, true, !LEAN>(a, tm, sBar, s_bias, n0, nlim, cb0, lrow, lane, plain); break;
      }
    } else if (LEAN) {
    } else if (a.out_mask || a.nadd == 0 || all_staged) {
      int tcount = 0;
      for (long long tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++tcount) {
        const int acc = tcount & 1;
        const long long row = tile * TC_BM + lrow;
        const bool row_ok = row < a.M;
        const bool rz = row_ok && a.row_mask && a.row_mask[row] == 0;
        if (tcount) gather_rows(tile, cg0, cg1);
        mbar_wait(sBar + 64 + 8 * acc, (tcount >> 1) & 1);
        tc_fence_after_sync();
        for (int col0 = cb0; col0 < a.Nb; col0 += 64) {
          uint32_t r[32];
          tmem_ld32(tmem + ((uint32_t)(lq * 32) << 16) + (uint32_t)(acc * a.acc_stride + col0), r);
          pfa.flags = 0;
          if (row_ok) epilogue_prefetch(a, row, cg0, cg1, n0 + col0, nlim, pfa);   // global latency overlaps the TMEM load
          tmem_ld_wait();
          if (row_ok) epilogue_block32<ACT>(a, row, n0, col0, r, s_bias, plain, rz, nlim, &pfa);
        }
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(sBar + 80 + 8 * acc);
      }
    } else {
    if ((long long)blockIdx.x < a.ntiles && blockIdx.x * (long long)TC_BM + lrow < a.M && cb0 < a.Nb)
      epilogue_prefetch(a, blockIdx.x * (long long)TC_BM + lrow, cg0, cg1, n0 + cb0, nlim, pfa);
    gather_rows((long long)blockIdx.x + gridDim.x, ng0, ng1);
    int tcount = 0, parity = 0;
    for (long long tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++tcount) {
      const int acc = tcount & 1;
      const long long ntile = tile + gridDim.x;
      gather_rows(ntile + gridDim.x, fg0, fg1);
      const long long row = tile * TC_BM + lrow, nrow = ntile * TC_BM + lrow;
      const bool row_ok = row < a.M;
      const bool rz = row_ok && a.row_mask && a.row_mask[row] == 0;
      mbar_wait(sBar + 64 + 8 * acc, (tcount >> 1) & 1);
      tc_fence_after_sync();
      for (int col0 = cb0; col0 < a.Nb; col0 += 64, parity ^= 1) {
        uint32_t r[32];
        tmem_ld32(tmem + ((uint32_t)(lq * 32) << 16) + (uint32_t)(acc * a.acc_stride + col0), r);
        const bool same = col0 + 64 < a.Nb;
        const long long prow = same ? row : nrow;
        const int pcol = n0 + (same ? col0 + 64 : cb0);
        const bool p_ok = same ? row_ok : (ntile < a.ntiles && nrow < a.M);
        if (parity == 0) {
          pfb.flags = 0;
          if (p_ok) epilogue_prefetch(a, prow, same ? cg0 : ng0, same ? cg1 : ng1, pcol, nlim, pfb);
          tmem_ld_wait();
          if (row_ok) epilogue_block32<ACT>(a, row, n0, col0, r, s_bias, plain, rz, nlim, &pfa);
        } else {
          pfa.flags = 0;
          if (p_ok) epilogue_prefetch(a, prow, same ? cg0 : ng0, same ? cg1 : ng1, pcol, nlim, pfa);
          tmem_ld_wait();
          if (row_ok) epilogue_block32<ACT>(a, row, n0, col0, r, s_bias, plain, rz, nlim, &pfb);
        }
      }
      cg0 = ng0; cg1 = ng1; ng0 = fg0; ng1 = fg1;
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(sBar + 80 + 8 * acc);
    }
    }
  }
  tc_fence_before_sync();
  if (a.cluster > 1) cluster_sync_all();      // no CTA leaves while a peer may still write into it or arrive on its barriers
  else __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, ncols);
}

// Row-major bf16 pack with every input segment padded to whole 64-column chunks (gathered / ragged segments of
// b3d_linear_tma each start on a chunk boundary): packed column kp of segment s, offset o < width_s, holds logical
// column (sum of earlier widths) + o; padding columns are zero.
struct SegWidths { int n, w[6]; };
__global__ void k_pack_weights_rm_segs(const float* __restrict__ W, int ldw, int n_log, int transpose, SegWidths sw,
                                       __nv_bfloat16* __restrict__ Wr, int Npad, int Kpad) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)Npad * Kpad) return;
  int n = (int)(i / Kpad), kp = (int)(i % Kpad);
  int k = -1, base = 0, pbase = 0;
  for (int s = 0; s < sw.n; ++s) {
    const int pw = (sw.w[s] + 63) / 64 * 64;
    if (kp < pbase + pw) { if (kp - pbase < sw.w[s]) k = base + kp - pbase; break; }
    base += sw.w[s]; pbase += pw;
  }
  float v = 0.f;
  if (n < n_log && k >= 0) v = transpose ? W[(long long)k * ldw + n] : W[(long long)n * ldw + k];
  Wr[i] = __float2bfloat16_rn(v);
}

// Row-major bf16 pack for the TMA path: Wr[n][k] = bf16(B[n][k]) zero padded to [Npad][Kpad].
__global__ void k_pack_weights_rm(const float* __restrict__ W, int ldw, int n_log, int k_log, int transpose,
                                  __nv_bfloat16* __restrict__ Wr, int Npad, int Kpad) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)Npad * Kpad) return;
  int n = (int)(i / Kpad), k = (int)(i % Kpad);
  float v = 0.f;
  if (n < n_log && k < k_log) v = transpose ? W[(long long)k * ldw + n] : W[(long long)n * ldw + k];
  Wr[i] = __float2bfloat16_rn(v);
}

// ------------------------------------------------------------------ TMA-fed weight gradient
// dW[n,k] = sum_r dY[r,n] A[r,k] for DENSE bf16 dY / A (1-2 segments). Both operands are MN-major
// tiles written by TMA (boxes {64 cols, 64 rows}, 128B swizzle): no thread touches them.
//   warp 0: TMA producer, warp 1: tcgen05 issuer, all 6 warps: final epilogue (fp32 partials).
// Bias gradient: a constant MN-major "ones" tile (column 0 = 1) multiplies dY^T in a second, N=16
// MMA per K step; rows past the end of the split are zero-filled by TMA, so they drop out.
constexpr int WG_STAGES = 4;
constexpr int WG_BOX = 64 * 128;   // bytes of one {64 cols, 64 rows} bf16 box

struct WgTmaArgs {
  int seg0_groups, ngroups_total;   // 64-column groups of A: in segment 0 / in total
  int Ktot, Nout, ktiles;
  long long M, rows_per_split;
  float* part;                       // [S][Nout][Ktot+1]
};

__device__ __forceinline__ uint64_t make_smem_desc_sw128_mn(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;   // next 64-wide MN group
  d |= (uint64_t)(1024 >> 4) << 32;                    // next 8-row K group
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;                              // SWIZZLE_128B
  return d;
}

__global__ void __launch_bounds__(WG_THREADS, 1)
k_wgrad_tma(const __grid_constant__ CUtensorMap mapDY, const __grid_constant__ CUtensorMap mapA0,
            const __grid_constant__ CUtensorMap mapA1, const WgTmaArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  using namespace tc;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int split = blockIdx.x;
  const int nt = blockIdx.y / a.ktiles, kt = blockIdx.y % a.ktiles;
  const int n0 = nt * TC_BM;
  const int g0 = kt * 4;                                       // first 64-column group of this k tile
  const int ng = min(4, a.ngroups_total - g0);                 // groups in this tile (N' = 64*ng)
  const bool with_bias = (kt == 0);
  const uint32_t stage_bytes = 2 * WG_BOX + 4 * WG_BOX;        // dY^T (2 boxes) + A^T (<= 4 boxes) = 48 KB
  const uint32_t pad = (1024u - (smem_u32(smem) & 1023u)) & 1023u;
  const uint32_t sStage = smem_u32(smem) + pad;
  const uint32_t sOnes = sStage + WG_STAGES * stage_bytes;     // 8 KB constant tile
  const uint32_t sBar = sOnes + WG_BOX;                        // full[4] @0, empty[4] @32, done @64, tmem ptr @72
  volatile uint32_t* s_tmem = reinterpret_cast<volatile uint32_t*>(smem + pad + WG_STAGES * stage_bytes + WG_BOX + 72);
  if (warp == 1) tmem_alloc(sBar + 72, 512);
  if (tid == 0) {
    for (int s = 0; s < WG_STAGES; ++s) { mbar_init(sBar + 8 * s, 1); mbar_init(sBar + 32 + 8 * s, 1); }
    mbar_init(sBar + 64, 1);
    fence_mbar_init();
  }
  // ones tile: element (row r, col 0) = 1.0 -> 16-byte chunk (0 ^ (r & 7)) of row r
  for (int i = tid; i < WG_BOX / 16; i += WG_THREADS) {
    const int r = i >> 3, ch = i & 7;
    st_shared_v4(sOnes + i * 16, (ch == (r & 7)) ? 0x00003F80u : 0u, 0u, 0u, 0u);
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *s_tmem;

  const long long r0 = (long long)split * a.rows_per_split;
  const long long r1 = min(a.M, r0 + a.rows_per_split);
  const int nchunks = (int)((r1 - r0 + 63) / 64);

  if (warp == 0 && lane == 0) {
    for (int c = 0; c < nchunks; ++c) {
      const int s = c % WG_STAGES;
      if (c >= WG_STAGES) mbar_wait(sBar + 32 + 8 * s, ((c / WG_STAGES) - 1) & 1);
      const uint32_t base = sStage + s * stage_bytes;
      const int row = (int)(r0 + (long long)c * 64);
      mbar_expect_tx(sBar + 8 * s, (2 + ng) * WG_BOX);
      tma_load_2d(base, &mapDY, n0, row, sBar + 8 * s);
      tma_load_2d(base + WG_BOX, &mapDY, n0 + 64, row, sBar + 8 * s);
      for (int g = 0; g < ng; ++g) {
        const int gg = g0 + g;
        if (gg < a.seg0_groups) tma_load_2d(base + (2 + g) * WG_BOX, &mapA0, gg * 64, row, sBar + 8 * s);
        else tma_load_2d(base + (2 + g) * WG_BOX, &mapA1, (gg - a.seg0_groups) * 64, row, sBar + 8 * s);
      }
    }
  } else if (warp == 1 && lane == 0) {
    const uint32_t idesc = make_idesc_bf16(TC_BM, (uint32_t)(64 * ng), 1, 1);
    const uint32_t idesc1 = make_idesc_bf16(TC_BM, 16u, 1, 1);
    for (int c = 0; c < nchunks; ++c) {
      const int s = c % WG_STAGES;
      mbar_wait(sBar + 8 * s, (c / WG_STAGES) & 1);
      tc_fence_after_sync();
      const uint32_t base = sStage + s * stage_bytes;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint64_t ad = make_smem_desc_sw128_mn(base + j * 2048, WG_BOX);
        mma_bf16_ss(tmem, ad, make_smem_desc_sw128_mn(base + 2 * WG_BOX + j * 2048, WG_BOX), idesc, (c | j) != 0);
        if (with_bias) mma_bf16_ss(tmem + 256, ad, make_smem_desc_sw128_mn(sOnes + j * 2048, WG_BOX), idesc1, (c | j) != 0);
      }
      mma_commit(sBar + 32 + 8 * s);
    }
    mma_commit(sBar + 64);
  }
  __syncwarp();
  // ---- epilogue (all warps): TMEM lane = output row n; warps w and w+4 share a lane quarter
  const int Kw = a.Ktot + 1;
  if (nchunks > 0) {
    mbar_wait(sBar + 64, 0);
    tc_fence_after_sync();
  }
  if (warp < 4 || warp < 6) {
    const int lq = warp & 3;
    const int n = n0 + lq * 32 + lane;
    // warps 0-3 take column blocks 0,2,4,..; warps 4,5 (quarters 0,1) help with blocks 1,3,.. of their quarter;
    // quarters 2,3 have a single warp, which therefore walks every block
    const bool shared_quarter = lq < 2;
    const int first = (warp >= 4) ? 1 : 0;
    const int step = shared_quarter ? 2 : 1;
    const int nblk = (64 * ng) / 32;
    if (!(warp >= 4 && !shared_quarter)) {
      for (int blk = first; blk < nblk; blk += step) {
        uint32_t r[32];
        if (nchunks > 0) { tmem_ld32(tmem + ((uint32_t)(lq * 32) << 16) + (uint32_t)(blk * 32), r); tmem_ld_wait(); }
        else {
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = 0u;
        }
        if (n < a.Nout) {
          float* prow = a.part + ((long long)split * a.Nout + n) * Kw;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int k = g0 * 64 + blk * 32 + j;
            if (k < a.Ktot) prow[k] = __uint_as_float(r[j]);
          }
        }
      }
      if (with_bias && warp < 4) {
        uint32_t r[32];
        if (nchunks > 0) { tmem_ld32(tmem + ((uint32_t)(lq * 32) << 16) + 256u, r); tmem_ld_wait(); }
        else r[0] = 0u;
        if (n < a.Nout) a.part[((long long)split * a.Nout + n) * Kw + a.Ktot] = __uint_as_float(r[0]);
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

}  // namespace b3d

using namespace b3d;

extern "C" size_t b3d_tc_packed_bytes(int32_t n_logical, int32_t k_logical) {
  return (size_t)round_up(n_logical, 16) * round_up(k_logical, TC_BK) * 2;
}

extern "C" int b3d_tc_pack_weights(const float* W, int32_t ldw, int32_t n_logical, int32_t k_logical,
                                   int32_t transpose, void* Wp, void* stream) {
  if (!W || !Wp || n_logical <= 0 || k_logical <= 0) return bad_arg("b3d_tc_pack_weights");
  int Npad = round_up(n_logical, 16), Kpad = round_up(k_logical, TC_BK);
  long long total = (long long)Npad * Kpad;
  if (transpose & 2) {   // B3D_PACK_SPLIT: tf32 hi + lo images for b3d_linear_tc(B3D_FLAG_SPLIT); Wp holds 4x b3d_tc_packed_bytes
    Kpad = round_up(k_logical, 32);
    total = (long long)Npad * Kpad;
    k_pack_weights_tf32<<<(unsigned)ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(
        W, ldw, n_logical, k_logical, transpose & 1, reinterpret_cast<float*>(Wp), Npad, Kpad);
    B3D_LAUNCH_CHECK("k_pack_weights_tf32");
    return 0;
  }
  k_pack_weights<<<(unsigned)ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(
      W, ldw, n_logical, k_logical, transpose & 1, reinterpret_cast<__nv_bfloat16*>(Wp), Npad, Kpad, 0);
  B3D_LAUNCH_CHECK("k_pack_weights");
  return 0;
}

extern "C" int b3d_linear_tc(const b3d_seg_t* segs, int32_t nseg, const void* Wp, int32_t n_logical,
                             int32_t k_logical, const float* bias, void* Y, int32_t ldy, int32_t y_dtype,
                             int64_t M, int32_t act, int32_t flags, const void* out_mask, int32_t ldm,
                             int32_t mask_dtype, const uint8_t* row_mask, const b3d_seg_t* adds, int32_t nadd,
                             void* relu_bits_out, void* stream) {
  if (M == 0) return 0;
  TcArgs a;
  if (nadd < 0 || nadd > 2 || (nadd && to_dev(adds, nadd, a.add))) return bad_arg("b3d_linear_tc adds");
  for (int q = 0; q < nadd; ++q)
    if (a.add[q].width != n_logical || (a.add[q].ld & (a.add[q].dtype == B3D_BF16 ? 7 : 3)) ||
        (reinterpret_cast<uintptr_t>(a.add[q].ptr) & 15))
      return bad_arg("b3d_linear_tc: adds must be [*, Nout] with 16-byte aligned rows");
  a.nadd = nadd;
  if (to_dev(segs, nseg, a.seg)) return bad_arg("b3d_linear_tc segments");
  int K = 0;
  for (int s = 0; s < nseg; ++s) {
    if (!seg_tc_ok(a.seg[s])) return bad_arg("b3d_linear_tc: segment widths % 8, 16-byte aligned rows, no operand masks");
    K += a.seg[s].width;
  }
  if (K != k_logical) return bad_arg("b3d_linear_tc: sum of segment widths != k_logical");
  if (!Wp || !Y || n_logical <= 0 || M < 0) return bad_arg("b3d_linear_tc W/Y/N/M");
  const bool split = (flags & B3D_FLAG_SPLIT) != 0;   // tf32 x3 arithmetic (Wp packed with B3D_PACK_SPLIT)
  if ((split ? round_up(K, 32) / 4 : round_up(K, TC_BK) / 8) > TC_TAB) return bad_arg("b3d_linear_tc: K too large");
  if (split && (y_dtype != B3D_F32 || relu_bits_out || mask_dtype == B3D_BITS))
    return bad_arg("b3d_linear_tc: the split (1e-4) mode is fp32 in / fp32 out");
  if (y_dtype == B3D_BF16 && ((ldy & 7) || (reinterpret_cast<uintptr_t>(Y) & 15) || (flags & B3D_FLAG_ACCUMULATE)))
    return bad_arg("b3d_linear_tc: bf16 output needs ld % 8 == 0, 16-byte alignment, no accumulate");
  a.nseg = nseg; a.Ktot = K; a.Wp = reinterpret_cast<const __nv_bfloat16*>(Wp);
  a.Npad = round_up(n_logical, 16); a.Kpad = round_up(k_logical, split ? 32 : TC_BK);
  a.bias = bias; a.Y = Y; a.ldy = ldy; a.y_bf16 = (y_dtype == B3D_BF16); a.M = M; a.Nout = n_logical; a.act = act;
  a.flags = flags; a.ldm = ldm; a.mask_bf16 = (mask_dtype == B3D_BF16); a.row_mask = row_mask;
  a.out_mask = mask_dtype == B3D_BITS ? nullptr : out_mask;
  a.mask_bits = mask_dtype == B3D_BITS ? reinterpret_cast<const uint32_t*>(out_mask) : nullptr;
  a.bits_out = reinterpret_cast<uint32_t*>(relu_bits_out);
  int Nb = a.Npad < TC_NMAX ? a.Npad : TC_NMAX;
  size_t smem = (split ? 2 : 1) * (2 * TC_A_STAGE + 2 * (size_t)Nb * 128) + 64 + 1024 + 4 * TC_TAB +
                sizeof(int32_t) * B3D_MAX_SEGS * TC_BM;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_linear_tc<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_linear_tc<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_linear_tc<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_linear_tc<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_linear_tc<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_linear_tc<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
    if (e != cudaSuccess) return fail("k_linear_tc smem attr", e);
    attr_set = true;
  }
  dim3 grid((unsigned)ceil_div(M, TC_BM), (unsigned)ceil_div(a.Npad, TC_NMAX));
  cudaStream_t st = (cudaStream_t)stream;
  if (split) {
    if (act == B3D_ACT_RELU) k_linear_tc<1, true><<<grid, TC_THREADS, smem, st>>>(a);
    else if (act == B3D_ACT_SIGMOID) k_linear_tc<2, true><<<grid, TC_THREADS, smem, st>>>(a);
    else k_linear_tc<0, true><<<grid, TC_THREADS, smem, st>>>(a);
  } else {
    if (act == B3D_ACT_RELU) k_linear_tc<1, false><<<grid, TC_THREADS, smem, st>>>(a);
    else if (act == B3D_ACT_SIGMOID) k_linear_tc<2, false><<<grid, TC_THREADS, smem, st>>>(a);
    else k_linear_tc<0, false><<<grid, TC_THREADS, smem, st>>>(a);
  }
  B3D_LAUNCH_CHECK("k_linear_tc");
  return 0;
}

extern "C" size_t b3d_wgrad_tc_workspace_bytes(int64_t M, int32_t Nout, int32_t K) {
  int S, ktiles, Kp; long long rps;
  wgrad_tc_plan(M > 0 ? M : 1, Nout, K, &S, &rps, &ktiles, &Kp);
  return sizeof(float) * (size_t)S * Nout * (K + 1) + 256;
}

extern "C" int b3d_wgrad_tc(const b3d_seg_t* dy, const b3d_seg_t* segs, int32_t nseg, float* dW, int32_t lddw,
                            float* db, int64_t M, int32_t Nout, int32_t flags, void* workspace,
                            size_t workspace_bytes, void* stream) {
  if (M <= 0) return bad_arg("b3d_wgrad_tc: M must be > 0 (use b3d_wgrad for empty inputs)");
  WgTcArgs a;
  if (to_dev(segs, nseg, a.seg) || to_dev(dy, 1, &a.dy)) return bad_arg("b3d_wgrad_tc segments");
  if (a.dy.idx || a.dy.width != Nout) return bad_arg("b3d_wgrad_tc: dy");
  int K = 0;
  for (int s = 0; s < nseg; ++s) {
    if (!seg_tc_ok(a.seg[s])) return bad_arg("b3d_wgrad_tc: segment widths % 8, 16-byte aligned rows, no masks");
    K += a.seg[s].width;
  }
  {
    SegDev d = a.dy;
    d.width = 8;   // Nout itself may be ragged; only alignment / mask rules apply to dy
    if (!seg_tc_ok(d)) return bad_arg("b3d_wgrad_tc: dy alignment / mask");
  }
  if (K / 8 + 2 > TC_TAB) return bad_arg("b3d_wgrad_tc: K too large");
  int S, ktiles, Kp; long long rps;
  wgrad_tc_plan(M, Nout, K, &S, &rps, &ktiles, &Kp);
  if (workspace_bytes < b3d_wgrad_tc_workspace_bytes(M, Nout, K)) return bad_arg("b3d_wgrad_tc workspace too small");
  a.nseg = nseg; a.Ktot = K; a.Kp = Kp; a.Nout = Nout; a.ktiles = ktiles; a.M = M; a.rows_per_split = rps;
  a.part = reinterpret_cast<float*>(workspace);
  cudaStream_t st = (cudaStream_t)stream;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_wgrad_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_wgrad_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
    if (e != cudaSuccess) return fail("k_wgrad_tc smem attr", e);
    attr_set = true;
  }
  int Nk = Kp < TC_NMAX ? Kp : TC_NMAX;
  const bool split = (flags & B3D_FLAG_SPLIT) != 0;
  size_t smem = (split ? 2 : 1) * (2 * TC_A_STAGE + 2 * (size_t)Nk * 128) + 64 + 4 * TC_TAB;
  dim3 grid((unsigned)S, (unsigned)(ceil_div(Nout, TC_BM) * ktiles));
  if (split) k_wgrad_tc<true><<<grid, TC_THREADS, smem, st>>>(a);
  else k_wgrad_tc<false><<<grid, TC_THREADS, smem, st>>>(a);
  B3D_LAUNCH_CHECK("k_wgrad_tc");
  long long tot = (long long)Nout * (K + 1);
  k_wgrad_tc_reduce<<<(unsigned)ceil_div(tot, 32), 32 * WG_RED_WARPS, 0, st>>>(a.part, S, Nout, K, dW, lddw, db,
                                                                  (flags & B3D_FLAG_ACCUMULATE) ? 1 : 0);
  B3D_LAUNCH_CHECK("k_wgrad_tc_reduce");
  return 0;
}

extern "C" size_t b3d_tma_packed_bytes(int32_t n_logical, int32_t k_logical) {
  return (size_t)round_up(n_logical, 16) * round_up(k_logical, TC_BK) * 2;
}

extern "C" int b3d_tma_pack_weights(const float* W, int32_t ldw, int32_t n_logical, int32_t k_logical,
                                    int32_t transpose, void* Wr, void* stream) {
  if (!W || !Wr || n_logical <= 0 || k_logical <= 0) return bad_arg("b3d_tma_pack_weights");
  int Npad = round_up(n_logical, 16), Kpad = round_up(k_logical, TC_BK);
  long long total = (long long)Npad * Kpad;
  k_pack_weights_rm<<<(unsigned)ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(
      W, ldw, n_logical, k_logical, transpose, reinterpret_cast<__nv_bfloat16*>(Wr), Npad, Kpad);
  B3D_LAUNCH_CHECK("k_pack_weights_rm");
  return 0;
}

static int tma_padded_k(const b3d_seg_t* segs, int nseg) {
  int kp = 0;
  for (int s = 0; s < nseg; ++s) kp += round_up(segs[s].width, TC_BK);
  return kp;
}

extern "C" size_t b3d_tma_packed_bytes_segs(int32_t n_logical, const int32_t* widths, int32_t nseg) {
  size_t kp = 0;
  for (int s = 0; s < nseg; ++s) kp += (size_t)round_up(widths[s], TC_BK);
  return (size_t)round_up(n_logical, 16) * kp * 2;
}

extern "C" int b3d_tma_pack_weights_segs(const float* W, int32_t ldw, int32_t n_logical, const int32_t* widths,
                                         int32_t nseg, int32_t transpose, void* Wr, void* stream) {
  if (!W || !Wr || !widths || n_logical <= 0 || nseg < 1 || nseg > TMA_MAX_SEGS) return bad_arg("b3d_tma_pack_weights_segs");
  SegWidths sw;
  sw.n = nseg;
  int Kpad = 0;
  for (int s = 0; s < nseg; ++s) { sw.w[s] = widths[s]; Kpad += round_up(widths[s], TC_BK); }
  const int Npad = round_up(n_logical, 16);
  long long total = (long long)Npad * Kpad;
  k_pack_weights_rm_segs<<<(unsigned)ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(
      W, ldw, n_logical, transpose, sw, reinterpret_cast<__nv_bfloat16*>(Wr), Npad, Kpad);
  B3D_LAUNCH_CHECK("k_pack_weights_rm_segs");
  return 0;
}

extern "C" int b3d_linear_tma(const b3d_seg_t* segs, int32_t nseg, const void* Wr, int32_t n_logical,
                              int32_t k_logical, const float* bias, void* Y, int32_t ldy, int32_t y_dtype,
                              int64_t M, int32_t act, int32_t flags, const void* out_mask, int32_t ldm,
                              int32_t mask_dtype, const uint8_t* row_mask, const b3d_seg_t* adds, int32_t nadd,
                              void* relu_bits_out, void* stream) {
  if (M == 0) return 0;
  static TmaArgs a;
  SegDev seg[TMA_MAX_SEGS];
  if (nseg < 1 || nseg > TMA_MAX_SEGS || to_dev(segs, nseg, seg)) return bad_arg("b3d_linear_tma: 1 to 6 segments");
  int K = 0;
  a.gidx[0] = a.gidx[1] = nullptr;
  for (int s = 0; s < nseg; ++s) {
    if (seg[s].dtype != B3D_BF16 || !seg_tc_ok(seg[s]))
      return bad_arg("b3d_linear_tma: segments must be bf16, widths % 8, 16-byte aligned rows, no operand masks");
    K += seg[s].width;
    a.seg_chunks[s] = round_up(seg[s].width, TC_BK) / TC_BK;
    a.seg_gsel[s] = -1;
    if (seg[s].idx) {       // row-gathered segment: at most two distinct index arrays per launch
      int q = 0;
      while (q < 2 && a.gidx[q] && a.gidx[q] != seg[s].idx) ++q;
      if (q == 2) return bad_arg("b3d_linear_tma: at most two distinct row-index arrays");
      a.gidx[q] = seg[s].idx;
      a.seg_gsel[s] = q;
    }
  }
  if (K != k_logical || !Wr || !Y || n_logical <= 0) return bad_arg("b3d_linear_tma: shapes");
  if (nadd < 0 || nadd > 2 || (nadd && to_dev(adds, nadd, a.add))) return bad_arg("b3d_linear_tma adds");
  for (int q = 0; q < nadd; ++q)
    if (a.add[q].width != n_logical || (a.add[q].ld & (a.add[q].dtype == B3D_BF16 ? 7 : 3)) ||
        (reinterpret_cast<uintptr_t>(a.add[q].ptr) & 15))
      return bad_arg("b3d_linear_tma: adds must be [*, Nout] with 16-byte aligned rows");
  if (y_dtype == B3D_BF16 && ((ldy & 7) || (reinterpret_cast<uintptr_t>(Y) & 15) || (flags & B3D_FLAG_ACCUMULATE)))
    return bad_arg("b3d_linear_tma: bf16 output needs ld % 8 == 0, 16-byte alignment, no accumulate");
  // every segment starts on a 64-column chunk boundary of the packed weights (b3d_tma_pack_weights_segs; the
  // plain pack b3d_tma_pack_weights is the same thing when only the last segment is ragged)
  const int Npad = round_up(n_logical, 16), Kpad = tma_padded_k(segs, nseg);
  a.nseg = nseg;
  a.nchunks = Kpad / TC_BK;
  // weight block resident in shared memory: Nb * Kpad * 2 <= 144 KB. Sign-bit words cover 32 columns,
  // so with bit masks every column block must start on a word boundary (granularity 32 instead of 16).
  const int gran = (relu_bits_out || mask_dtype == B3D_BITS) ? 32 : 16;
  auto plan_nb = [&](int budget_kb, int* ny_out) {
    int nb = (budget_kb * 1024 / (Kpad * 2)) / gran * gran;
    if (nb > TC_NMAX) nb = TC_NMAX;
    if (nb > round_up(Npad, gran)) nb = round_up(Npad, gran);
    if (nb < gran) { *ny_out = 0; return 0; }
    const int y = (Npad + nb - 1) / nb;
    *ny_out = y;
    return round_up((Npad + y - 1) / y, gran);      // balance the column blocks
  };
  int ny = 0;
  int Nb = plan_nb(144, &ny);
  // Long-K layers (K >= 384): a 192 KB weight block with a 2-stage operand ring when that means fewer column blocks,
  // i.e. fewer passes over the [M, K] operand (512 -> 384: 2 instead of 3; 384 -> 256: 1 instead of 2).
  int wide_stages = 0;
  {
    static int wide = -1;
    if (wide < 0) { const char* e = getenv("B3D_TMA_WIDE"); wide = (e && e[0] == '1') ? 1 : 0; }   // measured slower: opt-in
    int ny2 = 0;
    const int nb2 = wide ? plan_nb(192, &ny2) : 0;
    if (nb2 && ny2 < ny && nadd == 0 &&
        (long long)(Kpad / TC_BK) * nb2 * 128 + 2 * TC_A_STAGE + TMA_BAR_BYTES + 1024 + 1024 + 64 <= 227 * 1024) {
      Nb = nb2; ny = ny2; wide_stages = 2;
    }
  }
  if (Nb < gran) return bad_arg("b3d_linear_tma: K too large for a resident weight block");
  a.Nb = Nb;
  a.acc_stride = (int)tmem_cols_for(Nb);
  a.ntiles = ceil_div(M, TC_BM);
  a.bias = bias; a.Y = Y; a.ldy = ldy; a.y_bf16 = (y_dtype == B3D_BF16); a.M = M; a.Nout = n_logical; a.act = act;
  a.flags = flags; a.ldm = ldm; a.mask_bf16 = (mask_dtype == B3D_BF16); a.row_mask = row_mask;
  a.out_mask = mask_dtype == B3D_BITS ? nullptr : out_mask;
  a.mask_bits = mask_dtype == B3D_BITS ? reinterpret_cast<const uint32_t*>(out_mask) : nullptr;
  a.bits_out = reinterpret_cast<uint32_t*>(relu_bits_out);
  a.nadd = nadd;
  alignas(64) static TmaMaps maps;
  alignas(64) CUtensorMap mW;
  for (int s = 0; s < TMA_MAX_SEGS; ++s) {
    const int src = s < nseg ? s : 0;
    // gathered segments: the tensor is the whole source table (its row count is not part of b3d_seg_t; the indices
    // address it, so a bound of 2^28 rows is encoded) and the box is one row; dense segments: M rows, box of 128 rows
    const long long rows = seg[src].idx ? (1ll << 28) : M;
    if (make_tmap_bf16(&maps.m[s], seg[src].ptr, rows, seg[src].width, seg[src].ld, seg[src].idx ? 1 : TC_BM))
      return bad_arg("b3d_linear_tma: tensor map A");
  }
  // Staging plan for the row-gathered addends (see the producer warps 10-11): prefer double-buffered tiles; all
  // addends if they fit, else the LAST one (the source-side rows: targets are sorted, so the first addend's rows
  // repeat along a tile and are cheap to load from the epilogue). The operand ring shrinks to 2 stages for short K
  // when that buys the second buffer.
  a.stage_mask = 0; a.add_bufs = 1; a.add_slots = 0; a.add_tile_bytes = TC_BM * Nb * 2;
  a.stages = wide_stages ? wide_stages : TMA_STAGES;
  {
    static int enabled = -1, split = -1, fast = -1;
    if (enabled < 0) { const char* e = getenv("B3D_STAGE_ADDENDS"); enabled = (e && e[0] == '0') ? 0 : 1; }
    if (split < 0) { const char* e = getenv("B3D_STAGE_SPLIT"); split = (e && e[0] == '0') ? 0 : 1; }
    if (fast < 0) { const char* e = getenv("B3D_FAST_EPI"); fast = (e && e[0] == '0') ? 0 : 1; }
    a.fast_epi = fast;
    { static int ca = -1; if (ca < 0) { const char* e = getenv("B3D_ADD_CA"); ca = (e && e[0] == '1') ? 1 : 0; } a.add_ca = ca; }   // measured 2-6 % slower: opt-in
    bool ok = enabled && nadd > 0 && y_dtype == B3D_BF16;
    for (int q = 0; q < nadd; ++q) ok = ok && a.add[q].dtype == B3D_BF16;
    const long long limit = 227 * 1024;
    auto fits = [&](int nb, int n, int bufs, int stages) {
      return (long long)a.nchunks * nb * 128 + TMA_BAR_BYTES + 1024 + 1024 + 64 + TMA_ID_BYTES + (long long)stages * TC_A_STAGE +
                 (long long)n * bufs * TC_BM * nb * 2 <= limit;
    };
    // narrower column blocks when that lets EVERY addend be staged and double-buffered (the operand tile is then read
    // once per column block, from L2 after the first): all addends in the accumulator = the lean epilogue
    const int st_min = a.nchunks > 2 ? 4 : 2;
    if (ok && split && (n_logical % 64) == 0 && (Nb % 64) == 0 && !fits(Nb, nadd, 2, st_min)) {
      for (int nb = Nb - 64; nb >= 64; nb -= 64) {
        if ((n_logical % nb) || !fits(nb, nadd, 2, st_min)) continue;
        Nb = nb;
        ny = n_logical / nb;
        break;
      }
    }
    a.Nb = Nb;
    a.acc_stride = (int)tmem_cols_for(Nb);
    a.add_tile_bytes = TC_BM * Nb * 2;
    ok = ok && (Nb % 64) == 0 && (n_logical % Nb) == 0;
    if (ok) {
      // all addends staged (lean epilogue) before a partial plan; double-buffered before single
      struct { int n, bufs, stages; } opts[] = {{nadd, 2, 4}, {nadd, 2, 2}, {nadd, 1, 4}, {nadd, 1, 2},
                                                {1, 2, 4}, {1, 2, 2}, {1, 1, 4}, {0, 0, 0}};
      static int forced[3] = {-1, 0, 0};
      if (forced[0] == -1) {
        forced[0] = 0;
        const char* e = getenv("B3D_STAGE_PLAN");       // "n,bufs,stages": experiments only
        if (e && sscanf(e, "%d,%d,%d", &forced[0], &forced[1], &forced[2]) != 3) forced[0] = 0;
      }
      if (forced[0] > 0 && forced[0] <= nadd) opts[0] = {forced[0] == 2 ? nadd : 1, forced[1], forced[2]};
      for (auto& o : opts) {
        if (o.n == 0) break;
        if (o.stages == 2 && a.nchunks > 2) continue;
        if (!fits(Nb, o.n, o.bufs, o.stages)) continue;
        a.add_slots = o.n; a.add_bufs = o.bufs; a.stages = o.stages;
        a.stage_mask = o.n == nadd ? (1 << nadd) - 1 : 1 << (nadd - 1);
        break;
      }
    }
  }
  if (make_tmap_bf16(&mW, Wr, Npad, Kpad, Kpad, Nb)) return bad_arg("b3d_linear_tma: tensor map W");
  // the operand ring takes what is left, up to 8 stages: the ring is the bytes in flight per SM, and a long-K tile
  // (512 -> 384: 128 KB of operand per tile visit) is bound by ring size / TMA latency, not by L2 or HBM bandwidth
  const size_t rest = (size_t)a.nchunks * Nb * 128 + (size_t)a.add_bufs * a.add_slots * a.add_tile_bytes +
                      (a.stage_mask ? TMA_ID_BYTES : 0) + TMA_BAR_BYTES + 1024 + 1024;   // + alignment slack
  {
    static int deep = -1;
    if (deep < 0) { const char* e = getenv("B3D_TMA_DEEP"); deep = (e && e[0] == '0') ? 0 : 1; }
    if (deep && !wide_stages && a.nchunks >= 4) {      // short-K layers measured no faster / slower with a deeper ring
      int st = (int)((227 * 1024 - (long long)rest) / TC_A_STAGE);
      if (st > 8) st = 8;
      if (st > a.stages) a.stages = st;
    }
  }
  size_t smem = rest + (size_t)a.stages * TC_A_STAGE;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_linear_tma<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_linear_tma<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_linear_tma<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_linear_tma<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_linear_tma<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return fail("k_linear_tma smem attr", e);
    attr_set = true;
  }
  static int n_sm = 0;
  if (!n_sm) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    if (n_sm <= 0) n_sm = 148;
  }
  long long gx = n_sm / ny;
  if (gx < 1) gx = 1;
  if (gx > a.ntiles) gx = a.ntiles;
  cudaStream_t st = (cudaStream_t)stream;
  // the LEAN build when every block of every CTA is a lean block (see TmaCfg)
  const bool lean = a.fast_epi && act != B3D_ACT_SIGMOID && a.y_bf16 && !a.out_mask && !a.row_mask &&
                    !(flags & B3D_FLAG_ACCUMULATE) && (nadd == 0 || a.stage_mask == (1 << nadd) - 1) && (ldy & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(Y) & 31) == 0 && (Nb & 31) == 0 && (n_logical & 31) == 0;
  const void* fn = lean ? (act == B3D_ACT_RELU ? (const void*)k_linear_tma<1, true> : (const void*)k_linear_tma<0, true>)
                        : (act == B3D_ACT_RELU ? (const void*)k_linear_tma<1, false>
                           : act == B3D_ACT_SIGMOID ? (const void*)k_linear_tma<2, false> : (const void*)k_linear_tma<0, false>);
  const unsigned threads = lean ? TmaCfg<true>::THREADS : TmaCfg<false>::THREADS;
  // Cluster mode: the ny column-block CTAs of a row tile as one thread-block cluster sharing the operand loads (TMA
  // multicast). All clusters must be resident together (the tile loop is a static partition over the clusters).
  a.cluster = 1;
  {
    static int want = -1;
    // measured SLOWER than independent CTAs (512 -> 384: 1,217 vs 1,033 us; profiles/r2_lean_epilogue.md): opt-in
    if (want < 0) { const char* e = getenv("B3D_TMA_CLUSTER"); want = (e && e[0] == '1') ? 1 : 0; }
    bool dense = true;
    for (int sg = 0; sg < nseg; ++sg) dense = dense && a.seg_gsel[sg] < 0;
    if (want && ny >= 2 && ny <= 8 && dense) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)gx, (unsigned)ny);
      cfg.blockDim = dim3(threads);
      cfg.dynamicSmemBytes = smem;
      cfg.stream = st;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = (unsigned)ny; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int ncl = 0;
      if (cudaOccupancyMaxActiveClusters(&ncl, fn, &cfg) == cudaSuccess && ncl >= 1) {
        if (gx > ncl) gx = ncl;
        cfg.gridDim = dim3((unsigned)gx, (unsigned)ny);
        a.cluster = ny;
        void* args[] = {(void*)&maps, (void*)&mW, (void*)&a};
        cudaError_t e = cudaLaunchKernelExC(&cfg, fn, args);
        if (e != cudaSuccess) return fail("k_linear_tma cluster launch", e);
        B3D_LAUNCH_CHECK("k_linear_tma");
        return 0;
      }
      (void)cudaGetLastError();
    }
  }
  dim3 grid((unsigned)gx, (unsigned)ny);
  void* args[] = {(void*)&maps, (void*)&mW, (void*)&a};
  cudaError_t le = cudaLaunchKernel(fn, grid, dim3(threads), args, smem, st);
  if (le != cudaSuccess) return fail("k_linear_tma launch", le);
  B3D_LAUNCH_CHECK("k_linear_tma");
  return 0;
}

extern "C" size_t b3d_wgrad_tma_workspace_bytes(int64_t M, int32_t Nout, int32_t K) {
  // worst case: one split per 256 rows is never exceeded by the plan below
  long long tiles = (long long)((Nout + TC_BM - 1) / TC_BM) * ((round_up(K, 64) / 64 + 3) / 4);
  long long S = (148 + tiles - 1) / tiles;   // upper bound of the launch plan's split count
  long long smax = (M + WG_MIN_ROWS - 1) / WG_MIN_ROWS;
  if (S > smax) S = smax;
  if (S < 1) S = 1;
  return sizeof(float) * (size_t)S * Nout * (K + 1) + 256;
}

extern "C" int b3d_wgrad_tma(const b3d_seg_t* dy, const b3d_seg_t* segs, int32_t nseg, float* dW, int32_t lddw,
                             float* db, int64_t M, int32_t Nout, int32_t flags, void* workspace,
                             size_t workspace_bytes, void* stream) {
  if (M <= 0) return bad_arg("b3d_wgrad_tma: M must be > 0");
  SegDev seg[2], d;
  if (nseg < 1 || nseg > 2 || to_dev(segs, nseg, seg) || to_dev(dy, 1, &d)) return bad_arg("b3d_wgrad_tma segments");
  if (d.idx || d.width != Nout || d.dtype != B3D_BF16 || (d.ld & 7) || (reinterpret_cast<uintptr_t>(d.ptr) & 15) ||
      d.mask_mode != B3D_MASK_NONE)
    return bad_arg("b3d_wgrad_tma: dy must be dense bf16 with 16-byte aligned rows");
  int K = 0;
  for (int s = 0; s < nseg; ++s) {
    if (seg[s].dtype != B3D_BF16 || seg[s].idx || !seg_tc_ok(seg[s]))
      return bad_arg("b3d_wgrad_tma: segments must be dense bf16, widths % 8, 16-byte aligned rows");
    if (s + 1 < nseg && (seg[s].width % 64)) return bad_arg("b3d_wgrad_tma: leading segment width % 64");
    K += seg[s].width;
  }
  WgTmaArgs a;
  a.seg0_groups = nseg == 2 ? seg[0].width / 64 : round_up(K, 64) / 64;
  a.ngroups_total = nseg == 2 ? seg[0].width / 64 + round_up(seg[1].width, 64) / 64 : round_up(K, 64) / 64;
  a.Ktot = K; a.Nout = Nout; a.M = M;
  a.ktiles = (a.ngroups_total + 3) / 4;
  long long tiles = (long long)((Nout + TC_BM - 1) / TC_BM) * a.ktiles;
  long long S = 148 / tiles;          // ONE wave: S * tiles <= 148 CTAs (rounding up put 150 CTAs on 148 SMs)
  // at least 4 chunks of 64 rows per CTA. (1,024 rows per CTA left a 9,629-row launch — the reference's own 2-window
  // training batch — on 9 CTAs walking 16 chunks each: 19.6 us per launch, 25 % of that step.)
  long long smax = (M + WG_MIN_ROWS - 1) / WG_MIN_ROWS;
  if (S > smax) S = smax;
  if (S < 1) S = 1;
  long long rps = ((M + S - 1) / S + 63) / 64 * 64;
  S = (M + rps - 1) / rps;
  a.rows_per_split = rps;
  if (workspace_bytes < sizeof(float) * (size_t)S * Nout * (K + 1)) return bad_arg("b3d_wgrad_tma workspace too small");
  a.part = reinterpret_cast<float*>(workspace);
  alignas(64) CUtensorMap mDY, mA0, mA1;
  if (make_tmap_bf16(&mDY, d.ptr, M, Nout, d.ld, 64)) return bad_arg("b3d_wgrad_tma: tensor map dY");
  if (make_tmap_bf16(&mA0, seg[0].ptr, M, seg[0].width, seg[0].ld, 64)) return bad_arg("b3d_wgrad_tma: tensor map A0");
  if (nseg == 2) {
    if (make_tmap_bf16(&mA1, seg[1].ptr, M, seg[1].width, seg[1].ld, 64)) return bad_arg("b3d_wgrad_tma: tensor map A1");
  } else {
    mA1 = mA0;
  }
  cudaStream_t st = (cudaStream_t)stream;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_wgrad_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return fail("k_wgrad_tma smem attr", e);
    attr_set = true;
  }
  size_t smem = (size_t)WG_STAGES * 6 * WG_BOX + WG_BOX + 128 + 1024;
  dim3 grid((unsigned)S, (unsigned)tiles);
  k_wgrad_tma<<<grid, WG_THREADS, smem, st>>>(mDY, mA0, mA1, a);
  B3D_LAUNCH_CHECK("k_wgrad_tma");
  long long tot = (long long)Nout * (K + 1);
  k_wgrad_tc_reduce<<<(unsigned)ceil_div(tot, 32), 32 * WG_RED_WARPS, 0, st>>>(a.part, (int)S, Nout, K, dW, lddw, db,
                                                                  (flags & B3D_FLAG_ACCUMULATE) ? 1 : 0);
  B3D_LAUNCH_CHECK("k_wgrad_tc_reduce");
  return 0;
}
