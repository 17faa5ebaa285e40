// bf16 tensor-core ("2e-2 mode") dense layers for sm_100a: tcgen05.mma with fp32 accumulators
// in TMEM, operands staged in shared memory in the no-swizzle canonical UMMA layouts.
//
//   b3d_linear_tc : Y = act(cat_s(gather(A_s)) W^T + b)   (forward and, with a transposed pack,
//                   the input gradient).  A: fp32 in HBM, converted to bf16 while staging
//                   (K-major canonical layout); W: pre-packed bf16 slabs (b3d_tc_pack_weights).
//   b3d_wgrad_tc  : dW = dY^T cat_s(A_s), db = colsum(dY).  Both operands are staged MN-major
//                   (rows are the reduction dimension), rows split over CTAs, fixed-order reduce.
//
// Pipeline (both kernels): 128 threads stage chunk c+1 while the tensor core works on chunk c
// (tcgen05.mma is asynchronous; a tcgen05.commit -> mbarrier frees the stage). Two CTAs per SM
// overlap one CTA's epilogue with the other's main loop.
#include "b3d_common.cuh"
#include "tc_common.cuh"

namespace b3d {


constexpr int TC_BM = 128;      // rows per CTA (UMMA M)
constexpr int TC_BK = 64;       // K elements staged per chunk (4 MMAs of K=16)
constexpr int TC_NMAX = 256;    // max UMMA N per CTA
constexpr int TC_THREADS = 128;
constexpr int TC_A_STAGE = TC_BM * TC_BK * 2;  // 16 KB

__host__ __device__ inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
__host__ __device__ inline uint32_t tmem_cols_for(int n) { return n <= 32 ? 32u : n <= 64 ? 64u : n <= 128 ? 128u : 256u; }

// ------------------------------------------------------------------ weight packing
// Wp[((k/8) * Npad + n) * 8 + k%8] = bf16(B[n][k]),  B[n][k] = transpose ? W[k*ldw+n] : W[n*ldw+k],
// zero padded to Npad = round_up(n_logical,16), Kpad = round_up(k_logical,64): every (chunk, k-group)
// slab is a contiguous run of 16-byte rows == the shared-memory image of a K-major operand.
__global__ void k_pack_weights(const float* __restrict__ W, int ldw, int n_log, int k_log, int transpose,
                               __nv_bfloat16* __restrict__ Wp, int Npad, int Kpad) {
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
  Wp[i] = __float2bfloat16_rn(v);
}

// ------------------------------------------------------------------ forward / dgrad
struct TcArgs {
  SegDev seg[B3D_MAX_SEGS];
  int nseg, Ktot;
  const __nv_bfloat16* Wp;
  int Npad, Kpad;
  const float* bias;
  float* Y;
  int ldy;
  long long M;
  int Nout, act, flags;
  const float* out_mask;
  int ldm;
  const uint8_t* row_mask;
};

__global__ void __launch_bounds__(TC_THREADS) k_linear_tc(const TcArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  using namespace tc;
  const int tid = threadIdx.x, warp = tid >> 5;
  const long long m0 = (long long)blockIdx.x * TC_BM;
  const int n0 = blockIdx.y * TC_NMAX;
  const int Nb = min(TC_NMAX, a.Npad - n0);
  const uint32_t b_stage = (uint32_t)Nb * (TC_BK * 2);   // Nb rows x 128 B
  const uint32_t sA = smem_u32(smem);
  const uint32_t sB = sA + 2 * TC_A_STAGE;
  const uint32_t sBar = sB + 2 * b_stage;                // free[0], free[1], done, tmem ptr
  int32_t* s_grow = reinterpret_cast<int32_t*>(smem + 2 * TC_A_STAGE + 2 * b_stage + 64);  // [nseg][128]
  volatile uint32_t* s_tmem = reinterpret_cast<volatile uint32_t*>(smem + 2 * TC_A_STAGE + 2 * b_stage + 24);
  const uint32_t ncols = tmem_cols_for(Nb);

  if (warp == 0) tmem_alloc(sBar + 24, ncols);
  if (tid == 0) {
    mbar_init(sBar + 0, 1);
    mbar_init(sBar + 8, 1);
    mbar_init(sBar + 16, 1);
    fence_mbar_init();
  }
  const long long row = m0 + tid;
  const bool row_ok = row < a.M;
  for (int s = 0; s < a.nseg; ++s)
    s_grow[s * TC_BM + tid] = row_ok ? (a.seg[s].idx ? __ldg(a.seg[s].idx + row) : (int32_t)row) : 0;
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *s_tmem;

  const int nchunks = a.Kpad / TC_BK;
  const uint32_t idesc = make_idesc_bf16(TC_BM, (uint32_t)Nb, 0, 0);
  const uint32_t lbo_a = TC_BM * 16, lbo_b = (uint32_t)Nb * 16, sbo = 128;

  for (int c = 0; c < nchunks; ++c) {
    const int s = c & 1;
    if (c >= 2) mbar_wait(sBar + 8 * s, ((c >> 1) - 1) & 1);   // MMAs of chunk c-2 released this stage
    // ---- B: contiguous pre-packed slabs -> cp.async
    {
      const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(a.Wp) + ((size_t)(c * 8) * a.Npad + n0) * 16;
      const uint32_t dst = sB + s * b_stage;
#pragma unroll
      for (int g = 0; g < 8; ++g)
        for (int n = tid; n < Nb; n += TC_THREADS)
          cp_async16(dst + (uint32_t)(g * Nb + n) * 16, wsrc + ((size_t)g * a.Npad + n) * 16);
    }
    // ---- A: one row per thread, 64 columns = 8 groups of 8 (gather + fp32->bf16)
    {
      float4 v[16];
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        int off = c * TC_BK + g * 8, sg = 0;
        while (sg < a.nseg && off >= a.seg[sg].width) { off -= a.seg[sg].width; ++sg; }
        if (row_ok && sg < a.nseg) {
          const SegDev& S = a.seg[sg];
          const long long gr = s_grow[sg * TC_BM + tid];
          const float4* p = reinterpret_cast<const float4*>(S.ptr + gr * S.ld + off);
          v[2 * g] = __ldg(p);
          v[2 * g + 1] = __ldg(p + 1);
          if (S.mask_mode != B3D_MASK_NONE) {
            const float4* mp = reinterpret_cast<const float4*>(S.mask + gr * S.ldmask + off);
            const float4 m0v = __ldg(mp), m1v = __ldg(mp + 1);
            v[2 * g].x = apply_mask(v[2 * g].x, m0v.x, S.mask_mode); v[2 * g].y = apply_mask(v[2 * g].y, m0v.y, S.mask_mode);
            v[2 * g].z = apply_mask(v[2 * g].z, m0v.z, S.mask_mode); v[2 * g].w = apply_mask(v[2 * g].w, m0v.w, S.mask_mode);
            v[2 * g + 1].x = apply_mask(v[2 * g + 1].x, m1v.x, S.mask_mode); v[2 * g + 1].y = apply_mask(v[2 * g + 1].y, m1v.y, S.mask_mode);
            v[2 * g + 1].z = apply_mask(v[2 * g + 1].z, m1v.z, S.mask_mode); v[2 * g + 1].w = apply_mask(v[2 * g + 1].w, m1v.w, S.mask_mode);
          }
        } else {
          v[2 * g] = make_float4(0.f, 0.f, 0.f, 0.f);
          v[2 * g + 1] = v[2 * g];
        }
      }
      const uint32_t dst = sA + s * TC_A_STAGE + tid * 16;
#pragma unroll
      for (int g = 0; g < 8; ++g)
        st_shared_v4(dst + g * (TC_BM * 16), pack_bf16x2(v[2 * g].x, v[2 * g].y), pack_bf16x2(v[2 * g].z, v[2 * g].w),
                     pack_bf16x2(v[2 * g + 1].x, v[2 * g + 1].y), pack_bf16x2(v[2 * g + 1].z, v[2 * g + 1].w));
    }
    cp_async_wait_all();
    fence_proxy_async_smem();      // generic-proxy smem writes -> visible to the tensor core (async proxy)
    __syncthreads();
    if (tid == 0) {
      tc_fence_after_sync();
#pragma unroll
      for (int j = 0; j < TC_BK / 16; ++j) {
        const uint32_t aaddr = sA + s * TC_A_STAGE + j * 2 * lbo_a;
        const uint32_t baddr = sB + s * b_stage + j * 2 * lbo_b;
        const uint64_t ad = make_smem_desc(aaddr, lbo_a, sbo);
        const uint64_t bd = make_smem_desc(baddr, lbo_b, sbo);
        mma_bf16_ss(tmem, ad, bd, idesc, (c | j) != 0);
      }
      mma_commit(sBar + 8 * s);
      if (c == nchunks - 1) mma_commit(sBar + 16);
    }
  }
  mbar_wait(sBar + 16, 0);
  tc_fence_after_sync();

  // ---- epilogue: TMEM -> registers -> bias / activation / masks -> global (row per thread)
  const bool rz = row_ok && a.row_mask && a.row_mask[row] == 0;
  const bool vec_ok = ((a.ldy & 3) == 0) && ((reinterpret_cast<uintptr_t>(a.Y) & 15) == 0);
  for (int col0 = 0; col0 < Nb; col0 += 32) {
    uint32_t r[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)col0, r);
    tmem_ld_wait();
    if (!row_ok) continue;
    float* yrow = a.Y + row * a.ldy;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int cbase = n0 + col0 + 4 * q;
      if (cbase >= a.Nout) break;
      float o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int cc = cbase + j;
        float val = __uint_as_float(r[4 * q + j]);
        if (cc < a.Nout) {
          if (a.bias) val += __ldg(a.bias + cc);
          if (a.act == B3D_ACT_RELU) val = fmaxf(val, 0.f);
          else if (a.act == B3D_ACT_SIGMOID) val = 1.f / (1.f + expf(-val));
          if (a.out_mask) val = (__ldg(a.out_mask + row * a.ldm + cc) > 0.f) ? val : 0.f;
          if (rz) val = 0.f;
        }
        o[j] = val;
      }
      if (vec_ok && cbase + 3 < a.Nout) {
        float4* yp = reinterpret_cast<float4*>(yrow + cbase);
        if (a.flags & B3D_FLAG_ACCUMULATE) { float4 p = *yp; o[0] += p.x; o[1] += p.y; o[2] += p.z; o[3] += p.w; }
        *yp = make_float4(o[0], o[1], o[2], o[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (cbase + j < a.Nout) {
            float* yp = yrow + cbase + j;
            *yp = (a.flags & B3D_FLAG_ACCUMULATE) ? *yp + o[j] : o[j];
          }
      }
    }
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

__global__ void __launch_bounds__(TC_THREADS) k_wgrad_tc(const WgTcArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  using namespace tc;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int split = blockIdx.x;
  const int nt = blockIdx.y / a.ktiles, kt = blockIdx.y % a.ktiles;
  const int n0 = nt * TC_BM, k0 = kt * TC_NMAX;
  const int Nk = min(TC_NMAX, a.Kp - k0);
  const uint32_t b_stage = (uint32_t)Nk * 128;
  const uint32_t sA = smem_u32(smem);
  const uint32_t sB = sA + 2 * TC_A_STAGE;
  const uint32_t sBar = sB + 2 * b_stage;
  int32_t* s_grow = reinterpret_cast<int32_t*>(smem + 2 * TC_A_STAGE + 2 * b_stage + 64);  // [nseg][64]
  volatile uint32_t* s_tmem = reinterpret_cast<volatile uint32_t*>(smem + 2 * TC_A_STAGE + 2 * b_stage + 24);
  const uint32_t ncols = tmem_cols_for(Nk);
  if (warp == 0) tmem_alloc(sBar + 24, ncols);
  if (tid == 0) {
    mbar_init(sBar + 0, 1);
    mbar_init(sBar + 8, 1);
    mbar_init(sBar + 16, 1);
    fence_mbar_init();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *s_tmem;

  const long long r0 = (long long)split * a.rows_per_split;
  const long long r1 = min(a.M, r0 + a.rows_per_split);
  const int nchunks = (int)((r1 - r0 + TC_BK - 1) / TC_BK);
  const uint32_t idesc = make_idesc_bf16(TC_BM, (uint32_t)Nk, 1, 1);
  const int rl = tid & 7;        // row within an 8-row group
  const int gq = tid >> 3;       // 0..15: which 8-wide MN group this thread starts at

  for (int c = 0; c < nchunks; ++c) {
    const int s = c & 1;
    const long long rbase = r0 + (long long)c * TC_BK;
    if (c >= 2) mbar_wait(sBar + 8 * s, ((c >> 1) - 1) & 1);
    // gathered source rows of this chunk, per segment
    __syncthreads();   // previous chunk's readers of s_grow are done
    for (int i = tid; i < a.nseg * TC_BK; i += TC_THREADS) {
      const int sg = i / TC_BK;
      const long long r = rbase + (i % TC_BK);
      s_grow[i] = (r < r1) ? (a.seg[sg].idx ? __ldg(a.seg[sg].idx + r) : (int32_t)r) : -1;
    }
    __syncthreads();
    // ---- A' = dY^T chunk: [128 n][64 r]
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int rr = i * 8 + rl;
      const long long r = rbase + rr;
      const int n = n0 + gq * 8;
      float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
      if (r < r1 && n < a.Nout) {
        const float* p = a.dy.ptr + r * a.dy.ld + n;
        if (n + 7 < a.Nout) {
          v0 = __ldg(reinterpret_cast<const float4*>(p));
          v1 = __ldg(reinterpret_cast<const float4*>(p) + 1);
          if (a.dy.mask_mode != B3D_MASK_NONE) {
            const float4* mp = reinterpret_cast<const float4*>(a.dy.mask + r * a.dy.ldmask + n);
            const float4 m0v = __ldg(mp), m1v = __ldg(mp + 1);
            v0.x = apply_mask(v0.x, m0v.x, a.dy.mask_mode); v0.y = apply_mask(v0.y, m0v.y, a.dy.mask_mode);
            v0.z = apply_mask(v0.z, m0v.z, a.dy.mask_mode); v0.w = apply_mask(v0.w, m0v.w, a.dy.mask_mode);
            v1.x = apply_mask(v1.x, m1v.x, a.dy.mask_mode); v1.y = apply_mask(v1.y, m1v.y, a.dy.mask_mode);
            v1.z = apply_mask(v1.z, m1v.z, a.dy.mask_mode); v1.w = apply_mask(v1.w, m1v.w, a.dy.mask_mode);
          }
        } else {
          float t[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            t[j] = 0.f;
            if (n + j < a.Nout) {
              t[j] = __ldg(p + j);
              if (a.dy.mask_mode != B3D_MASK_NONE)
                t[j] = apply_mask(t[j], __ldg(a.dy.mask + r * a.dy.ldmask + n + j), a.dy.mask_mode);
            }
          }
          v0 = make_float4(t[0], t[1], t[2], t[3]); v1 = make_float4(t[4], t[5], t[6], t[7]);
        }
      }
      st_shared_v4(sA + s * TC_A_STAGE + gq * 1024 + i * 128 + rl * 16, pack_bf16x2(v0.x, v0.y), pack_bf16x2(v0.z, v0.w),
                   pack_bf16x2(v1.x, v1.y), pack_bf16x2(v1.z, v1.w));
    }
    // ---- B' = Acat^T chunk: [Nk k][64 r]
    for (int kg = gq; kg < Nk / 8; kg += 16) {
      const int col = k0 + kg * 8;
      int off = col, sg = 0;
      while (sg < a.nseg && off >= a.seg[sg].width) { off -= a.seg[sg].width; ++sg; }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int rr = i * 8 + rl;
        float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
        if (sg < a.nseg) {
          const int gr = s_grow[sg * TC_BK + rr];
          if (gr >= 0) {
            const SegDev& S = a.seg[sg];
            const float4* p = reinterpret_cast<const float4*>(S.ptr + (long long)gr * S.ld + off);
            v0 = __ldg(p);
            v1 = __ldg(p + 1);
          }
        } else if (col == a.Ktot && rbase + rr < r1) {
          v0.x = 1.f;   // virtual ones column -> bias gradient
        }
        st_shared_v4(sB + s * b_stage + kg * 1024 + i * 128 + rl * 16, pack_bf16x2(v0.x, v0.y), pack_bf16x2(v0.z, v0.w),
                     pack_bf16x2(v1.x, v1.y), pack_bf16x2(v1.z, v1.w));
      }
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after_sync();
#pragma unroll
      for (int j = 0; j < TC_BK / 16; ++j) {
        const uint32_t aaddr = sA + s * TC_A_STAGE + j * 256;
        const uint32_t baddr = sB + s * b_stage + j * 256;
        const uint64_t ad = make_smem_desc(aaddr, 128, 1024);   // LBO: next 8-row K group, SBO: next 8-wide MN group
        const uint64_t bd = make_smem_desc(baddr, 128, 1024);
        mma_bf16_ss(tmem, ad, bd, idesc, (c | j) != 0);
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
  const int n = n0 + tid;   // TMEM lane == output row n
  for (int col0 = 0; col0 < Nk; col0 += 32) {
    uint32_t r[32];
    if (nchunks > 0) {
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)col0, r);
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

__global__ void k_wgrad_tc_reduce(const float* __restrict__ part, int S, int Nout, int Ktot,
                                  float* __restrict__ dW, int lddw, float* __restrict__ db, int accumulate) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int Kw = Ktot + 1;
  const long long tot = (long long)Nout * Kw;
  if (i >= tot) return;
  float s = 0.f;
  for (int p = 0; p < S; ++p) s += part[(long long)p * tot + i];   // fixed order: deterministic
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
  long long smax = (M + 511) / 512;
  if (s > smax) s = smax;
  if (s < 1) s = 1;
  long long r = ((M + s - 1) / s + TC_BK - 1) / TC_BK * TC_BK;
  *S = (int)((M + r - 1) / r);
  *rps = r;
}

static bool segs_tc_ok(const SegDev* seg, int nseg) {
  for (int s = 0; s < nseg; ++s)
    if ((seg[s].width & 7) || (seg[s].ld & 3) || (reinterpret_cast<uintptr_t>(seg[s].ptr) & 15) ||
        (seg[s].mask_mode != B3D_MASK_NONE && ((seg[s].ldmask & 3) || (reinterpret_cast<uintptr_t>(seg[s].mask) & 15))))
      return false;
  return true;
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
  k_pack_weights<<<(unsigned)ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(
      W, ldw, n_logical, k_logical, transpose, reinterpret_cast<__nv_bfloat16*>(Wp), Npad, Kpad);
  B3D_LAUNCH_CHECK("k_pack_weights");
  return 0;
}

extern "C" int b3d_linear_tc(const b3d_seg_t* segs, int32_t nseg, const void* Wp, int32_t n_logical,
                             int32_t k_logical, const float* bias, float* Y, int32_t ldy, int64_t M, int32_t act,
                             int32_t flags, const float* out_mask, int32_t ldm, const uint8_t* row_mask,
                             void* stream) {
  if (M == 0) return 0;
  TcArgs a;
  if (to_dev(segs, nseg, a.seg)) return bad_arg("b3d_linear_tc segments");
  if (!segs_tc_ok(a.seg, nseg)) return bad_arg("b3d_linear_tc: segment widths must be multiples of 8, 16-byte aligned");
  int K = 0;
  for (int s = 0; s < nseg; ++s) K += a.seg[s].width;
  if (K != k_logical) return bad_arg("b3d_linear_tc: sum of segment widths != k_logical");
  if (!Wp || !Y || n_logical <= 0 || M < 0) return bad_arg("b3d_linear_tc W/Y/N/M");
  a.nseg = nseg; a.Ktot = K; a.Wp = reinterpret_cast<const __nv_bfloat16*>(Wp);
  a.Npad = round_up(n_logical, 16); a.Kpad = round_up(k_logical, TC_BK);
  a.bias = bias; a.Y = Y; a.ldy = ldy; a.M = M; a.Nout = n_logical; a.act = act; a.flags = flags;
  a.out_mask = out_mask; a.ldm = ldm; a.row_mask = row_mask;
  int Nb = a.Npad < TC_NMAX ? a.Npad : TC_NMAX;
  size_t smem = 2 * TC_A_STAGE + 2 * (size_t)Nb * 128 + 64 + sizeof(int32_t) * B3D_MAX_SEGS * TC_BM;
  static size_t smem_set = 0;
  if (smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(k_linear_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
    if (e != cudaSuccess) return fail("k_linear_tc smem attr", e);
    smem_set = 112 * 1024;
  }
  dim3 grid((unsigned)ceil_div(M, TC_BM), (unsigned)ceil_div(a.Npad, TC_NMAX));
  k_linear_tc<<<grid, TC_THREADS, smem, (cudaStream_t)stream>>>(a);
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
  for (int s = 0; s < nseg; ++s)
    if (a.seg[s].mask_mode != B3D_MASK_NONE) return bad_arg("b3d_wgrad_tc: masked A segments unsupported");
  if (!segs_tc_ok(a.seg, nseg) || (a.dy.ld & 3) || (reinterpret_cast<uintptr_t>(a.dy.ptr) & 15) ||
      (a.dy.mask_mode != B3D_MASK_NONE && ((a.dy.ldmask & 3) || (reinterpret_cast<uintptr_t>(a.dy.mask) & 15))))
    return bad_arg("b3d_wgrad_tc: alignment (widths % 8, 16-byte aligned rows)");
  int K = 0;
  for (int s = 0; s < nseg; ++s) K += a.seg[s].width;
  int S, ktiles, Kp; long long rps;
  wgrad_tc_plan(M, Nout, K, &S, &rps, &ktiles, &Kp);
  if (workspace_bytes < b3d_wgrad_tc_workspace_bytes(M, Nout, K)) return bad_arg("b3d_wgrad_tc workspace too small");
  a.nseg = nseg; a.Ktot = K; a.Kp = Kp; a.Nout = Nout; a.ktiles = ktiles; a.M = M; a.rows_per_split = rps;
  a.part = reinterpret_cast<float*>(workspace);
  cudaStream_t st = (cudaStream_t)stream;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_wgrad_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
    if (e != cudaSuccess) return fail("k_wgrad_tc smem attr", e);
    attr_set = true;
  }
  int Nk = Kp < TC_NMAX ? Kp : TC_NMAX;
  size_t smem = 2 * TC_A_STAGE + 2 * (size_t)Nk * 128 + 64 + sizeof(int32_t) * B3D_MAX_SEGS * TC_BK;
  dim3 grid((unsigned)S, (unsigned)(ceil_div(Nout, TC_BM) * ktiles));
  k_wgrad_tc<<<grid, TC_THREADS, smem, st>>>(a);
  B3D_LAUNCH_CHECK("k_wgrad_tc");
  long long tot = (long long)Nout * (K + 1);
  k_wgrad_tc_reduce<<<(unsigned)ceil_div(tot, 256), 256, 0, st>>>(a.part, S, Nout, K, dW, lddw, db,
                                                                  (flags & B3D_FLAG_ACCUMULATE) ? 1 : 0);
  B3D_LAUNCH_CHECK("k_wgrad_tc_reduce");
  return 0;
}
