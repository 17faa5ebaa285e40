// fp32 "exact mode" dense layers with fused gather / concat / bias / activation / masks:
// forward + input-gradient GEMM (b3d_linear) and deterministic weight-gradient GEMM
// (b3d_wgrad). FFMA register-tiled SIMT kernels: this is the 1e-4-tolerance path; the
// bf16 tcgen05 tiles live in fused_tc.cu.
#include "b3d_common.cuh"

namespace b3d {

// ------------------------------------------------------------------ forward / dgrad GEMM
constexpr int BM = 128, BK = 16, LIN_THREADS = 256;   // BN = 16 * TN columns per CTA (TN = 1, 2, 4)

struct LinArgs {
  SegDev seg[B3D_MAX_SEGS];
  int nseg;
  const float* W;
  int ldw, trans_w;
  const float* bias;
  float* Y;
  int ldy;
  long long M;
  int Nout, act, flags;
  const float* out_mask;
  int ldm;
  const uint8_t* row_mask;
  SegDev add[2];
  int nadd;
};

template <int TN>
__global__ void __launch_bounds__(LIN_THREADS) k_linear(const LinArgs a) {
  constexpr int BN = 16 * TN;
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int ty = tid >> 4, tx = tid & 15;
  // loader roles
  const int ar = tid & (BM - 1), akp = tid >> 7;   // A: row, 8-wide k part
  const int bn = tid & (BN - 1), bkp = tid / BN;   // B: col, TN-wide k part
  const long long arow = m0 + ar;
  const bool arow_ok = arow < a.M;
  const bool bcol_ok = (n0 + bn) < a.Nout;

  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float a_reg[8], b_reg[TN];
  int seg = 0, k0 = 0, koff = 0;  // current chunk: segment, offset in segment, column offset of segment in W
  long long grow = 0;             // gathered row for the current segment
  auto seg_row = [&](int s) { grow = arow_ok ? (a.seg[s].idx ? (long long)a.seg[s].idx[arow] : arow) : 0; };
  seg_row(0);

  auto load_chunk = [&]() {
    const SegDev& sg = a.seg[seg];
    const int kw = min(BK, sg.width - k0);
    // ---- A
    {
      const int kb = akp * 8;
      const float* p = sg.ptr + grow * sg.ld + k0 + kb;
      const bool vec = ((sg.ld & 3) == 0) && ((k0 & 3) == 0) && ((kw & 3) == 0) &&
                       ((reinterpret_cast<uintptr_t>(sg.ptr) & 15) == 0);
      if (!arow_ok) {
#pragma unroll
        for (int j = 0; j < 8; ++j) a_reg[j] = 0.f;
      } else if (vec) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (kb + 4 * h < kw) v = __ldg(reinterpret_cast<const float4*>(p + 4 * h));
          a_reg[4 * h + 0] = v.x; a_reg[4 * h + 1] = v.y; a_reg[4 * h + 2] = v.z; a_reg[4 * h + 3] = v.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) a_reg[j] = (kb + j < kw) ? __ldg(p + j) : 0.f;
      }
      if (sg.mask_mode != B3D_MASK_NONE && arow_ok) {
        const float* mp = sg.mask + grow * sg.ldmask + k0 + kb;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (kb + j < kw) a_reg[j] = apply_mask(a_reg[j], __ldg(mp + j), sg.mask_mode);
      }
    }
    // ---- B (weights): Bs[k][n]
    {
      const int kb = bkp * TN;
      if (!bcol_ok) {
#pragma unroll
        for (int j = 0; j < TN; ++j) b_reg[j] = 0.f;
      } else if (!a.trans_w) {
        const float* p = a.W + (long long)(n0 + bn) * a.ldw + koff + k0 + kb;
        const bool vec = (TN == 4) && ((a.ldw & 3) == 0) && (((koff + k0) & 3) == 0) && (kb + 3 < kw) &&
                         ((reinterpret_cast<uintptr_t>(a.W) & 15) == 0);
        if (vec) {
          float4 v = __ldg(reinterpret_cast<const float4*>(p));
          b_reg[0] = v.x; b_reg[TN > 1 ? 1 : 0] = v.y; b_reg[TN > 2 ? 2 : 0] = v.z; b_reg[TN > 3 ? 3 : 0] = v.w;
        } else {
#pragma unroll
          for (int j = 0; j < TN; ++j) b_reg[j] = (kb + j < kw) ? __ldg(p + j) : 0.f;
        }
      } else {
        const float* p = a.W + (long long)(koff + k0 + kb) * a.ldw + n0 + bn;
#pragma unroll
        for (int j = 0; j < TN; ++j) b_reg[j] = (kb + j < kw) ? __ldg(p + (long long)j * a.ldw) : 0.f;
      }
    }
  };
  auto advance = [&]() -> bool {  // move to next chunk; false when done
    k0 += BK;
    if (k0 >= a.seg[seg].width) {
      koff += a.seg[seg].width;
      k0 = 0;
      ++seg;
      if (seg >= a.nseg) return false;
      seg_row(seg);
    }
    return true;
  };

  load_chunk();
  bool more = true;
  while (more) {
#pragma unroll
    for (int j = 0; j < 8; ++j) As[akp * 8 + j][ar] = a_reg[j];
#pragma unroll
    for (int j = 0; j < TN; ++j) Bs[bkp * TN + j][bn] = b_reg[j];
    __syncthreads();
    more = advance();
    if (more) load_chunk();  // global loads overlap the FMAs below
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
      float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bv[TN];
#pragma unroll
      for (int j = 0; j < TN; ++j) bv[j] = Bs[k][tx * TN + j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

  // ---- epilogue
  float bv[TN];
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    int c = n0 + tx * TN + j;
    bv[j] = (a.bias && c < a.Nout) ? __ldg(a.bias + c) : 0.f;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    long long r = m0 + ty * 8 + i;
    if (r >= a.M) continue;
    const bool rz = a.row_mask && a.row_mask[r] == 0;
    const float* addp[2] = {nullptr, nullptr};
    for (int q = 0; q < a.nadd; ++q)
      addp[q] = a.add[q].ptr + (long long)(a.add[q].idx ? __ldg(a.add[q].idx + r) : r) * a.add[q].ld;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int c = n0 + tx * TN + j;
      if (c >= a.Nout) continue;
      float v = acc[i][j] + bv[j];
      if (addp[0]) v += __ldg(addp[0] + c);
      if (addp[1]) v += __ldg(addp[1] + c);
      if (a.act == B3D_ACT_RELU) v = fmaxf(v, 0.f);
      else if (a.act == B3D_ACT_SIGMOID) v = 1.f / (1.f + expf(-v));
      if (a.out_mask) v = (__ldg(a.out_mask + r * a.ldm + c) > 0.f) ? v : 0.f;
      if (rz) v = 0.f;
      float* yp = a.Y + r * a.ldy + c;
      if (a.flags & B3D_FLAG_ACCUMULATE) v += *yp;
      *yp = v;
    }
  }
}

// ------------------------------------------------------------------ weight gradient
constexpr int WG_T = 64, WG_RPT = 4, WG_BR = 16 * WG_RPT, WG_THREADS = 256;   // 64 rows per smem stage

struct WgArgs {
  SegDev dy;
  SegDev seg[B3D_MAX_SEGS];
  int nseg, K, Nout, ktiles;
  long long M, rows_per_split;
  float* part;   // [S][Nout][K]
  float* pbias;  // [S][Nout]
};

static void wgrad_plan(long long M, int Nout, int K, int* S, long long* rows_per_split) {
  long long tiles = ceil_div(Nout, WG_T) * ceil_div(K, WG_T);
  long long s = (4 * 148 + tiles - 1) / tiles;
  long long smax = ceil_div(M > 0 ? M : 1, 256);
  if (s > smax) s = smax;
  if (s < 1) s = 1;
  long long rps = ceil_div(ceil_div(M > 0 ? M : 1, s), WG_BR) * WG_BR;
  *S = (int)ceil_div(M > 0 ? M : 1, rps);
  *rows_per_split = rps;
}

__global__ void __launch_bounds__(WG_THREADS) k_wgrad(const WgArgs a) {
  __shared__ __align__(16) float Ds[WG_BR][WG_T + 4];
  __shared__ __align__(16) float As[WG_BR][WG_T + 4];
  const int tid = threadIdx.x;
  const int split = blockIdx.x;
  const int nt = blockIdx.y / a.ktiles, kt = blockIdx.y % a.ktiles;
  const int n0 = nt * WG_T, kc0 = kt * WG_T;
  const int ty = tid >> 4, tx = tid & 15;
  const int lr = tid >> 4, lc = (tid & 15) * 4;  // loader: row in stage, first of 4 columns
  const long long r0 = (long long)split * a.rows_per_split;
  const long long r1 = min(a.M, r0 + a.rows_per_split);

  // map my 4 concatenated A columns to (segment, offset)
  int sseg[4], soff[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int c = kc0 + lc + j, s = 0;
    sseg[j] = -1; soff[j] = 0;
    if (c < a.K) {
      while (c >= a.seg[s].width) { c -= a.seg[s].width; ++s; }
      sseg[j] = s; soff[j] = c;
    }
  }
  const bool a_same = sseg[0] >= 0 && sseg[3] == sseg[0] && soff[3] == soff[0] + 3;
  bool a_vec = false;
  if (a_same) {
    const SegDev& sg = a.seg[sseg[0]];
    a_vec = ((sg.ld & 3) == 0) && ((soff[0] & 3) == 0) && ((reinterpret_cast<uintptr_t>(sg.ptr) & 15) == 0) &&
            sg.mask_mode == B3D_MASK_NONE;
  }
  const bool d_vec = (n0 + lc + 3 < a.Nout) && ((a.dy.ld & 3) == 0) &&
                     ((reinterpret_cast<uintptr_t>(a.dy.ptr) & 15) == 0) && a.dy.mask_mode == B3D_MASK_NONE;

  float acc[4][4], accb[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    accb[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  }
  float dreg[WG_RPT][4], areg[WG_RPT][4];
  auto load_stage = [&](long long rbase) {
#pragma unroll
   for (int q = 0; q < WG_RPT; ++q) {
    const long long r = rbase + q * 16 + lr;
    const bool ok = r < r1;
    // dY
    if (!ok) {
#pragma unroll
      for (int j = 0; j < 4; ++j) dreg[q][j] = 0.f;
    } else if (d_vec) {
      float4 v = __ldg(reinterpret_cast<const float4*>(a.dy.ptr + r * a.dy.ld + n0 + lc));
      dreg[q][0] = v.x; dreg[q][1] = v.y; dreg[q][2] = v.z; dreg[q][3] = v.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int n = n0 + lc + j;
        float v = 0.f;
        if (n < a.Nout) {
          v = __ldg(a.dy.ptr + r * a.dy.ld + n);
          if (a.dy.mask_mode != B3D_MASK_NONE) v = apply_mask(v, __ldg(a.dy.mask + r * a.dy.ldmask + n), a.dy.mask_mode);
        }
        dreg[q][j] = v;
      }
    }
    // A
    if (!ok) {
#pragma unroll
      for (int j = 0; j < 4; ++j) areg[q][j] = 0.f;
    } else if (a_vec) {
      const SegDev& sg = a.seg[sseg[0]];
      long long gr = sg.idx ? (long long)__ldg(sg.idx + r) : r;
      float4 v = __ldg(reinterpret_cast<const float4*>(sg.ptr + gr * sg.ld + soff[0]));
      areg[q][0] = v.x; areg[q][1] = v.y; areg[q][2] = v.z; areg[q][3] = v.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float v = 0.f;
        if (sseg[j] >= 0) {
          const SegDev& sg = a.seg[sseg[j]];
          long long gr = sg.idx ? (long long)__ldg(sg.idx + r) : r;
          v = __ldg(sg.ptr + gr * sg.ld + soff[j]);
          if (sg.mask_mode != B3D_MASK_NONE) v = apply_mask(v, __ldg(sg.mask + gr * sg.ldmask + soff[j]), sg.mask_mode);
        }
        areg[q][j] = v;
      }
    }
   }
  };

  if (r0 < r1) load_stage(r0);
  for (long long rb = r0; rb < r1; rb += WG_BR) {
#pragma unroll
    for (int q = 0; q < WG_RPT; ++q) {
      *reinterpret_cast<float4*>(&Ds[q * 16 + lr][lc]) = make_float4(dreg[q][0], dreg[q][1], dreg[q][2], dreg[q][3]);
      *reinterpret_cast<float4*>(&As[q * 16 + lr][lc]) = make_float4(areg[q][0], areg[q][1], areg[q][2], areg[q][3]);
    }
    __syncthreads();
    if (rb + WG_BR < r1) load_stage(rb + WG_BR);
#pragma unroll
    for (int r = 0; r < WG_BR; ++r) {
      float4 d = *reinterpret_cast<const float4*>(&Ds[r][ty * 4]);
      float4 v = *reinterpret_cast<const float4*>(&As[r][tx * 4]);
      float dv[4] = {d.x, d.y, d.z, d.w}, av[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        accb[i] += dv[i];
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(dv[i], av[j], acc[i][j]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int n = n0 + ty * 4 + i;
    if (n >= a.Nout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int k = kc0 + tx * 4 + j;
      if (k < a.K) a.part[((long long)split * a.Nout + n) * a.K + k] = acc[i][j];
    }
    if (kt == 0 && tx == 0) a.pbias[(long long)split * a.Nout + n] = accb[i];
  }
}

__global__ void k_wgrad_reduce(const float* __restrict__ part, const float* __restrict__ pbias, int S,
                               int Nout, int K, float* __restrict__ dW, int lddw, float* __restrict__ db,
                               int accumulate) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long nk = (long long)Nout * K;
  if (i < nk) {
    float s = 0.f;
    for (int p = 0; p < S; ++p) s += part[(long long)p * nk + i];  // fixed order: deterministic
    int n = (int)(i / K), k = (int)(i % K);
    float* o = dW + (long long)n * lddw + k;
    *o = accumulate ? *o + s : s;
  } else if (db && i < nk + Nout) {
    int n = (int)(i - nk);
    float s = 0.f;
    for (int p = 0; p < S; ++p) s += pbias[(long long)p * Nout + n];
    db[n] = accumulate ? db[n] + s : s;
  }
}

}  // namespace b3d

using namespace b3d;

extern "C" int b3d_linear(const b3d_seg_t* segs, int32_t nseg, const float* W, int32_t ldw,
                          int32_t trans_w, const float* bias, float* Y, int32_t ldy, int64_t M,
                          int32_t Nout, int32_t act, int32_t flags, const float* out_mask, int32_t ldm,
                          const uint8_t* row_mask, const b3d_seg_t* adds, int32_t nadd, void* stream) {
  LinArgs a;
  if (M == 0) return 0;
  if (nadd < 0 || nadd > 2 || (nadd && (to_dev(adds, nadd, a.add) || !all_f32(a.add, nadd))))
    return bad_arg("b3d_linear adds");
  for (int q = 0; q < nadd; ++q)
    if (a.add[q].width != Nout) return bad_arg("b3d_linear: adds width != Nout");
  a.nadd = nadd;
  if (to_dev(segs, nseg, a.seg) || !all_f32(a.seg, nseg)) return bad_arg("b3d_linear segments (fp32 only)");
  if (!W || !Y || Nout <= 0 || M < 0) return bad_arg("b3d_linear W/Y/Nout/M");
  a.nseg = nseg; a.W = W; a.ldw = ldw; a.trans_w = trans_w; a.bias = bias; a.Y = Y; a.ldy = ldy;
  a.M = M; a.Nout = Nout; a.act = act; a.flags = flags; a.out_mask = out_mask; a.ldm = ldm;
  a.row_mask = row_mask;
  // narrow outputs (classifier / encoder layers) use narrower column tiles instead of padding to 64
  const int TN = Nout <= 16 ? 1 : (Nout <= 32 ? 2 : 4);
  dim3 grid((unsigned)ceil_div(M, BM), (unsigned)ceil_div(Nout, 16 * TN));
  if (TN == 1) k_linear<1><<<grid, LIN_THREADS, 0, (cudaStream_t)stream>>>(a);
  else if (TN == 2) k_linear<2><<<grid, LIN_THREADS, 0, (cudaStream_t)stream>>>(a);
  else k_linear<4><<<grid, LIN_THREADS, 0, (cudaStream_t)stream>>>(a);
  B3D_LAUNCH_CHECK("k_linear");
  return 0;
}

extern "C" size_t b3d_wgrad_workspace_bytes(int64_t M, int32_t Nout, int32_t K) {
  int S; long long rps;
  wgrad_plan(M, Nout, K, &S, &rps);
  return sizeof(float) * ((size_t)S * Nout * K + (size_t)S * Nout) + 256;
}

extern "C" int b3d_wgrad(const b3d_seg_t* dy, const b3d_seg_t* segs, int32_t nseg, float* dW,
                         int32_t lddw, float* db, int64_t M, int32_t Nout, int32_t flags,
                         void* workspace, size_t workspace_bytes, void* stream) {
  WgArgs a;
  cudaStream_t st = (cudaStream_t)stream;
  int K = 0;
  if (nseg < 1 || nseg > B3D_MAX_SEGS) return bad_arg("b3d_wgrad nseg");
  for (int s = 0; s < nseg; ++s) K += segs[s].width;
  if (M > 0) {
    if (to_dev(segs, nseg, a.seg) || to_dev(dy, 1, &a.dy) || !all_f32(a.seg, nseg) || !all_f32(&a.dy, 1))
      return bad_arg("b3d_wgrad segments (fp32 only)");
    if (a.dy.idx) return bad_arg("b3d_wgrad: dy must not be gathered");
    if (a.dy.width != Nout) return bad_arg("b3d_wgrad: dy.width != Nout");
  }
  if (M <= 0) {
    if (!(flags & B3D_FLAG_ACCUMULATE)) {
      for (int n = 0; n < Nout; ++n) cudaMemsetAsync(dW + (size_t)n * lddw, 0, sizeof(float) * K, st);
      if (db) cudaMemsetAsync(db, 0, sizeof(float) * Nout, st);
    }
    return 0;
  }
  int S; long long rps;
  wgrad_plan(M, Nout, K, &S, &rps);
  if (workspace_bytes < b3d_wgrad_workspace_bytes(M, Nout, K)) return bad_arg("wgrad workspace too small");
  a.nseg = nseg; a.K = K; a.Nout = Nout; a.M = M; a.rows_per_split = rps;
  a.ktiles = (int)ceil_div(K, WG_T);
  a.part = (float*)workspace;
  a.pbias = a.part + (size_t)S * Nout * K;
  dim3 grid((unsigned)S, (unsigned)(ceil_div(Nout, WG_T) * a.ktiles));
  k_wgrad<<<grid, WG_THREADS, 0, st>>>(a);
  B3D_LAUNCH_CHECK("k_wgrad");
  long long tot = (long long)Nout * K + Nout;
  k_wgrad_reduce<<<(unsigned)ceil_div(tot, 256), 256, 0, st>>>(a.part, a.pbias, S, Nout, K, dW, lddw, db,
                                                               (flags & B3D_FLAG_ACCUMULATE) ? 1 : 0);
  B3D_LAUNCH_CHECK("k_wgrad_reduce");
  return 0;
}
