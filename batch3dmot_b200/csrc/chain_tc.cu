// Fused MLP chains on tcgen05: a whole nn.Sequential(Linear, ReLU, ..., Linear) — or the chain of input-gradient
// GEMMs of its backward pass — over one 128-row tile of edges per persistent CTA, with every hidden activation
// kept in shared memory (the epilogue of layer l writes the bf16 tile straight into the 128-byte-swizzled K-major
// image that layer l+1's tcgen05.mma reads as its A operand). Replaces the per-layer launches of k_linear_tma for
//   edge_update            cat[e, att] -> 256 -> 128 -> 64           (clr_att_gnn.py:196-201, :314-317)
//   + both message MLPs'   first layers 64 -> 192 (+ReLU)            (clr_att_gnn.py:203-213, :319-327)
//   att_edge_encoder       64 (+ node-side addends) -> 512 -> 384 ... (clr_att_gnn.py:82-91, :164)
// and of their backward chains dZ_{l-1} = (dZ_l W_l) * relu'(z_{l-1}).
//
// Roles (one CTA per SM, 320 threads):
//   warp 0    producer: TMA box loads of the input tile (128B swizzle) and cp.async.bulk copies of the weight
//             chunks, which are PRE-SWIZZLED in global memory (b3d_chain_pack_weights) and stream from L2 through
//             a ring of shared-memory stages in the fixed order (layer, column block, K chunk);
//   warp 1    MMA issuer (one thread): walks the same program; a K chunk of layer l+1 may issue as soon as the
//             column block of layer l that holds those 64 columns has been written (per-block mbarriers), so
//             the epilogue of one block overlaps the MMAs of the next;
//   warps 2-9 epilogue: TMEM -> registers -> bias / row-gathered node-side addends / ReLU or sign-bit mask ->
//             bf16 -> shared memory (next layer's operand) and, where asked, global memory (+ sign bits).
// Two 256-column TMEM accumulators alternate between consecutive column blocks.
// Safety of buffer reuse rests on two facts: tcgen05.mma complete in issue order, and an epilogue starts only
// after the commit that follows its block's MMAs — so a buffer last read by layer c may be overwritten by the
// epilogue of any layer d > c (host-side planner, chain_plan()).
#include <cuda.h>

#include "b3d_common.cuh"
#include "tc_common.cuh"

namespace b3d {

constexpr int CH_MAX_LAYERS = B3D_CHAIN_MAX_LAYERS;
constexpr int CH_EPI_WARPS = 16;              // 4 per TMEM lane quarter
constexpr int CH_THREADS = 64 + 32 * CH_EPI_WARPS;
constexpr int CH_BM = 128;
constexpr int CH_SUB = CH_BM * 128;     // one [128 rows x 64 cols] bf16 sub-tile: 16 KB
constexpr int CH_MAX_STAGES = 8;
constexpr int CH_SMEM_LIMIT = 227 * 1024;
// barrier block (bytes from its base)
constexpr int CH_ACT_PER_LAYER = 8;             // one "columns written" barrier per 64 output columns (N <= 512)
constexpr int BAR_IN_FULL = 0, BAR_IN_EMPTY = 8, BAR_ACC_FULL = 16, BAR_ACC_EMPTY = 32, BAR_W_FULL = 48,
              BAR_W_EMPTY = 48 + 8 * CH_MAX_STAGES, BAR_ACT = 48 + 16 * CH_MAX_STAGES,
              BAR_TMEM = BAR_ACT + 8 * CH_ACT_PER_LAYER * CH_MAX_LAYERS, BAR_BYTES = 1024;

struct ChainLayerDev {
  int K, N, nblk, Nb, kchunks;
  int src_off, src_layer, dst_off;     // byte offsets inside the activation arena (dst_off < 0: not kept)
  int act, nadd;
  int add_sel[2], add_ld[2];
  const __nv_bfloat16* add_ptr[2];
  const float* bias;
  __nv_bfloat16* out;
  int ldo, bias_off;
  uint32_t* bits_out;
  const uint32_t* bits_in;
  long long w_off;
};

struct ChainArgs {
  ChainLayerDev L[CH_MAX_LAYERS];
  int nl, in_chunks0, in_chunks, in_release;
  int stage_bytes, nstages, arena_bytes, bias_total;
  long long M, ntiles;
  const int32_t* idx0;
  const int32_t* idx1;
  const uint8_t* W;
};

__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// tcgen05.ld of 16 consecutive fp32 columns of this warp's 32 lanes.
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// Per-row state of one (layer, column block): everything the 16-column steps need, resolved ONCE per block so that
// the steps themselves hold no parameter-space loads or per-layer branches beyond the mode bits.
struct ChainRow {
  const __nv_bfloat16* add0;   // row of addend 0 (already gathered through its index), or null
  const __nv_bfloat16* add1;
  __nv_bfloat16* out;          // output row or null
  uint16_t* bits_out;          // this row's sign-bit words viewed as half-words: column c -> [(c / 32) * 2M + (c / 16) % 2]
  const uint16_t* bits_in;
  uint32_t smem_row;           // shared-memory address of this row in the destination tile, 0 = not kept
  const float* bias;           // shared-memory bias of the layer
  int act;
};

// Global operands of one 16-column step (32 bytes per addend), requested one step ahead of their use.
struct ChainPf {
  uint4 a0[2], a1[2];
  uint32_t bits;
};

__device__ __forceinline__ void chain_prefetch(const ChainRow& R, int c, long long M, ChainPf& pf) {
  if (R.add0) ldg256(R.add0 + c, pf.a0[0], pf.a0[1]);
  if (R.add1) ldg256(R.add1 + c, pf.a1[0], pf.a1[1]);
  if (R.bits_in) pf.bits = __ldg(R.bits_in + (long long)(c >> 5) * M * 2 + ((c >> 4) & 1));
}

__device__ __forceinline__ void add_bf16x8(float* o, const uint4& v) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    o[2 * j] += __uint_as_float(w[j] << 16);
    o[2 * j + 1] += __uint_as_float(w[j] & 0xFFFF0000u);
  }
}

// One 16-column step of one row: c = first column of the step in the layer output.
__device__ __forceinline__ void chain_step(const ChainRow& R, int c, long long M, bool ok, const uint32_t (&r)[16],
                                           const ChainPf& pf, int lrow) {
  using namespace tc;
  float o[16];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 b4 = *reinterpret_cast<const float4*>(R.bias + c + 4 * q);
    o[4 * q + 0] = __uint_as_float(r[4 * q + 0]) + b4.x;
    o[4 * q + 1] = __uint_as_float(r[4 * q + 1]) + b4.y;
    o[4 * q + 2] = __uint_as_float(r[4 * q + 2]) + b4.z;
    o[4 * q + 3] = __uint_as_float(r[4 * q + 3]) + b4.w;
  }
  if (R.add0) { add_bf16x8(o, pf.a0[0]); add_bf16x8(o + 8, pf.a0[1]); }
  if (R.add1) { add_bf16x8(o, pf.a1[0]); add_bf16x8(o + 8, pf.a1[1]); }
  if (R.act == B3D_ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 16; ++j) o[j] = fmaxf(o[j], 0.f);
  } else if (R.act == B3D_ACT_MASKBITS) {
    const uint32_t word = R.bits_in ? pf.bits : 0u;
#pragma unroll
    for (int j = 0; j < 16; ++j) o[j] = ((word >> j) & 1u) ? o[j] : 0.f;
  }
  if (R.bits_out) {
    uint32_t word = 0u;
#pragma unroll
    for (int j = 0; j < 16; ++j) word |= (o[j] > 0.f) ? (1u << j) : 0u;
    R.bits_out[(long long)(c >> 5) * M * 2 + ((c >> 4) & 1)] = (uint16_t)word;
  }
  uint4 p0 = make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
  uint4 p1 = make_uint4(pack_bf16x2(o[8], o[9]), pack_bf16x2(o[10], o[11]), pack_bf16x2(o[12], o[13]), pack_bf16x2(o[14], o[15]));
  if (R.out) stg256(R.out + c, p0, p1);
  if (R.smem_row) {
    // K-major, 128B swizzle: 16-byte chunk ch of row lrow sits at chunk position ch ^ (lrow & 7)
    const uint32_t base = R.smem_row + (uint32_t)(c >> 6) * CH_SUB;
    const int ch0 = (c & 63) >> 3;
    st_shared_v4(base + (uint32_t)((ch0 ^ (lrow & 7)) << 4), p0.x, p0.y, p0.z, p0.w);
    st_shared_v4(base + (uint32_t)(((ch0 + 1) ^ (lrow & 7)) << 4), p1.x, p1.y, p1.z, p1.w);
  }
  (void)ok;
}

__global__ void __launch_bounds__(CH_THREADS, 1)
k_chain(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
        const __grid_constant__ ChainArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  using namespace tc;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t pad = (1024u - (smem_u32(smem) & 1023u)) & 1023u;
  const uint32_t sArena = smem_u32(smem) + pad;
  const uint32_t sRing = sArena + (uint32_t)a.arena_bytes;
  const uint32_t misc = (uint32_t)a.arena_bytes + (uint32_t)(a.nstages * a.stage_bytes);
  const uint32_t sBar = sArena + misc;
  volatile uint32_t* s_tmem = reinterpret_cast<volatile uint32_t*>(smem + pad + misc + BAR_TMEM);
  float* s_bias = reinterpret_cast<float*>(smem + pad + misc + BAR_BYTES);
  const int S = a.nstages;

  if (warp == 1) tmem_alloc(sBar + BAR_TMEM, 512);
  if (tid == 0) {
    mbar_init(sBar + BAR_IN_FULL, 1);
    mbar_init(sBar + BAR_IN_EMPTY, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(sBar + BAR_ACC_FULL + 8 * s, 1); mbar_init(sBar + BAR_ACC_EMPTY + 8 * s, CH_EPI_WARPS); }
    for (int s = 0; s < CH_MAX_STAGES; ++s) { mbar_init(sBar + BAR_W_FULL + 8 * s, 1); mbar_init(sBar + BAR_W_EMPTY + 8 * s, 1); }
    for (int s = 0; s < CH_ACT_PER_LAYER * CH_MAX_LAYERS; ++s) mbar_init(sBar + BAR_ACT + 8 * s, CH_EPI_WARPS);
    fence_mbar_init();
  }
  for (int l = 0; l < a.nl; ++l)
    for (int i = tid; i < a.L[l].nblk * a.L[l].Nb; i += CH_THREADS)
      s_bias[a.L[l].bias_off + i] = (a.L[l].bias && i < a.L[l].N) ? __ldg(a.L[l].bias + i) : 0.f;
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *s_tmem;

  if (warp == 0) {
    // ------------------------------------------------------------------ producer
    if (lane == 0) {
      int it = 0, tcount = 0;
      for (long long tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++tcount) {
        if (tcount > 0) mbar_wait(sBar + BAR_IN_EMPTY, (tcount - 1) & 1);
        mbar_expect_tx(sBar + BAR_IN_FULL, (uint32_t)a.in_chunks * CH_SUB);
        for (int c = 0; c < a.in_chunks; ++c) {
          if (c < a.in_chunks0) tma_load_2d(sArena + c * CH_SUB, &mapA0, c * 64, (int)(tile * CH_BM), sBar + BAR_IN_FULL);
          else tma_load_2d(sArena + c * CH_SUB, &mapA1, (c - a.in_chunks0) * 64, (int)(tile * CH_BM), sBar + BAR_IN_FULL);
        }
        for (int l = 0; l < a.nl; ++l) {
          const ChainLayerDev& L = a.L[l];
          const uint32_t bytes = (uint32_t)L.Nb * 128u;
          const uint8_t* src = a.W + L.w_off;
          for (int q = 0; q < L.nblk * L.kchunks; ++q, ++it) {
            const int s = it % S;
            if (it >= S) mbar_wait(sBar + BAR_W_EMPTY + 8 * s, ((it / S) - 1) & 1);
            mbar_expect_tx(sBar + BAR_W_FULL + 8 * s, bytes);
            bulk_copy_g2s(sRing + (uint32_t)(s * a.stage_bytes), src + (size_t)q * bytes, bytes, sBar + BAR_W_FULL + 8 * s);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      int it = 0, tcount = 0, use = 0;
      for (long long tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++tcount) {
        for (int l = 0; l < a.nl; ++l) {
          const ChainLayerDev& L = a.L[l];
          for (int nb = 0; nb < L.nblk; ++nb, ++use) {
            const int slot = use & 1;
            if (use >= 2) mbar_wait(sBar + BAR_ACC_EMPTY + 8 * slot, ((use >> 1) - 1) & 1);
            tc_fence_after_sync();
            const int ncol = min(L.Nb, L.N - nb * L.Nb);
            const uint32_t idesc = make_idesc_bf16(CH_BM, (uint32_t)ncol, 0, 0);
            const uint32_t d_tmem = tmem + (uint32_t)(slot * 256);
            for (int kc = 0; kc < L.kchunks; ++kc, ++it) {
              if (L.src_layer < 0) {
                if (kc == 0) mbar_wait(sBar + BAR_IN_FULL, tcount & 1);
              } else {
                // the 64 columns of the producing layer that make up this K chunk have been written
                mbar_wait(sBar + BAR_ACT + 8 * (CH_ACT_PER_LAYER * L.src_layer + kc), tcount & 1);
              }
              const int s = it % S;
              mbar_wait(sBar + BAR_W_FULL + 8 * s, (it / S) & 1);
              tc_fence_after_sync();
              const int ksteps = min(4, (L.K - kc * 64) >> 4);
              const uint32_t a_base = sArena + (uint32_t)L.src_off + (uint32_t)kc * CH_SUB;
              const uint32_t b_base = sRing + (uint32_t)(s * a.stage_bytes);
              for (int j = 0; j < ksteps; ++j)
                mma_bf16_ss(d_tmem, make_smem_desc_sw128(a_base + j * 32), make_smem_desc_sw128(b_base + j * 32), idesc,
                            (kc | j) != 0);
              mma_commit(sBar + BAR_W_EMPTY + 8 * s);
            }
            mma_commit(sBar + BAR_ACC_FULL + 8 * slot);
            if (l == a.in_release && nb == L.nblk - 1) mma_commit(sBar + BAR_IN_EMPTY);
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue
    // 16 warps: warp w may read TMEM lanes 32 * (w % 4) ..; the four warps of a lane quarter split every 64-column
    // chunk of an accumulator into 16-column steps (sub-warp s takes columns 16 s .. 16 s + 15), so each chunk is
    // finished by all 16 warps together and its "written" barrier fires as early as possible: the next layer's
    // MMAs over that K chunk start while the following chunk is still in the epilogue. Four warps per scheduler
    // cover each other's TMEM-load, shared-memory and global latencies (two per scheduler ran at 27 % issue
    // utilisation: profiles/r2_chain_v1.md).
    const int lq = warp & 3;
    const int sub = (warp - 2) >> 2;
    const int lrow = lq * 32 + lane;
    const uint32_t t_lane = tmem + ((uint32_t)(lq * 32) << 16);
    int tcount = 0, use = 0;
    ChainPf pfa, pfb;
    long long tile = blockIdx.x;
    long long row = tile * CH_BM + lrow;
    bool ok = tile < a.ntiles && row < a.M;
    int i0 = (ok && a.idx0) ? __ldg(a.idx0 + row) : 0, i1 = (ok && a.idx1) ? __ldg(a.idx1 + row) : 0;
    for (; tile < a.ntiles; tile += gridDim.x, ++tcount) {
      const long long ntile = tile + gridDim.x, nrow = ntile * CH_BM + lrow;
      const bool nok = ntile < a.ntiles && nrow < a.M;
      const int n0 = (nok && a.idx0) ? __ldg(a.idx0 + nrow) : 0, n1 = (nok && a.idx1) ? __ldg(a.idx1 + nrow) : 0;
      for (int l = 0; l < a.nl; ++l) {
        const ChainLayerDev& L = a.L[l];
        ChainRow R;
        R.add0 = R.add1 = nullptr;
        if (ok && L.nadd > 0) {
          const long long g = L.add_sel[0] == 0 ? i0 : (L.add_sel[0] == 1 ? i1 : row);
          R.add0 = L.add_ptr[0] + g * L.add_ld[0];
        }
        if (ok && L.nadd > 1) {
          const long long g = L.add_sel[1] == 0 ? i0 : (L.add_sel[1] == 1 ? i1 : row);
          R.add1 = L.add_ptr[1] + g * L.add_ld[1];
        }
        R.out = (ok && L.out) ? L.out + row * L.ldo : nullptr;
        R.bits_out = (ok && L.bits_out) ? reinterpret_cast<uint16_t*>(L.bits_out) + row * 2 : nullptr;
        R.bits_in = (ok && L.act == B3D_ACT_MASKBITS) ? reinterpret_cast<const uint16_t*>(L.bits_in) + row * 2 : nullptr;
        R.smem_row = L.dst_off >= 0 ? sArena + (uint32_t)L.dst_off + (uint32_t)lrow * 128u : 0u;
        R.bias = s_bias + L.bias_off;
        R.act = L.act;
        const int nblk = L.nblk, Nb = L.Nb, N = L.N;
        for (int nb = 0; nb < nblk; ++nb, ++use) {
          const int slot = use & 1;
          const int ncol = min(Nb, N - nb * Nb);
          const int c0 = nb * Nb + sub * 16;
          chain_prefetch(R, c0, a.M, pfa);                  // first step of the block: in flight during the wait
          mbar_wait(sBar + BAR_ACC_FULL + 8 * slot, (use >> 1) & 1);
          tc_fence_after_sync();
          const uint32_t t_acc = t_lane + (uint32_t)(slot * 256 + sub * 16);
          for (int ch = 0; ch < ncol; ch += 128) {        // two 64-column chunks per iteration (static prefetch roles)
            uint32_t r[16];
            tmem_ld16(t_acc + (uint32_t)ch, r);
            if (ch + 64 < ncol) chain_prefetch(R, c0 + ch + 64, a.M, pfb);
            tmem_ld_wait();
            chain_step(R, c0 + ch, a.M, ok, r, pfa, lrow);
            if (R.smem_row) fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(sBar + BAR_ACT + 8 * (CH_ACT_PER_LAYER * l + ((nb * Nb + ch) >> 6)));
            if (ch + 64 < ncol) {
              tmem_ld16(t_acc + (uint32_t)(ch + 64), r);
              if (ch + 128 < ncol) chain_prefetch(R, c0 + ch + 128, a.M, pfa);
              tmem_ld_wait();
              chain_step(R, c0 + ch + 64, a.M, ok, r, pfb, lrow);
              if (R.smem_row) fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) mbar_arrive(sBar + BAR_ACT + 8 * (CH_ACT_PER_LAYER * l + ((nb * Nb + ch + 64) >> 6)));
            }
          }
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(sBar + BAR_ACC_EMPTY + 8 * slot);
        }
      }
      row = nrow; ok = nok; i0 = n0; i1 = n1;
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// Pre-swizzled chunk stream of one layer: for column block nb and K chunk kc a [Nb rows][64 k] bf16 image whose
// 16-byte chunk c of row n sits at position c ^ (n & 7) — the shared-memory image of a K-major 128B-swizzled
// operand, so the producer moves it with one linear bulk copy.
__global__ void k_pack_chain(const float* __restrict__ W, int ldw, int transpose, int K, int N, int Nb, int nblk,
                             int kchunks, uint8_t* __restrict__ dst) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // one 16-byte chunk
  const long long total = (long long)nblk * kchunks * Nb * 8;
  if (i >= total) return;
  const int c = (int)(i & 7);
  const long long q = i >> 3;
  const int n = (int)(q % Nb);
  const long long blk = q / Nb;
  const int kc = (int)(blk % kchunks), nb = (int)(blk / kchunks);
  const int ng = nb * Nb + n, k0 = kc * 64 + c * 8;
  float v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int k = k0 + e;
    v[e] = (ng < N && k < K) ? (transpose ? W[(long long)k * ldw + ng] : W[(long long)ng * ldw + k]) : 0.f;
  }
  uint4 pk = make_uint4(tc::pack_bf16x2(v[0], v[1]), tc::pack_bf16x2(v[2], v[3]), tc::pack_bf16x2(v[4], v[5]),
                        tc::pack_bf16x2(v[6], v[7]));
  *reinterpret_cast<uint4*>(dst + (blk * Nb + n) * 128 + ((c ^ (n & 7)) << 4)) = pk;
}

// ------------------------------------------------------------------ host-side plan
struct ChainPlan {
  int nblk[CH_MAX_LAYERS], Nb[CH_MAX_LAYERS], kchunks[CH_MAX_LAYERS], dst_off[CH_MAX_LAYERS], src_off[CH_MAX_LAYERS];
  int bias_off[CH_MAX_LAYERS];
  long long w_off[CH_MAX_LAYERS], w_total;
  int in_chunks, in_release, stage_bytes, nstages, arena_bytes, bias_total;
  size_t smem;
};

static int chain_plan_cap(const b3d_chain_layer_t* Ls, int nl, int k_in, int cap, ChainPlan* P) {
  if (nl < 1 || nl > CH_MAX_LAYERS) return -1;
  P->in_chunks = (k_in + 63) / 64;
  long long woff = 0;
  int bias = 0, stage = 0;
  int last_use[CH_MAX_LAYERS + 1];     // index 0: the input, 1 + l: output of layer l
  for (int i = 0; i <= nl; ++i) last_use[i] = -2;
  for (int l = 0; l < nl; ++l) {
    const b3d_chain_layer_t& L = Ls[l];
    if (L.N < 64 || L.N > 512 || (L.N & 63) || L.K < 16 || (L.K & 15) || L.src >= l || L.src < -1) return -1;
    if (L.K != (L.src < 0 ? k_in : Ls[L.src].N)) return -1;
    int nblk = (L.N + cap - 1) / cap;
    int Nb = ((L.N + nblk - 1) / nblk + 63) / 64 * 64;
    if (nblk > 2) return -1;
    P->nblk[l] = nblk; P->Nb[l] = Nb; P->kchunks[l] = (L.K + 63) / 64;
    P->w_off[l] = woff;
    woff += (long long)nblk * P->kchunks[l] * Nb * 128;
    P->bias_off[l] = bias;
    bias += nblk * Nb;
    if (Nb * 128 > stage) stage = Nb * 128;
    last_use[L.src + 1] = l;
  }
  P->w_total = woff; P->bias_total = bias; P->stage_bytes = stage;
  // activation arena: first fit; a buffer defined by layer d may reuse the bytes of one last read by layer c < d
  int off[CH_MAX_LAYERS + 1], size[CH_MAX_LAYERS + 1], def[CH_MAX_LAYERS + 1];
  int arena = 0;
  for (int b = 0; b <= nl; ++b) {
    def[b] = b - 1;
    size[b] = b == 0 ? P->in_chunks * CH_SUB : (last_use[b] >= 0 ? Ls[b - 1].N / 64 * CH_SUB : 0);
    off[b] = -1;
    if (size[b] == 0) continue;
    int cand = 0;
    for (bool moved = true; moved;) {
      moved = false;
      for (int o = 0; o < b; ++o) {
        if (off[o] < 0 || def[b] > last_use[o]) continue;                  // dead by then: may overlap
        if (cand < off[o] + size[o] && off[o] < cand + size[b]) { cand = off[o] + size[o]; moved = true; }
      }
    }
    off[b] = cand;
    if (cand + size[b] > arena) arena = cand + size[b];
  }
  P->arena_bytes = arena;
  P->in_release = last_use[0];
  for (int b = 1; b <= nl; ++b)
    if (off[b] >= 0 && off[b] < off[0] + size[0] && off[0] < off[b] + size[b] && last_use[b] > P->in_release)
      P->in_release = last_use[b];
  for (int l = 0; l < nl; ++l) {
    P->dst_off[l] = off[l + 1];
    P->src_off[l] = off[Ls[l].src + 1];
  }
  const int fixed = arena + BAR_BYTES + bias * 4 + 1024;
  int ns = (CH_SMEM_LIMIT - fixed) / stage;
  if (ns > CH_MAX_STAGES) ns = CH_MAX_STAGES;
  P->nstages = ns;
  P->smem = (size_t)fixed + (size_t)(ns > 0 ? ns : 0) * stage;
  return ns >= 2 ? 0 : -1;
}

static int chain_plan(const b3d_chain_layer_t* Ls, int nl, int k_in, ChainPlan* P) {
  // prefer 256-wide column blocks; fall back to narrower ones when that buys a deeper weight ring
  ChainPlan best;
  int have = 0;
  for (int cap : {256, 192, 128}) {
    ChainPlan p;
    if (chain_plan_cap(Ls, nl, k_in, cap, &p)) continue;
    if (!have || (best.nstages < 3 && (long long)p.nstages * p.stage_bytes > (long long)best.nstages * best.stage_bytes)) {
      best = p;
      have = 1;
    }
    if (best.nstages >= 3) break;
  }
  if (!have) return -1;
  *P = best;
  return 0;
}

}  // namespace b3d

using namespace b3d;

extern "C" int b3d_chain_supported(const b3d_chain_layer_t* layers, int32_t nl, int32_t k_in) {
  ChainPlan P;
  return layers && chain_plan(layers, nl, k_in, &P) == 0 ? 1 : 0;
}

extern "C" size_t b3d_chain_packed_bytes(const b3d_chain_layer_t* layers, int32_t nl, int32_t k_in) {
  ChainPlan P;
  if (!layers || chain_plan(layers, nl, k_in, &P)) return 0;
  return (size_t)P.w_total;
}

extern "C" int b3d_chain_pack_weights(const b3d_chain_layer_t* layers, int32_t nl, int32_t k_in, int32_t layer,
                                      const float* W, int32_t ldw, int32_t transpose, void* packed, void* stream) {
  ChainPlan P;
  if (!layers || !W || !packed || layer < 0 || layer >= nl || chain_plan(layers, nl, k_in, &P))
    return bad_arg("b3d_chain_pack_weights");
  const long long total = (long long)P.nblk[layer] * P.kchunks[layer] * P.Nb[layer] * 8;
  k_pack_chain<<<(unsigned)ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(
      W, ldw, transpose, layers[layer].K, layers[layer].N, P.Nb[layer], P.nblk[layer], P.kchunks[layer],
      reinterpret_cast<uint8_t*>(packed) + P.w_off[layer]);
  B3D_LAUNCH_CHECK("k_pack_chain");
  return 0;
}

extern "C" int b3d_chain_run(const b3d_seg_t* in_segs, int32_t nseg, const b3d_chain_layer_t* layers, int32_t nl,
                             const void* packed, const int32_t* idx0, const int32_t* idx1, int64_t M, void* stream) {
  if (M == 0) return 0;
  SegDev seg[2];
  if (nseg < 1 || nseg > 2 || to_dev(in_segs, nseg, seg)) return bad_arg("b3d_chain_run: 1 or 2 input segments");
  int k_in = 0;
  for (int s = 0; s < nseg; ++s) {
    if (seg[s].dtype != B3D_BF16 || seg[s].idx || seg[s].mask_mode != B3D_MASK_NONE || (seg[s].width & 15) ||
        (seg[s].ld & 7) || (reinterpret_cast<uintptr_t>(seg[s].ptr) & 15))
      return bad_arg("b3d_chain_run: input segments must be dense bf16, widths % 16, 16-byte aligned rows");
    if (s + 1 < nseg && (seg[s].width & 63)) return bad_arg("b3d_chain_run: leading segment width % 64");
    k_in += seg[s].width;
  }
  ChainPlan P;
  if (!layers || !packed || M < 0 || chain_plan(layers, nl, k_in, &P)) return bad_arg("b3d_chain_run: unsupported chain");
  static ChainArgs a;       // ~1 KB; filled per call (single host thread per process drives the library)
  a.nl = nl;
  for (int l = 0; l < nl; ++l) {
    const b3d_chain_layer_t& H = layers[l];
    ChainLayerDev& D = a.L[l];
    D.K = H.K; D.N = H.N; D.nblk = P.nblk[l]; D.Nb = P.Nb[l]; D.kchunks = P.kchunks[l];
    D.src_off = P.src_off[l]; D.src_layer = H.src; D.dst_off = P.dst_off[l];
    D.act = H.act; D.nadd = H.nadd;
    if (H.nadd < 0 || H.nadd > 2) return bad_arg("b3d_chain_run: nadd");
    for (int t = 0; t < 2; ++t) {
      D.add_sel[t] = H.add_idx[t]; D.add_ld[t] = H.add_ld[t];
      D.add_ptr[t] = reinterpret_cast<const __nv_bfloat16*>(H.add_ptr[t]);
      if (t < H.nadd) {
        if (!H.add_ptr[t] || (H.add_ld[t] & 15) || (reinterpret_cast<uintptr_t>(H.add_ptr[t]) & 31) || H.add_idx[t] < -1 ||
            H.add_idx[t] > 1 || (H.add_idx[t] == 0 && !idx0) || (H.add_idx[t] == 1 && !idx1))
          return bad_arg("b3d_chain_run: addends must be bf16 [*, N] with 32-byte aligned rows and a valid index selector");
      }
    }
    D.bias = H.bias;
    D.out = reinterpret_cast<__nv_bfloat16*>(H.out); D.ldo = H.ldo; D.bias_off = P.bias_off[l];
    if (H.out && ((H.ldo & 15) || (reinterpret_cast<uintptr_t>(H.out) & 31) || H.ldo < H.N))
      return bad_arg("b3d_chain_run: outputs must be bf16 with 32-byte aligned rows (ld % 16 == 0)");
    D.bits_out = reinterpret_cast<uint32_t*>(H.bits_out);
    D.bits_in = reinterpret_cast<const uint32_t*>(H.bits_in);
    if (H.act == B3D_ACT_MASKBITS && !H.bits_in) return bad_arg("b3d_chain_run: B3D_ACT_MASKBITS needs bits_in");
    if (H.act != B3D_ACT_NONE && H.act != B3D_ACT_RELU && H.act != B3D_ACT_MASKBITS) return bad_arg("b3d_chain_run: act");
    D.w_off = P.w_off[l];
  }
  a.in_chunks = P.in_chunks;
  a.in_chunks0 = nseg == 2 ? seg[0].width / 64 : P.in_chunks;
  a.in_release = P.in_release;
  a.stage_bytes = P.stage_bytes; a.nstages = P.nstages; a.arena_bytes = P.arena_bytes; a.bias_total = P.bias_total;
  a.M = M; a.ntiles = ceil_div(M, CH_BM);
  a.idx0 = idx0; a.idx1 = idx1;
  a.W = reinterpret_cast<const uint8_t*>(packed);
  alignas(64) CUtensorMap mA0, mA1;
  if (make_tmap_bf16(&mA0, seg[0].ptr, M, seg[0].width, seg[0].ld, CH_BM)) return bad_arg("b3d_chain_run: tensor map A0");
  if (nseg == 2) {
    if (make_tmap_bf16(&mA1, seg[1].ptr, M, seg[1].width, seg[1].ld, CH_BM)) return bad_arg("b3d_chain_run: tensor map A1");
  } else {
    mA1 = mA0;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, CH_SMEM_LIMIT);
    if (e != cudaSuccess) return fail("k_chain smem attr", e);
    attr_set = true;
  }
  static int n_sm = 0;
  if (!n_sm) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    if (n_sm <= 0) n_sm = 148;
  }
  long long gx = n_sm < a.ntiles ? n_sm : a.ntiles;
  k_chain<<<(unsigned)gx, CH_THREADS, P.smem, (cudaStream_t)stream>>>(mA0, mA1, a);
  B3D_LAUNCH_CHECK("k_chain");
  return 0;
}
