// Experiment: TMA tile::gather4 semantics on sm_100a (tensor-map box shape, shared-memory image with 128B swizzle).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cstdint>

typedef CUresult (*enc_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                           const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                           CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void k(const __grid_constant__ CUtensorMap map, const int* idx, int nrows, int col0, uint8_t* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  uint32_t sb = (uint32_t)__cvta_generic_to_shared(&bar);
  uint32_t base = (uint32_t)__cvta_generic_to_shared(smem);
  base = (base + 1023) & ~1023u;
  uint32_t off = base - (uint32_t)__cvta_generic_to_shared(smem);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sb));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  for (int i = threadIdx.x; i < nrows * 128 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem + off)[i] = 0xDEADBEEF;
  asm volatile("fence.proxy.async.shared::cta;");
  __syncthreads();
  if (threadIdx.x == 0)
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sb), "r"(nrows * 128));
  __syncthreads();
  if (threadIdx.x < nrows / 4) {
    int i = threadIdx.x;
    int4 r = reinterpret_cast<const int4*>(idx)[i];
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
        ::"r"(base + i * 512), "l"(reinterpret_cast<uint64_t>(&map)), "r"(col0), "r"(r.x), "r"(r.y), "r"(r.z), "r"(r.w), "r"(sb)
        : "memory");
  }
  if (threadIdx.x == 0) {
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0,1,0,p;}" : "=r"(ok) : "r"(sb) : "memory");
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nrows * 128; i += blockDim.x) out[i] = smem[off + i];
}

int main(int argc, char** argv) {
  int box_rows = argc > 1 ? atoi(argv[1]) : 1;
  const int R = 1000, C = 128, nrows = 128, col0 = 64;
  std::vector<__nv_bfloat16> h(R * C);
  for (int r = 0; r < R; ++r) for (int c = 0; c < C; ++c) h[r * C + c] = __float2bfloat16((float)(r % 256) + c / 256.0f);
  std::vector<int> idx(nrows);
  for (int i = 0; i < nrows; ++i) idx[i] = (i * 37 + 11) % R;
  __nv_bfloat16* d; int* di; uint8_t* dout;
  cudaMalloc(&d, h.size() * 2); cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  cudaMalloc(&di, nrows * 4); cudaMemcpy(di, idx.data(), nrows * 4, cudaMemcpyHostToDevice);
  cudaMalloc(&dout, nrows * 128);
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  enc_fn enc = (enc_fn)p;
  CUtensorMap m;
  cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)R};
  cuuint64_t strides[1] = {(cuuint64_t)C * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult rc = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode box_rows=%d rc=%d\n", box_rows, (int)rc);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  k<<<1, 128, 32 * 1024>>>(m, di, nrows, col0, dout);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  std::vector<uint8_t> o(nrows * 128);
  cudaMemcpy(o.data(), dout, o.size(), cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int j = 0; j < nrows; ++j)
    for (int ch = 0; ch < 8; ++ch) {
      const __nv_bfloat16* got = reinterpret_cast<const __nv_bfloat16*>(o.data() + j * 128 + ((ch ^ (j & 7)) * 16));
      for (int e2 = 0; e2 < 8; ++e2) {
        float want = __bfloat162float(h[idx[j] * C + col0 + ch * 8 + e2]);
        if (__bfloat162float(got[e2]) != want) { if (bad < 5) printf("row %d ch %d e %d got %f want %f\n", j, ch, e2, __bfloat162float(got[e2]), want); ++bad; }
      }
    }
  printf("mismatches: %d of %d\n", bad, nrows * 64);
  return 0;
}
