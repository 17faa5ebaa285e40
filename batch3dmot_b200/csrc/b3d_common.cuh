// Shared helpers for libb3d.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/b3d.h"

namespace b3d {

extern thread_local char g_err[256];
extern long long g_launches;

inline int fail(const char* what, cudaError_t e) {
  snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
  return (int)e;
}
inline int bad_arg(const char* what) {
  snprintf(g_err, sizeof(g_err), "bad argument: %s", what);
  return -1;
}

#define B3D_LAUNCH_CHECK(name)                                  \
  do {                                                          \
    ::b3d::g_launches++;                                        \
    cudaError_t e__ = cudaGetLastError();                       \
    if (e__ != cudaSuccess) return ::b3d::fail(name, e__);      \
  } while (0)

__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ float apply_mask(float v, float m, int mode) {
  if (mode == B3D_MASK_RELU) return m > 0.f ? v : 0.f;
  if (mode == B3D_MASK_SIGMOID) return v * m * (1.f - m);
  return v;
}

struct SegDev {
  const float* ptr;
  const int32_t* idx;
  const float* mask;
  int width, ld, ldmask, mask_mode, dtype;
};

inline int to_dev(const b3d_seg_t* in, int nseg, SegDev* out) {
  if (nseg < 1 || nseg > B3D_MAX_SEGS) return -1;
  for (int s = 0; s < nseg; ++s) {
    if (!in[s].ptr || in[s].width <= 0 || in[s].ld < in[s].width) return -1;
    out[s] = SegDev{in[s].ptr, in[s].idx, in[s].mask, in[s].width, in[s].ld,
                    in[s].mask ? in[s].ldmask : 0, in[s].mask ? in[s].mask_mode : B3D_MASK_NONE, in[s].dtype};
  }
  return 0;
}
inline bool all_f32(const SegDev* seg, int nseg) {
  for (int s = 0; s < nseg; ++s)
    if (seg[s].dtype != B3D_F32) return false;
  return true;
}

}  // namespace b3d
