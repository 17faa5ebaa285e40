// Weighted BCE loss (forward + input gradient, deterministic two-stage reduction) and a
// flat-buffer Adam step.
#include <math.h>

#include "b3d_common.cuh"

namespace b3d {

constexpr int BCE_THREADS = 256, BCE_PER_THREAD = 8, BCE_CHUNK = BCE_THREADS * BCE_PER_THREAD;

__global__ void __launch_bounds__(BCE_THREADS) k_bce(const float* __restrict__ in, const int64_t* __restrict__ y,
                                                    const float* __restrict__ w, long long E, float gscale,
                                                    int from_logits, float* __restrict__ grad,
                                                    float* __restrict__ partials) {
  __shared__ float red[BCE_THREADS];
  float acc = 0.f;
  const long long base = (long long)blockIdx.x * BCE_CHUNK;
#pragma unroll
  for (int i = 0; i < BCE_PER_THREAD; ++i) {
    long long e = base + i * BCE_THREADS + threadIdx.x;
    if (e < E) {
      const float x = in[e];
      const float t = (float)y[e];
      const float we = w ? w[e] : 1.f;
      float l, g;
      if (from_logits) {  // BCEWithLogits: max(x,0) - x t + log1p(exp(-|x|))
        l = fmaxf(x, 0.f) - x * t + log1pf(expf(-fabsf(x)));
        g = 1.f / (1.f + expf(-x)) - t;
      } else {            // torch BCELoss: log clamped at -100; grad (p - t) / max(p (1-p), 1e-12)
        l = -(t * fmaxf(logf(x), -100.f) + (1.f - t) * fmaxf(logf(1.f - x), -100.f));
        g = (x - t) / fmaxf((1.f - x) * x, 1e-12f);
      }
      acc += we * l;
      if (grad) grad[e] = we * g * gscale;
    }
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = BCE_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partials[blockIdx.x] = red[0];
}

// Focal loss (config 5 "focal/BCE edge loss"; not in the reference): l = a_t (1 - p_t)^gamma (-log p_t),
// p_t = p for positives and 1 - p for negatives, a_t = alpha / 1 - alpha. Same two-stage reduction.
__global__ void __launch_bounds__(BCE_THREADS) k_focal(const float* __restrict__ in, const int64_t* __restrict__ y,
                                                      const float* __restrict__ w, long long E, float gscale,
                                                      int from_logits, float alpha, float gamma,
                                                      float* __restrict__ grad, float* __restrict__ partials) {
  __shared__ float red[BCE_THREADS];
  float acc = 0.f;
  const long long base = (long long)blockIdx.x * BCE_CHUNK;
#pragma unroll
  for (int i = 0; i < BCE_PER_THREAD; ++i) {
    long long e = base + i * BCE_THREADS + threadIdx.x;
    if (e < E) {
      const float x = in[e];
      const float p = from_logits ? 1.f / (1.f + expf(-x)) : x;
      const bool pos = y[e] != 0;
      const float we = w ? w[e] : 1.f;
      const float pt = pos ? p : 1.f - p;
      const float at = pos ? alpha : 1.f - alpha;
      const float lg = fmaxf(logf(pt), -100.f);
      const float om = 1.f - pt;
      const float f = powf(om, gamma);
      acc += we * at * f * -lg;
      if (grad) {
        // dl/dp_t = a_t [gamma (1-p_t)^(gamma-1) log p_t - (1-p_t)^gamma / p_t]
        const float fm1 = gamma == 0.f ? 0.f : gamma * powf(om, gamma - 1.f);
        float g = at * (fm1 * lg - f / fmaxf(pt, 1e-12f));
        g = pos ? g : -g;                       // dp_t/dp
        if (from_logits) g *= p * (1.f - p);    // dp/dx
        grad[e] = we * g * gscale;
      }
    }
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = BCE_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partials[blockIdx.x] = red[0];
}

__global__ void __launch_bounds__(1024) k_bce_final(const float* __restrict__ partials, long long n, float scale,
                                                    float* __restrict__ loss) {
  __shared__ float red[1024];
  float acc = 0.f;
  for (long long i = threadIdx.x; i < n; i += 1024) acc += partials[i];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) loss[0] = red[0] * scale;
}

__global__ void k_adam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                       float* __restrict__ v, long long n, float lr, float b1, float b2, float eps, float wd,
                       float bc1, float bc2_sqrt, float gscale) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float pi = p[i];
  float gi = g[i] * gscale;
  if (wd != 0.f) gi = fmaf(wd, pi, gi);
  float mi = b1 * m[i] + (1.f - b1) * gi;
  float vi = b2 * v[i] + (1.f - b2) * gi * gi;
  m[i] = mi;
  v[i] = vi;
  float denom = sqrtf(vi) / bc2_sqrt + eps;
  p[i] = pi - (lr / bc1) * (mi / denom);
}

// Same update with the step number kept ON THE DEVICE (bias corrections computed per thread from *step + 1), so
// that a captured CUDA graph of the training step replays with the right correction; k_adam_tick then advances it.
__global__ void k_adam_dev(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                           float* __restrict__ v, long long n, float lr, float b1, float b2, float eps, float wd,
                           const int32_t* __restrict__ step, float gscale) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float t = (float)(*step + 1);
  const float bc1 = 1.f - powf(b1, t), bc2_sqrt = sqrtf(1.f - powf(b2, t));
  float pi = p[i];
  float gi = g[i] * gscale;
  if (wd != 0.f) gi = fmaf(wd, pi, gi);
  float mi = b1 * m[i] + (1.f - b1) * gi;
  float vi = b2 * v[i] + (1.f - b2) * gi * gi;
  m[i] = mi;
  v[i] = vi;
  float denom = sqrtf(vi) / bc2_sqrt + eps;
  p[i] = pi - (lr / bc1) * (mi / denom);
}
__global__ void k_adam_tick(int32_t* step) { *step += 1; }

}  // namespace b3d

using namespace b3d;

extern "C" int b3d_adam_step_dev(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1,
                                 float beta2, float eps, float weight_decay, int32_t* step_counter, float grad_scale,
                                 void* stream) {
  if (!p || !g || !m || !v || n < 0 || !step_counter) return bad_arg("b3d_adam_step_dev");
  if (n > 0) {
    k_adam_dev<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr, beta1, beta2, eps,
                                                                              weight_decay, step_counter, grad_scale);
    B3D_LAUNCH_CHECK("k_adam_dev");
  }
  k_adam_tick<<<1, 1, 0, (cudaStream_t)stream>>>(step_counter);
  B3D_LAUNCH_CHECK("k_adam_tick");
  return 0;
}

extern "C" int64_t b3d_bce_partials(int64_t E) { return ceil_div(E > 0 ? E : 1, BCE_CHUNK); }

extern "C" int b3d_bce_fwd_bwd(const float* input, const int64_t* y, const float* w, int64_t E, float scale,
                               int32_t from_logits, float* loss_out, float* grad_out, float* partials,
                               void* stream) {
  if (!input || !y || !loss_out || !partials || E <= 0) return bad_arg("b3d_bce_fwd_bwd");
  cudaStream_t st = (cudaStream_t)stream;
  long long nb = b3d_bce_partials(E);
  k_bce<<<(unsigned)nb, BCE_THREADS, 0, st>>>(input, y, w, E, scale / (float)E, from_logits, grad_out, partials);
  B3D_LAUNCH_CHECK("k_bce");
  k_bce_final<<<1, 1024, 0, st>>>(partials, nb, scale / (float)E, loss_out);
  B3D_LAUNCH_CHECK("k_bce_final");
  return 0;
}

extern "C" int b3d_focal_fwd_bwd(const float* input, const int64_t* y, const float* w, int64_t E, float scale,
                                 int32_t from_logits, float alpha, float gamma, float* loss_out, float* grad_out,
                                 float* partials, void* stream) {
  if (!input || !y || !loss_out || !partials || E <= 0 || gamma < 0.f) return bad_arg("b3d_focal_fwd_bwd");
  cudaStream_t st = (cudaStream_t)stream;
  long long nb = b3d_bce_partials(E);
  k_focal<<<(unsigned)nb, BCE_THREADS, 0, st>>>(input, y, w, E, scale / (float)E, from_logits, alpha, gamma, grad_out,
                                                partials);
  B3D_LAUNCH_CHECK("k_focal");
  k_bce_final<<<1, 1024, 0, st>>>(partials, nb, scale / (float)E, loss_out);
  B3D_LAUNCH_CHECK("k_bce_final");
  return 0;
}

extern "C" int b3d_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1,
                             float beta2, float eps, float weight_decay, int32_t step, float grad_scale,
                             void* stream) {
  if (!p || !g || !m || !v || n < 0 || step < 1) return bad_arg("b3d_adam_step");
  if (n == 0) return 0;
  float bc1 = 1.f - powf(beta1, (float)step);
  float bc2 = 1.f - powf(beta2, (float)step);
  k_adam<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr, beta1, beta2, eps,
                                                                        weight_decay, bc1, sqrtf(bc2), grad_scale);
  B3D_LAUNCH_CHECK("k_adam");
  return 0;
}
