// Flat-buffer optimizer step of the DualVGR train loop: global-norm clipping + Adam in two launches over ONE fp32
// buffer (the reference walks ~250 parameter tensors: nn.utils.clip_grad_norm_(max_norm=12) + optim.Adam(lr=1e-4),
// train.py:85,158-159). Matches torch.optim.Adam (no weight decay, no amsgrad) and clip_grad_norm_ (eps 1e-6).
#include <cuda_bf16.h>

#include "capi_internal.h"
#include "ptx.cuh"

namespace dvgr {

__global__ void sumsq_partial_kernel(const float* __restrict__ g, long long n, float* __restrict__ partial) {
  float acc = 0.f;
  const long long n4 = n / 4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(g)[i];
    acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (long long i = n4 * 4; i < n; ++i) acc += g[i] * g[i];
  acc = warp_sum(acc);
  __shared__ float red[32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) red[warp] = acc;
  __syncthreads();
  if (warp == 0) {
    float v = lane < (blockDim.x >> 5) ? red[lane] : 0.f;
    v = warp_sum(v);
    if (lane == 0) partial[blockIdx.x] = v;
  }
}
__global__ void sumsq_final_kernel(const float* __restrict__ partial, int n, float* __restrict__ out) {
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += 32) acc += partial[i];
  acc = warp_sum(acc);
  if (threadIdx.x == 0) out[0] = acc;
}

// norm_sq: device scalar = sum of squares of the (already averaged) gradient. grad_scale multiplies g before use.
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, long long n, float lr, float b1, float b2, float eps, float bc1,
                            float bc2_sqrt, float max_norm, const float* __restrict__ norm_sq, float grad_scale,
                            const int* __restrict__ step_dev, __nv_bfloat16* __restrict__ shadow) {
  if (step_dev != nullptr) {      // device-resident step counter (CUDA-graph replay): bias corrections computed here
    const float t = (float)step_dev[0];
    bc1 = 1.f - powf(b1, t);
    bc2_sqrt = sqrtf(1.f - powf(b2, t));
  }
  float clip = 1.f;
  if (max_norm > 0.f && norm_sq != nullptr) {
    const float total = sqrtf(norm_sq[0]) * grad_scale;
    clip = fminf(1.f, max_norm / (total + 1e-6f));
  }
  const float gs = grad_scale * clip;
  const float step = lr / bc1;
  // 4 parameters per thread: 16-byte loads / stores of p, g, m, v (+ one 8-byte store of the bf16 shadow)
  const long long n4 = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                         reinterpret_cast<uintptr_t>(v)) & 15) == 0 && (shadow == nullptr || (reinterpret_cast<uintptr_t>(shadow) & 7) == 0)
                           ? n / 4 : 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 g4 = reinterpret_cast<const float4*>(g)[i];
    float4 m4 = reinterpret_cast<const float4*>(m)[i], v4 = reinterpret_cast<const float4*>(v)[i];
    float4 p4 = reinterpret_cast<const float4*>(p)[i];
    const float gi[4] = {g4.x * gs, g4.y * gs, g4.z * gs, g4.w * gs};
    float* mp = reinterpret_cast<float*>(&m4);
    float* vp = reinterpret_cast<float*>(&v4);
    float* pp = reinterpret_cast<float*>(&p4);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      mp[q] = b1 * mp[q] + (1.f - b1) * gi[q];
      vp[q] = b2 * vp[q] + (1.f - b2) * gi[q] * gi[q];
      pp[q] = pp[q] - step * mp[q] / (sqrtf(vp[q]) / bc2_sqrt + eps);
    }
    reinterpret_cast<float4*>(m)[i] = m4;
    reinterpret_cast<float4*>(v)[i] = v4;
    reinterpret_cast<float4*>(p)[i] = p4;
    if (shadow != nullptr) {
      uint2 o;
      o.x = pack_bf16x2(pp[0], pp[1]);
      o.y = pack_bf16x2(pp[2], pp[3]);
      reinterpret_cast<uint2*>(shadow)[i] = o;                     // bf16 operand copy for the next step's GEMMs
    }
  }
  for (long long i = n4 * 4 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * gs;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float pn = p[i] - step * mi / (sqrtf(vi) / bc2_sqrt + eps);
    p[i] = pn;
    if (shadow != nullptr) shadow[i] = __float2bfloat16_rn(pn);
  }
}

// End-of-step bookkeeping in one tiny launch: total = ce + sum of the auxiliary-loss partials (coef-scaled per-sample values
// of common_loss / loss_dependence, train.py:148-154), and the step's health: if any LSTM dependency poll gave up (sticky words
// of dvgr_lstm_seq_*), the reported loss is NaN — a replayed CUDA graph can then never train silently on stale state.
struct FlagList {
  const int* flag[16];
  int n;
};
__global__ void finalize_loss_kernel(const float* __restrict__ ce, const float* __restrict__ parts, int rows,
                                     const FlagList F, float* __restrict__ out) {
  __shared__ float red[3][32];
  float a[3] = {0.f, 0.f, 0.f};
  for (int i = threadIdx.x; i < rows; i += blockDim.x) {
    a[0] += parts[3 * i]; a[1] += parts[3 * i + 1]; a[2] += parts[3 * i + 2];
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    a[c] = warp_sum(a[c]);
    if (lane == 0) red[c][warp] = a[c];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float s[3] = {0.f, 0.f, 0.f};
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w)
      for (int c = 0; c < 3; ++c) s[c] += red[c][w];
    int bad = 0;
    for (int i = 0; i < F.n; ++i) bad += (F.flag[i][0] != 0) ? 1 : 0;
    const float total = ce[0] + s[0] + s[1] + s[2];
    out[0] = bad ? __int_as_float(0x7fc00000) : total;
    out[1] = s[0];
    out[2] = s[1] + s[2];
    out[3] = (float)bad;
  }
}

}  // namespace dvgr

using namespace dvgr;

extern "C" int dvgr_finalize_loss(const float* ce, const float* parts, int rows, const int* const* flags, int n_flags,
                                  float* out, void* stream) {
  if (!ce || !out) return set_error("finalize_loss: null buffer");
  if (n_flags < 0 || n_flags > 16) return set_error("finalize_loss: n_flags=%d out of [0,16]", n_flags);
  FlagList F;
  F.n = n_flags;
  for (int i = 0; i < n_flags; ++i) {
    if (!flags[i]) return set_error("finalize_loss: null flag %d", i);
    F.flag[i] = flags[i];
  }
  finalize_loss_kernel<<<1, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(ce, parts, parts ? rows : 0, F, out);
  DVGR_CHECK_LAUNCH("finalize_loss");
  return 0;
}

extern "C" int dvgr_sumsq_blocks(void) { return 1024; }

extern "C" int dvgr_sumsq(const float* g, long long n, float* partial_ws, float* out, void* stream) {
  if (n <= 0) return 0;
  if ((reinterpret_cast<uintptr_t>(g) & 15) != 0) return set_error("sumsq: buffer must be 16-byte aligned");
  const int blocks = dvgr_sumsq_blocks();
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  sumsq_partial_kernel<<<blocks, 256, 0, st>>>(g, n, partial_ws);
  DVGR_CHECK_LAUNCH("sumsq_partial");
  sumsq_final_kernel<<<1, 32, 0, st>>>(partial_ws, blocks, out);
  DVGR_CHECK_LAUNCH("sumsq_final");
  return 0;
}

extern "C" int dvgr_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n,
                              float lr, float beta1, float beta2, float eps, int step, float max_norm,
                              const float* norm_sq, float grad_scale, const int* step_dev, void* bf16_shadow,
                              void* stream) {
  if (n <= 0) return 0;
  if (step < 1 && step_dev == nullptr) return set_error("adam: step must be >= 1");
  if (step < 1) step = 1;
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2 = 1.f - powf(beta2, (float)step);
  long long blocks = (n / 4 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  adam_kernel<<<(int)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, bc1, sqrtf(bc2), max_norm, norm_sq, grad_scale, step_dev,
      reinterpret_cast<__nv_bfloat16*>(bf16_shadow));
  DVGR_CHECK_LAUNCH("adam_step");
  return 0;
}
