// Auxiliary losses of the DualVGR train step, fused per video, value AND gradient in one launch.
//
//   common_loss      reference utils.py:10-18  : centre over nodes, L2-normalise rows, || E^E^T - E'^E'^T ||^2 (mean)
//   loss_dependence  reference utils.py:20-31  : HSIC = sum_b tr(R K1 R K2), R = I - 11^T/N
//
// Both only need the node-centred Gram matrices C = (RE)(RE)^T (N x N, in shared memory):
//   HSIC   = sum_ij C1_ij C2_ij                       d/dE1 = 2 C2 (R E1)
//   common = sum_ij (G1_ij - G2_ij)^2, G = C / (n_i n_j), n_i = sqrt(C_ii)
// The reference's B-iteration torch.trace loop (utils.py:29-30) and the .cpu()/.cuda() round trip of the four
// [B,N,D] tensors (model/models.py:153-160, train.py:152-153) do not exist here.
// Inputs are fp32: after centring, bf16 inputs would be rounding noise (SURVEY.md §7 "ill-conditioned auxiliary loss").
#include "capi_internal.h"
#include "ptx.cuh"

namespace dvgr {

constexpr int kLossThreads = 256;
constexpr int kChunk = 256;

// smem: tile [2][N][kChunk] f32, C [2][N][N] f32, Delta [N][N], nrm [2][N], rdot [2][N]
__host__ __device__ inline size_t pair_loss_smem(int N) {
  return (size_t)(2 * N * kChunk + 3 * N * N + 4 * N) * sizeof(float);
}

// Loads a column chunk of both tensors, centred over the nodes.
__device__ __forceinline__ void load_centered_chunk(const float* __restrict__ x, const float* __restrict__ y, int N, int D,
                                                    int c0, float* tile) {
  for (int cc = threadIdx.x; cc < kChunk; cc += blockDim.x) {
    const int c = c0 + cc;
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      const float* src = which ? y : x;
      float* t = tile + (size_t)which * N * kChunk;
      float s = 0.f;
      if (c < D)
        for (int n = 0; n < N; ++n) s += src[(long long)n * D + c];
      const float mean = s / N;
      for (int n = 0; n < N; ++n) t[n * kChunk + cc] = (c < D) ? src[(long long)n * D + c] - mean : 0.f;
    }
  }
}

// mode 0: common loss, mode 1: HSIC.   loss_part[b] = coef * value_b ; dx, dy (+)= coef * d value_b / d{x,y}
__global__ void __launch_bounds__(kLossThreads)
pair_loss_kernel(const float* __restrict__ X, const float* __restrict__ Y, int N, int D, int mode, float coef,
                 float* __restrict__ loss_part, float* __restrict__ dX, float* __restrict__ dY, int accumulate_x,
                 int accumulate_y) {
  extern __shared__ __align__(16) float sm[];
  float* tile = sm;                               // [2][N][kChunk]
  float* C = tile + 2 * N * kChunk;               // [2][N][N]
  float* Delta = C + 2 * N * N;                   // [N][N]
  float* nrm = Delta + N * N;                     // [2][N]
  float* rdot = nrm + 2 * N;                      // [2][N]
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = kLossThreads / 32;
  const float* x = X + (long long)b * N * D;
  const float* y = Y + (long long)b * N * D;

  for (int i = tid; i < 2 * N * N; i += kLossThreads) C[i] = 0.f;
  __syncthreads();
  // ---- pass 1: centred Gram matrices
  for (int c0 = 0; c0 < D; c0 += kChunk) {
    load_centered_chunk(x, y, N, D, c0, tile);
    __syncthreads();
    for (int pr = warp; pr < 2 * N * N; pr += nwarps) {
      const int which = pr / (N * N), r = pr - which * N * N, i = r / N, j = r - i * N;
      if (j < i) continue;
      const float* ti = tile + ((size_t)which * N + i) * kChunk;
      const float* tj = tile + ((size_t)which * N + j) * kChunk;
      float acc = 0.f;
#pragma unroll
      for (int q = 0; q < kChunk / 32; ++q) acc += ti[lane + 32 * q] * tj[lane + 32 * q];
      acc = warp_sum(acc);
      if (lane == 0) {
        C[which * N * N + i * N + j] += acc;
        if (i != j) C[which * N * N + j * N + i] += acc;
      }
    }
    __syncthreads();
  }
  // ---- N x N algebra
  if (mode == 0) {
    for (int i = tid; i < 2 * N; i += kLossThreads) {
      const int which = i / N, n = i - which * N;
      nrm[i] = fmaxf(sqrtf(fmaxf(C[which * N * N + n * N + n], 0.f)), 1e-12f);
    }
    __syncthreads();
    float part = 0.f;
    for (int e = tid; e < N * N; e += kLossThreads) {
      const int i = e / N, j = e - i * N;
      const float g1 = C[e] / (nrm[i] * nrm[j]);
      const float g2 = C[N * N + e] / (nrm[N + i] * nrm[N + j]);
      const float d = g1 - g2;
      part += d * d;
      Delta[e] = 2.f * coef * d;      // dL/dG1 ; dL/dG2 = -Delta
      C[e] = g1;                      // keep the normalised Grams for r_i
      C[N * N + e] = g2;
    }
    part = warp_sum(part);
    __shared__ float red[kLossThreads / 32];
    if (lane == 0) red[warp] = part;
    __syncthreads();
    if (tid == 0) {
      float s = 0.f;
      for (int w = 0; w < nwarps; ++w) s += red[w];
      loss_part[b] = coef * s;
    }
    // r_i = E^_i . dE^_i = 2 sum_j (+-Delta_ij) G_ij
    for (int i = tid; i < 2 * N; i += kLossThreads) {
      const int which = i / N, n = i - which * N;
      float s = 0.f;
      for (int j = 0; j < N; ++j) s += Delta[n * N + j] * C[which * N * N + n * N + j];
      rdot[i] = (which ? -2.f : 2.f) * s;
    }
  } else {
    float part = 0.f;
    for (int e = tid; e < N * N; e += kLossThreads) part += C[e] * C[N * N + e];
    part = warp_sum(part);
    __shared__ float red2[kLossThreads / 32];
    if (lane == 0) red2[warp] = part;
    __syncthreads();
    if (tid == 0) {
      float s = 0.f;
      for (int w = 0; w < nwarps; ++w) s += red2[w];
      loss_part[b] = coef * s;
    }
  }
  __syncthreads();
  if (dX == nullptr && dY == nullptr) return;

  // ---- pass 2: gradients, one thread per column of the chunk
  float* dx = dX ? dX + (long long)b * N * D : nullptr;
  float* dy = dY ? dY + (long long)b * N * D : nullptr;
  for (int c0 = 0; c0 < D; c0 += kChunk) {
    load_centered_chunk(x, y, N, D, c0, tile);
    __syncthreads();
    for (int cc = tid; cc < kChunk; cc += kLossThreads) {
      const int c = c0 + cc;
      if (c >= D) continue;
#pragma unroll
      for (int which = 0; which < 2; ++which) {
        float* dst = which ? dy : dx;
        if (dst == nullptr) continue;
        const int acc_flag = which ? accumulate_y : accumulate_x;
        const float* mine = tile + (size_t)which * N * kChunk;          // centred values of this tensor
        const float* other = tile + (size_t)(1 - which) * N * kChunk;
        if (mode == 1) {
          // d/dE_which = 2 coef * C_other (R E_which)   (column mean of the result is already zero)
          const float* Co = C + (1 - which) * N * N;
          for (int i = 0; i < N; ++i) {
            float s = 0.f;
            for (int j = 0; j < N; ++j) s += Co[i * N + j] * mine[j * kChunk + cc];
            s *= 2.f * coef;
            const long long o = (long long)i * D + c;
            dst[o] = acc_flag ? dst[o] + s : s;
          }
          (void)other;
        } else {
          // E^_j[c] for this column (the thread owns column cc of both tiles: in-place scaling is race-free)
          const float sign = which ? -1.f : 1.f;
          const float* nr = nrm + which * N;
          const float* rd = rdot + which * N;
          float* col = const_cast<float*>(mine);
          for (int j = 0; j < N; ++j) col[j * kChunk + cc] /= nr[j];
          // dE'_i = (dE^_i - E^_i r_i) / n_i with dE^_i = 2 sign sum_j Delta_ij E^_j ; dE = dE' - column mean(dE')
          float colsum = 0.f;
          for (int i = 0; i < N; ++i) {
            float sacc = 0.f;
            for (int j = 0; j < N; ++j) sacc += Delta[i * N + j] * col[j * kChunk + cc];
            colsum += (2.f * sign * sacc - col[i * kChunk + cc] * rd[i]) / nr[i];
          }
          const float mean = colsum / N;
          for (int i = 0; i < N; ++i) {
            float sacc = 0.f;
            for (int j = 0; j < N; ++j) sacc += Delta[i * N + j] * col[j * kChunk + cc];
            const float v = (2.f * sign * sacc - col[i * kChunk + cc] * rd[i]) / nr[i] - mean;
            const long long o = (long long)i * D + c;
            dst[o] = acc_flag ? dst[o] + v : v;
          }
        }
      }
    }
    __syncthreads();
  }
}

}  // namespace dvgr

using namespace dvgr;

extern "C" int dvgr_pair_loss(const float* x, const float* y, int B, int N, int D, int mode, float coef,
                              float* loss_part, float* dx, float* dy, int accumulate_x, int accumulate_y,
                              void* stream) {
  if (B <= 0) return 0;
  if (N < 1 || N > 64) return set_error("pair_loss: N=%d out of [1,64]", N);
  if (mode != 0 && mode != 1) return set_error("pair_loss: mode must be 0 (common) or 1 (HSIC)");
  const size_t smem = pair_loss_smem(N);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(pair_loss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error("pair_loss: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured = smem;
  }
  pair_loss_kernel<<<B, kLossThreads, smem, reinterpret_cast<cudaStream_t>(stream)>>>(x, y, N, D, mode, coef, loss_part,
                                                                                      dx, dy, accumulate_x, accumulate_y);
  DVGR_CHECK_LAUNCH("pair_loss");
  return 0;
}
