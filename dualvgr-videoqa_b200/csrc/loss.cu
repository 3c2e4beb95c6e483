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
constexpr int kMaxJobs = 4;

struct PairJob {
  const float* x;
  const float* y;
  float* dx;          // may be null
  float* dy;
  float* loss_part;   // [B][loss_ld], this job writes column loss_col
  int loss_col, loss_ld;
  int mode;           // 0 common, 1 HSIC
  int acc_x, acc_y;   // 0 write, 1 read-modify-write, 2 atomicAdd (buffer zeroed by the caller; two addends => deterministic)
  float coef;
};
struct PairParams {
  PairJob job[kMaxJobs];
  int B, N, D, chunks;
  float* gram_ws;     // [jobs][B][chunks][2][N][N]
};

// Loads a column chunk of both tensors, centred over the nodes (columns beyond D are zero).
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

// pass 1: partial centred Gram matrices of one (video, job, column chunk)
__global__ void __launch_bounds__(kLossThreads) pair_gram_kernel(const PairParams p) {
  extern __shared__ __align__(16) float sm[];
  float* tile = sm;   // [2][N][kChunk]
  const int b = blockIdx.x, jb = blockIdx.y, ch = blockIdx.z, N = p.N, D = p.D;
  const PairJob& J = p.job[jb];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = kLossThreads / 32;
  load_centered_chunk(J.x + (long long)b * N * D, J.y + (long long)b * N * D, N, D, ch * kChunk, tile);
  __syncthreads();
  float* out = p.gram_ws + ((((long long)jb * p.B + b) * p.chunks + ch) * 2) * N * N;
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
      out[which * N * N + i * N + j] = acc;
      out[which * N * N + j * N + i] = acc;
    }
  }
}

__device__ __forceinline__ void emit(float* dst, long long o, float v, int acc) {
  if (acc == 0) dst[o] = v;
  else if (acc == 1) dst[o] += v;
  else atomicAdd(dst + o, v);
}

// pass 2: N x N algebra (redundantly per chunk CTA: it is tiny), loss value (chunk 0), gradient of this column chunk
__global__ void __launch_bounds__(kLossThreads) pair_grad_kernel(const PairParams p) {
  extern __shared__ __align__(16) float sm[];
  const int b = blockIdx.x, jb = blockIdx.y, ch = blockIdx.z, N = p.N, D = p.D;
  const PairJob& J = p.job[jb];
  float* tile = sm;                               // [2][N][kChunk]
  float* C = tile + 2 * N * kChunk;               // [2][N][N]
  float* Delta = C + 2 * N * N;                   // [N][N]
  float* nrm = Delta + N * N;                     // [2][N]
  float* rdot = nrm + 2 * N;                      // [2][N]
  __shared__ float red[kLossThreads / 32];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = kLossThreads / 32;
  const float* x = J.x + (long long)b * N * D;
  const float* y = J.y + (long long)b * N * D;
  const float coef = J.coef;
  const float* gw = p.gram_ws + (((long long)jb * p.B + b) * p.chunks * 2) * N * N;
  for (int e = tid; e < 2 * N * N; e += kLossThreads) {
    float s = 0.f;
    for (int k = 0; k < p.chunks; ++k) s += gw[(long long)k * 2 * N * N + e];
    C[e] = s;
  }
  load_centered_chunk(x, y, N, D, ch * kChunk, tile);
  __syncthreads();
  float part = 0.f;
  if (J.mode == 0) {
    for (int i = tid; i < 2 * N; i += kLossThreads) {
      const int which = i / N, n = i - which * N;
      nrm[i] = fmaxf(sqrtf(fmaxf(C[which * N * N + n * N + n], 0.f)), 1e-12f);
    }
    __syncthreads();
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
  } else {
    for (int e = tid; e < N * N; e += kLossThreads) part += C[e] * C[N * N + e];
  }
  part = warp_sum(part);
  if (lane == 0) red[warp] = part;
  __syncthreads();
  if (tid == 0 && ch == 0) {
    float s = 0.f;
    for (int w = 0; w < nwarps; ++w) s += red[w];
    J.loss_part[(long long)b * J.loss_ld + J.loss_col] = coef * s;
  }
  if (J.mode == 0) {
    for (int i = tid; i < 2 * N; i += kLossThreads) {     // r_i = E^_i . dE^_i = 2 sum_j (+-Delta_ij) G_ij
      const int which = i / N, n = i - which * N;
      float s = 0.f;
      for (int j = 0; j < N; ++j) s += Delta[n * N + j] * C[which * N * N + n * N + j];
      rdot[i] = (which ? -2.f : 2.f) * s;
    }
  }
  __syncthreads();
  if (J.dx == nullptr && J.dy == nullptr) return;
  float* dx = J.dx ? J.dx + (long long)b * N * D : nullptr;
  float* dy = J.dy ? J.dy + (long long)b * N * D : nullptr;
  for (int cc = tid; cc < kChunk; cc += kLossThreads) {
    const int c = ch * kChunk + cc;
    if (c >= D) continue;
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      float* dst = which ? dy : dx;
      if (dst == nullptr) continue;
      const int acc_flag = which ? J.acc_y : J.acc_x;
      float* col = tile + (size_t)which * N * kChunk;          // centred values of this tensor, column cc
      if (J.mode == 1) {
        // d/dE_which = 2 coef * C_other (R E_which)   (its column mean is already zero: rows of C sum to zero)
        const float* Co = C + (1 - which) * N * N;
        for (int i = 0; i < N; ++i) {
          float s = 0.f;
          for (int j = 0; j < N; ++j) s += Co[i * N + j] * col[j * kChunk + cc];
          emit(dst, (long long)i * D + c, 2.f * coef * s, acc_flag);
        }
      } else {
        // E^_j[c] (the thread owns column cc of both tiles: in-place scaling is race-free)
        const float sign = which ? -1.f : 1.f;
        const float* nr = nrm + which * N;
        const float* rd = rdot + which * N;
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
          emit(dst, (long long)i * D + c, v, acc_flag);
        }
      }
    }
  }
}

}  // namespace dvgr

using namespace dvgr;

extern "C" long long dvgr_pair_loss_workspace(int n_jobs, int B, int N, int D) {
  const int chunks = (D + kChunk - 1) / kChunk;
  return (long long)n_jobs * B * chunks * 2 * N * N;
}

extern "C" int dvgr_pair_loss_multi(const dvgr_pair_job* jobs, int n_jobs, int B, int N, int D, float* gram_ws,
                                    void* stream) {
  if (B <= 0 || n_jobs <= 0) return 0;
  if (n_jobs > kMaxJobs) return set_error("pair_loss: n_jobs=%d > %d", n_jobs, kMaxJobs);
  if (N < 1 || N > 64) return set_error("pair_loss: N=%d out of [1,64]", N);
  if (!gram_ws) return set_error("pair_loss: null workspace");
  PairParams p;
  memset(&p, 0, sizeof(p));
  p.B = B; p.N = N; p.D = D; p.chunks = (D + kChunk - 1) / kChunk; p.gram_ws = gram_ws;
  for (int i = 0; i < n_jobs; ++i) {
    const dvgr_pair_job& s = jobs[i];
    if (s.mode != 0 && s.mode != 1) return set_error("pair_loss: mode must be 0 (common) or 1 (HSIC)");
    if (!s.x || !s.y || !s.loss_part) return set_error("pair_loss: job %d has a null buffer", i);
    PairJob& d = p.job[i];
    d.x = s.x; d.y = s.y; d.dx = s.dx; d.dy = s.dy; d.loss_part = s.loss_part; d.loss_col = s.loss_col;
    d.loss_ld = s.loss_ld; d.mode = s.mode; d.acc_x = s.accumulate_x; d.acc_y = s.accumulate_y; d.coef = s.coef;
  }
  const size_t smem1 = (size_t)2 * N * kChunk * sizeof(float);
  const size_t smem2 = (size_t)(2 * N * kChunk + 3 * N * N + 4 * N) * sizeof(float);
  static size_t conf1 = 0, conf2 = 0;
  if (smem1 > 48 * 1024 && smem1 > conf1) {
    cudaError_t e = cudaFuncSetAttribute(pair_gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1);
    if (e != cudaSuccess) return set_error("pair_loss: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    conf1 = smem1;
  }
  if (smem2 > 48 * 1024 && smem2 > conf2) {
    cudaError_t e = cudaFuncSetAttribute(pair_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
    if (e != cudaSuccess) return set_error("pair_loss: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    conf2 = smem2;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid(B, n_jobs, p.chunks);
  pair_gram_kernel<<<grid, kLossThreads, smem1, st>>>(p);
  DVGR_CHECK_LAUNCH("pair_gram");
  pair_grad_kernel<<<grid, kLossThreads, smem2, st>>>(p);
  DVGR_CHECK_LAUNCH("pair_grad");
  return 0;
}
