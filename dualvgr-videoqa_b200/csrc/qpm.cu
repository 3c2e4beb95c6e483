// Query Punishment Module, fused: question-conditioned word attention (QueryAttn tail) and the per-clip sigmoid gates
// (QueryPunish) — forward and backward. Memory-bound: vectorised coalesced loads, warp-shuffle reductions.
//
//   reference model/utils.py:66-84  (QueryAttn.forward: normalize, fc, softmax over ALL L, mask, renormalise, bmm)
//   reference model/utils.py:92-105 (QueryPunish.forward: query = W q_c + b, sigmoid(X . query))
//
// The 768->768 feat_enhance product and the 300->768 query projections are tcgen05 GEMMs (gemm.cu); these kernels own
// everything between them. The reference's per-sample Python loop with a host sync per sample (utils.py:73-75) is a
// register-level mask from question_len here.
#include <cuda_bf16.h>

#include "act.cuh"
#include "capi_internal.h"
#include "ptx.cuh"

namespace dvgr {
namespace DVGR_VNS {
using namespace act;

constexpr int kQThreads = 256;
constexpr int kMaxL = 128;

// ---------------------------------------------------------------------------------------------- word attention fwd
// y [B][L][D] bf16 = feat_enhance(dynamic_q) (bias included); words [B][L][ld_w] bf16
// out: alpha [B][L], nrm [B][L] (clamped norm), prob [B][L] (softmax over all L), ssum [B]; qc [B][ld_qc] bf16 (zero padded)
__global__ void __launch_bounds__(kQThreads)
qattn_fwd_kernel(const act_t* __restrict__ y, const float* __restrict__ wf, const float* __restrict__ cf,
                 const int* __restrict__ qlen, const act_t* __restrict__ words, long long ld_w, int L, int D,
                 int W, float* __restrict__ alpha, float* __restrict__ nrm, float* __restrict__ prob,
                 float* __restrict__ ssum, act_t* __restrict__ qc, long long ld_qc) {
  __shared__ float sc[kMaxL];
  __shared__ float al[kMaxL];
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = kQThreads / 32;
  const act_t* yb = y + (long long)b * L * D;
  for (int l = warp; l < L; l += nwarps) {
    float n2 = 0.f, dot = 0.f;
    const act_t* row = yb + (long long)l * D;
    for (int c = lane * 8; c < D; c += 256) {
      float w8_[8];
      ld8(row + c, w8_);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float2 f = make_float2(w8_[2 * q], w8_[2 * q + 1]);
        n2 += f.x * f.x + f.y * f.y;
        dot += f.x * wf[c + 2 * q] + f.y * wf[c + 2 * q + 1];
      }
    }
    n2 = warp_sum(n2);
    dot = warp_sum(dot);
    if (lane == 0) {
      const float n = fmaxf(sqrtf(n2), 1e-12f);
      nrm[(long long)b * L + l] = n;
      sc[l] = dot / n + cf[0];
    }
  }
  __syncthreads();
  if (warp == 0) {
    float m = -INFINITY;
    for (int l = lane; l < L; l += 32) m = fmaxf(m, sc[l]);
    m = warp_max(m);
    float s = 0.f;
    for (int l = lane; l < L; l += 32) s += __expf(sc[l] - m);
    s = warp_sum(s);
    const int len = qlen[b];
    float ms = 0.f;
    for (int l = lane; l < L; l += 32) {
      const float p = __expf(sc[l] - m) / s;
      prob[(long long)b * L + l] = p;
      sc[l] = p;
      if (l < len) ms += p;
    }
    ms = warp_sum(ms) + 1e-5f;
    if (lane == 0) ssum[b] = ms;
    for (int l = lane; l < L; l += 32) {
      const float a = (l < len ? sc[l] : 0.f) / ms;
      al[l] = a;
      alpha[(long long)b * L + l] = a;
    }
  }
  __syncthreads();
  const act_t* wb = words + (long long)b * L * ld_w;
  for (int w = tid; w < ld_qc; w += kQThreads) {
    float acc = 0.f;
    if (w < W)
      for (int l = 0; l < L; ++l) acc += al[l] * ld1(&wb[(long long)l * ld_w + w]);
    st1(&qc[(long long)b * ld_qc + w], acc);
  }
}

// ---------------------------------------------------------------------------------------------- word attention bwd
// dqc [B][ld_qc] bf16 -> dy [B][L][D] bf16, dwords [B][L][ld_w] bf16 (+= when accumulate), dwf_part [B][D], dcf_part [B]
__global__ void __launch_bounds__(kQThreads)
qattn_bwd_kernel(const act_t* __restrict__ dqc, long long ld_qc, const act_t* __restrict__ y,
                 const float* __restrict__ wf, const int* __restrict__ qlen, const act_t* __restrict__ words,
                 long long ld_w, int L, int D, int W, const float* __restrict__ alpha, const float* __restrict__ nrm,
                 const float* __restrict__ prob, const float* __restrict__ ssum, act_t* __restrict__ dy,
                 act_t* __restrict__ dwords, int accumulate, float* __restrict__ dwf_part,
                 float* __restrict__ dcf_part) {
  __shared__ float dal[kMaxL];    // d alpha, then d score
  __shared__ float dq[512];       // dqc of this sample (W <= 512)
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = kQThreads / 32;
  for (int w = tid; w < W; w += kQThreads) dq[w] = ld1(&dqc[(long long)b * ld_qc + w]);
  __syncthreads();
  const act_t* wb = words + (long long)b * L * ld_w;
  act_t* dwb = dwords + (long long)b * L * ld_w;
  // d alpha_l = dqc . words_l ; dwords_l = alpha_l * dqc
  for (int l = warp; l < L; l += nwarps) {
    const float a = alpha[(long long)b * L + l];
    float acc = 0.f;
    for (int w = lane; w < W; w += 32) {
      acc += dq[w] * ld1(&wb[(long long)l * ld_w + w]);
      float g = a * dq[w];
      if (accumulate) g += ld1(&dwb[(long long)l * ld_w + w]);
      st1(&dwb[(long long)l * ld_w + w], g);
    }
    acc = warp_sum(acc);
    if (lane == 0) dal[l] = acc;
  }
  __syncthreads();
  if (warp == 0) {
    const int len = qlen[b];
    const float s = ssum[b];
    float t = 0.f;   // sum_k dalpha_k alpha_k
    for (int l = lane; l < L; l += 32) t += dal[l] * alpha[(long long)b * L + l];
    t = warp_sum(t);
    float pd = 0.f;  // sum_k p_k dp_k
    float dp[kMaxL / 32];
    int q = 0;
    for (int l = lane; l < L; l += 32, ++q) {
      dp[q] = (l < len) ? (dal[l] - t) / s : 0.f;
      pd += prob[(long long)b * L + l] * dp[q];
    }
    pd = warp_sum(pd);
    float dcsum = 0.f;
    q = 0;
    for (int l = lane; l < L; l += 32, ++q) {
      const float ds = prob[(long long)b * L + l] * (dp[q] - pd);
      dal[l] = ds;
      dcsum += ds;
    }
    dcsum = warp_sum(dcsum);
    if (lane == 0) dcf_part[b] = dcsum;
  }
  __syncthreads();
  // score_l = wf.y_l / n_l + cf  ->  dy_l = ds_l * (wf / n - (wf.y) y / n^3)   (n clamped: gradient of the clamp branch is wf/n)
  const act_t* yb = y + (long long)b * L * D;
  act_t* dyb = dy + (long long)b * L * D;
  for (int l = warp; l < L; l += nwarps) {
    const float n = nrm[(long long)b * L + l];
    const float ds = dal[l];
    const act_t* row = yb + (long long)l * D;
    float dot = 0.f;
    for (int c = lane * 8; c < D; c += 256) {
      float w8_[8];
      ld8(row + c, w8_);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float2 f = make_float2(w8_[2 * q], w8_[2 * q + 1]);
        dot += f.x * wf[c + 2 * q] + f.y * wf[c + 2 * q + 1];
      }
    }
    dot = warp_sum(dot);
    const bool clamped = n <= 1e-12f;
    const float k1 = ds / n, k2 = clamped ? 0.f : ds * dot / (n * n * n);
    for (int c = lane * 8; c < D; c += 256) {
      float w8_[8];
      ld8(row + c, w8_);
      float o8_[8];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float2 f = make_float2(w8_[2 * q], w8_[2 * q + 1]);
        o8_[2 * q] = k1 * wf[c + 2 * q] - k2 * f.x;
        o8_[2 * q + 1] = k1 * wf[c + 2 * q + 1] - k2 * f.y;
      }
      st8(dyb + (long long)l * D + c, o8_);
    }
  }
  // d wf (partial for this sample) = sum_l ds_l * y_l / n_l
  for (int c = tid; c < D; c += kQThreads) {
    float acc = 0.f;
    for (int l = 0; l < L; ++l) acc += dal[l] / nrm[(long long)b * L + l] * ld1(&yb[(long long)l * D + c]);
    dwf_part[(long long)b * D + c] = acc;
  }
}

// ---------------------------------------------------------------------------------------------- gates fwd / bwd
// X [S streams][B][N][D] via per-stream pointers; query [B][ld_q] bf16 with stream s at column s*D
struct GateParams {
  const act_t* X[2];
  act_t* dX[2];
  const act_t* query;
  act_t* dquery;
  long long ld_q;
  float* gate[2];
  const float* dgate_a[2];
  const float* dgate_b[2];
  int B, N, D, nstream;
};

__global__ void __launch_bounds__(256) gate_fwd_kernel(const GateParams p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long rows = (long long)p.B * p.N;
  const int s = blockIdx.y;
  for (long long r = (long long)blockIdx.x * 8 + warp; r < rows; r += (long long)gridDim.x * 8) {
    const int b = (int)(r / p.N);
    const act_t* x = p.X[s] + r * p.D;
    const act_t* q = p.query + (long long)b * p.ld_q + (long long)s * p.D;
    float acc = 0.f;
    for (int c = lane * 8; c < p.D; c += 256) {
      float xw8_[8], qw8_[8];
      ld8(x + c, xw8_);
      ld8(q + c, qw8_);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float2 a = make_float2(xw8_[2 * k], xw8_[2 * k + 1]), c2 = make_float2(qw8_[2 * k], qw8_[2 * k + 1]);
        acc += a.x * c2.x + a.y * c2.y;
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) p.gate[s][r] = sigmoidf_(acc);
  }
}

// one CTA per (video, stream): dz_n = (dgate_a + dgate_b) g (1-g); dX_n += dz_n * query ; dquery = sum_n dz_n X_n
__global__ void __launch_bounds__(256) gate_bwd_kernel(const GateParams p) {
  __shared__ float dz[64];
  const int b = blockIdx.x, s = blockIdx.y, tid = threadIdx.x;
  for (int n = tid; n < p.N; n += blockDim.x) {
    const long long r = (long long)b * p.N + n;
    const float g = p.gate[s][r];
    float d = p.dgate_a[s][r];
    if (p.dgate_b[s] != nullptr) d += p.dgate_b[s][r];
    dz[n] = d * g * (1.f - g);
  }
  __syncthreads();
  const act_t* q = p.query + (long long)b * p.ld_q + (long long)s * p.D;
  act_t* dq = p.dquery + (long long)b * p.ld_q + (long long)s * p.D;
  for (int c = tid * 2; c < p.D; c += blockDim.x * 2) {
    const float2 qv = ld2(q + c);
    float ax = 0.f, ay = 0.f;
    for (int n = 0; n < p.N; ++n) {
      const long long off = ((long long)b * p.N + n) * p.D + c;
      const float2 xv = ld2(p.X[s] + off);
      ax += dz[n] * xv.x;
      ay += dz[n] * xv.y;
      float2 old = ld2(p.dX[s] + off);
      old.x += dz[n] * qv.x;
      old.y += dz[n] * qv.y;
      st2(p.dX[s] + off, old.x, old.y);
    }
    st2(dq + c, ax, ay);
  }
}

}  // namespace DVGR_VNS
}  // namespace dvgr

using namespace dvgr;
using namespace dvgr::DVGR_VNS;
typedef act_t bf16;      // (the launchers below cast the ABI's void* activations to the build's storage type)

extern "C" int DVGR_FN(dvgr_qattn_fwd)(const void* y, const float* wf, const float* cf, const int* qlen, const void* words,
                              long long ld_w, int B, int L, int D, int W, float* alpha, float* nrm, float* prob,
                              float* ssum, void* qc, long long ld_qc, void* stream) {
  if (B <= 0) return 0;
  if (L > kMaxL) return set_error("qattn: L=%d > %d", L, kMaxL);
  if (D % 8 != 0) return set_error("qattn: D=%d must be a multiple of 8", D);
  qattn_fwd_kernel<<<B, kQThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const bf16*>(y), wf, cf, qlen, reinterpret_cast<const bf16*>(words), ld_w, L, D, W, alpha, nrm,
      prob, ssum, reinterpret_cast<bf16*>(qc), ld_qc);
  DVGR_CHECK_LAUNCH("qattn_fwd");
  return 0;
}

extern "C" int DVGR_FN(dvgr_qattn_bwd)(const void* dqc, long long ld_qc, const void* y, const float* wf, const int* qlen,
                              const void* words, long long ld_w, int B, int L, int D, int W, const float* alpha,
                              const float* nrm, const float* prob, const float* ssum, void* dy, void* dwords,
                              int accumulate_dwords, float* dwf_part, float* dcf_part, void* stream) {
  if (B <= 0) return 0;
  if (L > kMaxL) return set_error("qattn: L=%d > %d", L, kMaxL);
  if (W > 512) return set_error("qattn: W=%d > 512", W);
  if (D % 8 != 0) return set_error("qattn: D=%d must be a multiple of 8", D);
  qattn_bwd_kernel<<<B, kQThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const bf16*>(dqc), ld_qc, reinterpret_cast<const bf16*>(y), wf, qlen,
      reinterpret_cast<const bf16*>(words), ld_w, L, D, W, alpha, nrm, prob, ssum, reinterpret_cast<bf16*>(dy),
      reinterpret_cast<bf16*>(dwords), accumulate_dwords, dwf_part, dcf_part);
  DVGR_CHECK_LAUNCH("qattn_bwd");
  return 0;
}

extern "C" int DVGR_FN(dvgr_gate_fwd)(const void* x0, const void* x1, const void* query, long long ld_q, int B, int N, int D,
                             float* gate0, float* gate1, void* stream) {
  if (B <= 0 || N <= 0) return 0;
  if (D % 8 != 0) return set_error("gate: D=%d must be a multiple of 8", D);
  GateParams p;
  memset(&p, 0, sizeof(p));
  p.X[0] = reinterpret_cast<const bf16*>(x0); p.X[1] = reinterpret_cast<const bf16*>(x1);
  p.query = reinterpret_cast<const bf16*>(query); p.ld_q = ld_q;
  p.gate[0] = gate0; p.gate[1] = gate1;
  p.B = B; p.N = N; p.D = D; p.nstream = x1 ? 2 : 1;
  const long long rows = (long long)B * N;
  int blocks = (int)std::min<long long>((rows + 7) / 8, 148 * 8);
  gate_fwd_kernel<<<dim3(blocks, p.nstream), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  DVGR_CHECK_LAUNCH("gate_fwd");
  return 0;
}

extern "C" int DVGR_FN(dvgr_gate_bwd)(const void* x0, const void* x1, const void* query, long long ld_q, int B, int N, int D,
                             const float* gate0, const float* gate1, const float* dg0a, const float* dg0b,
                             const float* dg1a, const float* dg1b, void* dx0, void* dx1, void* dquery, void* stream) {
  if (B <= 0 || N <= 0) return 0;
  if (N > 64) return set_error("gate: N=%d > 64", N);
  GateParams p;
  memset(&p, 0, sizeof(p));
  p.X[0] = reinterpret_cast<const bf16*>(x0); p.X[1] = reinterpret_cast<const bf16*>(x1);
  p.dX[0] = reinterpret_cast<bf16*>(dx0); p.dX[1] = reinterpret_cast<bf16*>(dx1);
  p.query = reinterpret_cast<const bf16*>(query); p.dquery = reinterpret_cast<bf16*>(dquery); p.ld_q = ld_q;
  p.gate[0] = const_cast<float*>(gate0); p.gate[1] = const_cast<float*>(gate1);
  p.dgate_a[0] = dg0a; p.dgate_b[0] = dg0b; p.dgate_a[1] = dg1a; p.dgate_b[1] = dg1b;
  p.B = B; p.N = N; p.D = D; p.nstream = x1 ? 2 : 1;
  gate_bwd_kernel<<<dim3(B, p.nstream), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  DVGR_CHECK_LAUNCH("gate_bwd");
  return 0;
}
