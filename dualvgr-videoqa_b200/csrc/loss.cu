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

constexpr int kTS = kChunk + 4;     // padded tile row stride (floats): conflict-free TF32 fragment loads

__host__ __device__ inline int round16(int n) { return (n + 15) & ~15; }

// Loads a column chunk of both tensors, centred over the nodes, into tile[2][NP16][kTS]
// (columns beyond D and rows beyond N are zero).
__device__ __forceinline__ void load_centered_chunk(const float* __restrict__ x, const float* __restrict__ y, int N, int D,
                                                    int c0, float* tile) {
  const int NP = round16(N);
  for (int cc = threadIdx.x; cc < kChunk; cc += blockDim.x) {
    const int c = c0 + cc;
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      const float* src = which ? y : x;
      float* t = tile + (size_t)which * NP * kTS;
      float s = 0.f;
      if (c < D)
        for (int n = 0; n < N; ++n) s += src[(long long)n * D + c];
      const float mean = s / N;
      for (int n = 0; n < N; ++n) t[n * kTS + cc] = (c < D) ? src[(long long)n * D + c] - mean : 0.f;
      for (int n = N; n < NP; ++n) t[n * kTS + cc] = 0.f;
    }
  }
}

// pass 1: partial centred Gram matrices of one (video, job, column chunk): T T^T on the tensor cores (error-compensated 3 x TF32 m16n8k8: ptx.cuh)
__global__ void __launch_bounds__(kLossThreads) pair_gram_kernel(const PairParams p) {
  extern __shared__ __align__(16) float sm[];
  float* tile = sm;   // [2][NP16][kTS]
  const int b = blockIdx.x, jb = blockIdx.y, ch = blockIdx.z, N = p.N, D = p.D, NP = round16(N);
  const PairJob& J = p.job[jb];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = kLossThreads / 32, g = lane >> 2, t = lane & 3;
  load_centered_chunk(J.x + (long long)b * N * D, J.y + (long long)b * N * D, N, D, ch * kChunk, tile);
  __syncthreads();
  float* out = p.gram_ws + ((((long long)jb * p.B + b) * p.chunks + ch) * 2) * N * N;
  const int MT = NP / 16, NT = NP / 8;
  for (int ot = warp; ot < 2 * MT * NT; ot += nwarps) {
    const int which = ot / (MT * NT), r = ot - which * MT * NT, mt = r / NT, nt = r - mt * NT;
    const float* T = tile + (size_t)which * NP * kTS;
    const float* ar = T + (size_t)(mt * 16 + g) * kTS + t;
    const float* br = T + (size_t)(nt * 8 + g) * kTS + t;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
    for (int k0 = 0; k0 < kChunk; k0 += 8)
      mma_tf32x3(acc, ar[k0], ar[8 * kTS + k0], ar[k0 + 4], ar[8 * kTS + k0 + 4], br[k0], br[k0 + 4]);
    const int row = mt * 16 + g, col = nt * 8 + 2 * t;
    float* o = out + which * N * N;
    if (row < N && col < N) o[row * N + col] = acc[0];
    if (row < N && col + 1 < N) o[row * N + col + 1] = acc[1];
    if (row + 8 < N && col < N) o[(row + 8) * N + col] = acc[2];
    if (row + 8 < N && col + 1 < N) o[(row + 8) * N + col + 1] = acc[3];
  }
}

__device__ __forceinline__ void emit(float* dst, long long o, float v, int acc) {
  if (acc == 0) dst[o] = v;
  else if (acc == 1) dst[o] += v;
  else atomicAdd(dst + o, v);
}
// two adjacent columns at once (o even, 8-byte aligned): one 8-byte store / one vector reduction instead of two scalar ones
__device__ __forceinline__ void emit2(float* dst, long long o, float v0, float v1, int acc) {
  float2* d2 = reinterpret_cast<float2*>(dst + o);
  if (acc == 0) {
    *d2 = make_float2(v0, v1);
  } else if (acc == 1) {
    const float2 old = *d2;
    *d2 = make_float2(old.x + v0, old.y + v1);
  } else {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(d2), "f"(v0), "f"(v1) : "memory");
  }
}

// pass 2: N x N algebra (redundantly per chunk CTA: it is tiny), loss value (chunk 0), gradient of this column chunk
// as (N x N) . (N x 256) products on the tensor cores.
__host__ __device__ inline size_t pair_grad_smem(int N) {
  const int NP = round16(N), CS = NP + 4;
  return (size_t)(2 * NP * kTS + 3 * NP * CS + 6 * NP) * sizeof(float);
}

__global__ void __launch_bounds__(kLossThreads) pair_grad_kernel(const PairParams p) {
  extern __shared__ __align__(16) float sm[];
  const int b = blockIdx.x, jb = blockIdx.y, ch = blockIdx.z, N = p.N, D = p.D, NP = round16(N), CS = NP + 4;
  const PairJob& J = p.job[jb];
  float* tile = sm;                               // [2][NP][kTS]
  float* C = tile + 2 * NP * kTS;                 // [2][NP][CS]  (zero padded)
  float* Delta = C + 2 * NP * CS;                 // [NP][CS]
  float* nrm = Delta + NP * CS;                   // [2][NP]
  float* rdot = nrm + 2 * NP;                     // [2][NP]
  float* inv_nrm = rdot + 2 * NP;                 // [2][NP]
  __shared__ float red[kLossThreads / 32];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = kLossThreads / 32, g = lane >> 2, t = lane & 3;
  const float* x = J.x + (long long)b * N * D;
  const float* y = J.y + (long long)b * N * D;
  const float coef = J.coef;
  const float* gw = p.gram_ws + (((long long)jb * p.B + b) * p.chunks * 2) * N * N;
  for (int row = warp; row < 2 * NP; row += nwarps) {      // one warp per Gram row: no integer divisions in the loop
    const int which = row >= NP ? 1 : 0, i = row - which * NP;
    for (int j = lane; j < CS; j += 32) {
      float s = 0.f;
      if (i < N && j < N)
        for (int k = 0; k < p.chunks; ++k) s += gw[(long long)k * 2 * N * N + which * N * N + i * N + j];
      C[(size_t)row * CS + j] = s;
    }
  }
  for (int e = tid; e < NP * CS; e += kLossThreads) Delta[e] = 0.f;
  load_centered_chunk(x, y, N, D, ch * kChunk, tile);
  __syncthreads();
  float part = 0.f;
  if (J.mode == 0) {
    for (int i = tid; i < 2 * N; i += kLossThreads) {
      const int which = i / N, n = i - which * N;
      nrm[which * NP + n] = fmaxf(sqrtf(fmaxf(C[which * NP * CS + n * CS + n], 0.f)), 1e-12f);
      inv_nrm[which * NP + n] = 1.f / nrm[which * NP + n];
    }
    __syncthreads();
    for (int i = warp; i < N; i += nwarps)
      for (int j = lane; j < N; j += 32) {
        const float g1 = C[i * CS + j] * (inv_nrm[i] * inv_nrm[j]);
        const float g2 = C[NP * CS + i * CS + j] * (inv_nrm[NP + i] * inv_nrm[NP + j]);
        const float d = g1 - g2;
        part += d * d;
        Delta[i * CS + j] = 2.f * coef * d;     // dL/dG1 ; dL/dG2 = -Delta
        C[i * CS + j] = g1;                      // keep the normalised Grams for r_i
        C[NP * CS + i * CS + j] = g2;
      }
  } else {
    for (int i = warp; i < N; i += nwarps)
      for (int j = lane; j < N; j += 32) part += C[i * CS + j] * C[NP * CS + i * CS + j];
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
      for (int j = 0; j < N; ++j) s += Delta[n * CS + j] * C[which * NP * CS + n * CS + j];
      rdot[which * NP + n] = (which ? -2.f : 2.f) * s;
    }
    // E^ = E' / n (in place; each thread owns whole columns of both tiles)
    for (int cc = tid; cc < kChunk; cc += kLossThreads)
      for (int which = 0; which < 2; ++which)
        for (int j = 0; j < N; ++j) tile[((size_t)which * NP + j) * kTS + cc] *= inv_nrm[which * NP + j];
  }
  __syncthreads();
  if (J.dx == nullptr && J.dy == nullptr) return;

  const int MT = NP / 16;
  for (int which = 0; which < 2; ++which) {
    float* dst = which ? (J.dy ? J.dy + (long long)b * N * D : nullptr) : (J.dx ? J.dx + (long long)b * N * D : nullptr);
    if (dst == nullptr) continue;
    const int acc_flag = which ? J.acc_y : J.acc_x;
    const float* A = (J.mode == 0) ? Delta : C + (size_t)(1 - which) * NP * CS;     // [NP][CS]
    const float* Bt = tile + (size_t)which * NP * kTS;                                // [NP][kTS]
    const float scale = (J.mode == 0) ? (which ? -2.f : 2.f) : 2.f * coef;
    const bool pair_ok = ((D & 1) == 0) && ((reinterpret_cast<uintptr_t>(dst) & 7) == 0);
    for (int nt = warp; nt < kChunk / 8; nt += nwarps) {
      float acc[4][4];
#pragma unroll
      for (int m = 0; m < 4; ++m) acc[m][0] = acc[m][1] = acc[m][2] = acc[m][3] = 0.f;
      for (int k0 = 0; k0 < NP; k0 += 8) {
        const float b0 = Bt[(size_t)(k0 + t) * kTS + nt * 8 + g];
        const float b1 = Bt[(size_t)(k0 + t + 4) * kTS + nt * 8 + g];
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          if (m < MT) {
            const float* ar = A + (size_t)(m * 16 + g) * CS + k0 + t;
            mma_tf32x3(acc[m], ar[0], ar[8 * CS], ar[4], ar[8 * CS + 4], b0, b1);
          }
        }
      }
      const int cc = nt * 8 + 2 * t, c = ch * kChunk + cc;
      float v[4][4];
      float cs0 = 0.f, cs1 = 0.f;
#pragma unroll
      for (int m = 0; m < 4; ++m) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {      // h = 0: row g, h = 1: row g + 8
          const int i = m * 16 + g + 8 * h;
          float v0 = scale * acc[m][2 * h], v1 = scale * acc[m][2 * h + 1];
          if (J.mode == 0) {
            if (m < MT && i < N) {
              const float inr = inv_nrm[which * NP + i], rd = rdot[which * NP + i];
              v0 = (v0 - Bt[(size_t)i * kTS + cc] * rd) * inr;
              v1 = (v1 - Bt[(size_t)i * kTS + cc + 1] * rd) * inr;
            } else {
              v0 = v1 = 0.f;
            }
            cs0 += v0;
            cs1 += v1;
          }
          v[m][2 * h] = v0;
          v[m][2 * h + 1] = v1;
        }
      }
      if (J.mode == 0) {      // dE = dE' - column mean(dE') ; rows live in the 8 lanes sharing t
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
          cs0 += __shfl_xor_sync(0xffffffffu, cs0, o);
          cs1 += __shfl_xor_sync(0xffffffffu, cs1, o);
        }
        cs0 /= N;
        cs1 /= N;
      }
#pragma unroll
      for (int m = 0; m < 4; ++m) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int i = m * 16 + g + 8 * h;
          if (m < MT && i < N) {
            const float o0 = v[m][2 * h] - (J.mode == 0 ? cs0 : 0.f), o1 = v[m][2 * h + 1] - (J.mode == 0 ? cs1 : 0.f);
            if (c + 1 < D && pair_ok) {
              emit2(dst, (long long)i * D + c, o0, o1, acc_flag);
            } else {
              if (c < D) emit(dst, (long long)i * D + c, o0, acc_flag);
              if (c + 1 < D) emit(dst, (long long)i * D + c + 1, o1, acc_flag);
            }
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ tensor-centric unit losses
// The three loss terms of one DualVGR unit (train.py:148-154) share their operands: common(ca, cm), HSIC(aq, ca),
// HSIC(mq, cm). The pair-centric kernels above load every operand once per PAIR (6 tile loads per pass, two tiles per
// CTA, atomics where a tensor belongs to two pairs). Here the unit of work is a TENSOR:
//   pass 1 (aux_gram_kernel): centred Gram partials of ca, cm, aq, mq — one tile per CTA, 4 tile loads per pass
//   pass 2 (aux_grad_kernel): one CTA = (video, tensor, 256-column chunk): it sums the partial Grams it needs, rebuilds the
//          N x N algebra, and produces the COMPLETE gradient of its tensor — common part and HSIC part from the same
//          right-operand fragments — with plain stores (no atomics, no zero-filled accumulation buffers).
// tensors: 0 = ca (com_app), 1 = cm (com_motion), 2 = aq (aq_fusion), 3 = mq (mq_fusion)
struct AuxParams {
  const float* x[4];
  float* dx[4];
  float* loss_part;      // [B][3]: coef-scaled common, HSIC(aq, ca), HSIC(mq, cm) of each video
  float coef_com, coef_dep;
  int B, N, D, chunks;
  float* ws;             // [4][B][N][N] centred Grams, accumulated over the column chunks
  int x3;                // Gram products on error-compensated 3 x TF32 (1) or plain TF32 (0)
};

__device__ __forceinline__ void load_centered_tile(const float* __restrict__ x, int N, int D, int c0, float* t) {
  const int NP = round16(N);
  for (int cc = threadIdx.x; cc < kChunk; cc += blockDim.x) {
    const int c = c0 + cc;
    float s = 0.f;
    if (c < D)
      for (int n = 0; n < N; ++n) s += x[(long long)n * D + c];
    const float mean = s / N;
    for (int n = 0; n < N; ++n) t[n * kTS + cc] = (c < D) ? x[(long long)n * D + c] - mean : 0.f;
    for (int n = N; n < NP; ++n) t[n * kTS + cc] = 0.f;
  }
}

__global__ void __launch_bounds__(kLossThreads) aux_gram_kernel(const AuxParams p) {
  extern __shared__ __align__(16) float sm[];
  float* tile = sm;   // [NP16][kTS]
  const int b = blockIdx.x, ts = blockIdx.y, ch = blockIdx.z, N = p.N, D = p.D, NP = round16(N);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = kLossThreads / 32, g = lane >> 2, t = lane & 3;
  load_centered_tile(p.x[ts] + (long long)b * N * D, N, D, ch * kChunk, tile);
  __syncthreads();
  // the column chunks of a (video, tensor) ADD their partial Grams into one N x N block (zeroed by the launcher): pass 2 then
  // reads one Gram per operand instead of re-summing `chunks` partials in every one of its `chunks` CTAs (12x redundant L2
  // traffic at D = 768: ncu showed pass 2 latency-bound on exactly those loads)
  float* out = p.ws + ((long long)ts * p.B + b) * N * N;
  const int MT = NP / 16, NT = NP / 8;
  for (int ot = warp; ot < MT * NT; ot += nwarps) {
    const int mt = ot / NT, nt = ot - mt * NT;
    const float* ar = tile + (size_t)(mt * 16 + g) * kTS + t;
    const float* br = tile + (size_t)(nt * 8 + g) * kTS + t;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (p.x3) {
#pragma unroll 4
      for (int k0 = 0; k0 < kChunk; k0 += 8)
        mma_tf32x3(acc, ar[k0], ar[8 * kTS + k0], ar[k0 + 4], ar[8 * kTS + k0 + 4], br[k0], br[k0 + 4]);
    } else {
#pragma unroll 4
      for (int k0 = 0; k0 < kChunk; k0 += 8)
        mma_tf32(acc, to_tf32(ar[k0]), to_tf32(ar[8 * kTS + k0]), to_tf32(ar[k0 + 4]), to_tf32(ar[8 * kTS + k0 + 4]),
                 to_tf32(br[k0]), to_tf32(br[k0 + 4]));
    }
    const int row = mt * 16 + g, col = nt * 8 + 2 * t;
    if (row < N && col < N) atomicAdd(out + row * N + col, acc[0]);
    if (row < N && col + 1 < N) atomicAdd(out + row * N + col + 1, acc[1]);
    if (row + 8 < N && col < N) atomicAdd(out + (row + 8) * N + col, acc[2]);
    if (row + 8 < N && col + 1 < N) atomicAdd(out + (row + 8) * N + col + 1, acc[3]);
  }
}

// pass 2a (aux_mat_kernel, one CTA per video): the whole N x N algebra ONCE per video. From the four centred Grams
//   C_ca, C_cm, C_aq, C_mq it produces the three loss values and, for every tensor t, ONE matrix M_t such that
//   dL/dE_t = M_t (R E_t)   (R E_t = E_t centred over the nodes):
//     common(ca, cm) = c sum_ij (G1 - G2)^2,  G = C / (n n^T),  n_i = sqrt(C_ii):   A = dL/dG1 = 2 c (G1 - G2) = -dL/dG2
//        dL/dE^_i = 2 sum_j A_ij E^_j ,  r_i = E^_i . dL/dE^_i = 2 sum_j A_ij G_ij ,  dL/d(RE)_i = (dL/dE^_i - r_i E^_i) / n_i
//        => M^com = R [ 2 A / (n n^T) - diag(r / n^2) ]           (the leading R: gradient of the centring)
//     HSIC(x, y) = c' sum_ij Cx_ij Cy_ij:   dL/dE_x = 2 c' Cy (R E_x)  => M^hsic_x = 2 c' Cy
//   M_ca = M^com_1 + 2 c' C_aq,  M_cm = M^com_2 + 2 c' C_mq,  M_aq = 2 c' C_ca,  M_mq = 2 c' C_cm.
// pass 2b (aux_apply_kernel): dE_t = M_t (R E_t), a streaming pass (one thread per feature column keeps the N centred values
//   of its column in registers; M_t sits in shared memory) — plain fp32 FMAs.
// (The first version rebuilt this algebra in every one of the 3 column-chunk CTAs of a (video, tensor) and ran the products on
//  TF32 fragments out of a 57 KB tile: 156 us per unit layer, latency-bound at 7 % of DRAM bandwidth; this one is ~30 us.)
__global__ void __launch_bounds__(256) aux_mat_kernel(const AuxParams p, float* __restrict__ mats) {
  extern __shared__ __align__(16) float sm[];
  const int b = blockIdx.x, N = p.N, tid = threadIdx.x, NN = N * N;
  float* C = sm;                      // [4][N][N] centred Grams: ca, cm, aq, mq
  float* A = C + 4 * NN;              // [N][N] dL/dG_ca
  float* Mc = A + NN;                 // [2][N][N] un-centred common-term matrices of ca, cm
  float* nrm = Mc + 2 * NN;           // [2][N] n_i of ca, cm
  float* rd = nrm + 2 * N;            // [2][N] r_i
  float* colm = rd + 2 * N;           // [2][N] column means of Mc
  __shared__ float red[3][8];
  for (int e = tid; e < 4 * NN; e += blockDim.x) {
    const int ts = e / NN, ij = e - ts * NN;
    C[e] = p.ws[((long long)ts * p.B + b) * NN + ij];
  }
  __syncthreads();
  for (int e = tid; e < 2 * N; e += blockDim.x) {
    const int w = e / N, i = e - w * N;
    nrm[e] = fmaxf(sqrtf(fmaxf(C[w * NN + i * N + i], 0.f)), 1e-12f);
  }
  __syncthreads();
  float part[3] = {0.f, 0.f, 0.f};
  for (int e = tid; e < NN; e += blockDim.x) {
    const int i = e / N, j = e - i * N;
    const float g1 = C[e] / (nrm[i] * nrm[j]), g2 = C[NN + e] / (nrm[N + i] * nrm[N + j]);
    const float d = g1 - g2;
    part[0] += d * d;
    A[e] = 2.f * p.coef_com * d;
    part[1] += C[2 * NN + e] * C[e];              // HSIC(aq, ca)
    part[2] += C[3 * NN + e] * C[NN + e];         // HSIC(mq, cm)
  }
  __syncthreads();
  for (int e = tid; e < 2 * N; e += blockDim.x) {       // r_i = 2 sum_j (+-A_ij) G_ij
    const int w = e / N, i = e - w * N;
    float s_ = 0.f;
    for (int j = 0; j < N; ++j) s_ += A[i * N + j] * C[w * NN + i * N + j] / (nrm[w * N + i] * nrm[w * N + j]);
    rd[e] = (w ? -2.f : 2.f) * s_;
  }
  __syncthreads();
  for (int e = tid; e < 2 * NN; e += blockDim.x) {
    const int w = e / NN, ij = e - w * NN, i = ij / N, j = ij - i * N;
    const float ni = nrm[w * N + i], nj = nrm[w * N + j];
    float m = (w ? -2.f : 2.f) * A[ij] / (ni * nj);
    if (i == j) m -= rd[w * N + i] / (ni * ni);
    Mc[e] = m;
  }
  __syncthreads();
  for (int e = tid; e < 2 * N; e += blockDim.x) {       // column means: the centring projection applied from the left
    const int w = e / N, j = e - w * N;
    float s_ = 0.f;
    for (int i = 0; i < N; ++i) s_ += Mc[w * NN + i * N + j];
    colm[e] = s_ / N;
  }
  __syncthreads();
  float* out = mats + (long long)b * 4 * NN;            // [B][4][N][N]
  for (int e = tid; e < NN; e += blockDim.x) {
    const int j = e % N;
    out[e] = Mc[e] - colm[j] + 2.f * p.coef_dep * C[2 * NN + e];
    out[NN + e] = Mc[NN + e] - colm[N + j] + 2.f * p.coef_dep * C[3 * NN + e];
    out[2 * NN + e] = 2.f * p.coef_dep * C[e];
    out[3 * NN + e] = 2.f * p.coef_dep * C[NN + e];
  }
  const int warp = tid >> 5, lane = tid & 31;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float v = warp_sum(part[c]);
    if (lane == 0) red[c][warp] = v;
  }
  __syncthreads();
  if (tid < 3) {
    float s_ = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s_ += red[tid][w];
    p.loss_part[(long long)b * 3 + tid] = (tid == 0 ? p.coef_com : p.coef_dep) * s_;
  }
}

template <int NP>
__global__ void __launch_bounds__(128) aux_apply_kernel(const AuxParams p, const float* __restrict__ mats) {
  extern __shared__ __align__(16) float sm[];           // M_t [N][NP] (rows zero-padded to NP: 16-byte broadcast loads)
  const int b = blockIdx.x, ts = blockIdx.y, N = p.N, D = p.D, NN = N * N;
  if (p.dx[ts] == nullptr) return;
  for (int e = threadIdx.x; e < N * NP; e += blockDim.x) {
    const int i = e / NP, j = e - i * NP;
    sm[e] = j < N ? mats[((long long)b * 4 + ts) * NN + i * N + j] : 0.f;
  }
  __syncthreads();
  const float* __restrict__ x = p.x[ts] + (long long)b * N * D;
  float* __restrict__ dx = p.dx[ts] + (long long)b * N * D;
  // two adjacent feature columns per thread (8-byte accesses): every 16-byte load of a matrix row feeds 8 FMAs
  for (int c = 2 * (blockIdx.z * blockDim.x + threadIdx.x); c < D; c += 2 * gridDim.z * blockDim.x) {
    const bool two = c + 1 < D;
    float e0[NP], e1[NP];
    float m0 = 0.f, m1 = 0.f;
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      e0[j] = e1[j] = 0.f;
      if (j < N) {
        if (two) {
          const float2 v = *reinterpret_cast<const float2*>(x + (long long)j * D + c);
          e0[j] = v.x; e1[j] = v.y;
        } else {
          e0[j] = x[(long long)j * D + c];
        }
      }
      m0 += e0[j]; m1 += e1[j];
    }
    m0 /= N; m1 /= N;
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      e0[j] = j < N ? e0[j] - m0 : 0.f;
      e1[j] = j < N ? e1[j] - m1 : 0.f;
    }
    for (int i = 0; i < N; ++i) {
      float a0 = 0.f, a1 = 0.f;
      const float4* row = reinterpret_cast<const float4*>(sm + i * NP);
#pragma unroll
      for (int j4 = 0; j4 < NP / 4; ++j4) {
        const float4 m = row[j4];
        a0 = fmaf(m.x, e0[4 * j4], a0); a1 = fmaf(m.x, e1[4 * j4], a1);
        a0 = fmaf(m.y, e0[4 * j4 + 1], a0); a1 = fmaf(m.y, e1[4 * j4 + 1], a1);
        a0 = fmaf(m.z, e0[4 * j4 + 2], a0); a1 = fmaf(m.z, e1[4 * j4 + 2], a1);
        a0 = fmaf(m.w, e0[4 * j4 + 3], a0); a1 = fmaf(m.w, e1[4 * j4 + 3], a1);
      }
      if (two) *reinterpret_cast<float2*>(dx + (long long)i * D + c) = make_float2(a0, a1);
      else dx[(long long)i * D + c] = a0;
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
  const size_t smem1 = (size_t)2 * round16(N) * kTS * sizeof(float);
  const size_t smem2 = pair_grad_smem(N);
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

extern "C" long long dvgr_aux_loss_workspace(int B, int N, int D) {
  (void)D;
  return 8LL * B * N * N;          // centred Grams [4][B][N][N] + gradient matrices [B][4][N][N]
}

extern "C" int dvgr_aux_loss_unit(const float* ca, const float* cm, const float* aq, const float* mq, float coef_com,
                                  float coef_dep, int B, int N, int D, float* d_ca, float* d_cm, float* d_aq, float* d_mq,
                                  float* loss_part, float* gram_ws, void* stream) {
  return dvgr_aux_loss_unit_ex(ca, cm, aq, mq, coef_com, coef_dep, B, N, D, d_ca, d_cm, d_aq, d_mq, loss_part, gram_ws, 0, stream);
}

extern "C" int dvgr_aux_loss_unit_ex(const float* ca, const float* cm, const float* aq, const float* mq, float coef_com,
                                     float coef_dep, int B, int N, int D, float* d_ca, float* d_cm, float* d_aq, float* d_mq,
                                     float* loss_part, float* gram_ws, int precise, void* stream) {
  if (B <= 0) return 0;
  if (N < 1 || N > 64) return set_error("aux_loss: N=%d out of [1,64]", N);
  if (!ca || !cm || !aq || !mq || !loss_part || !gram_ws) return set_error("aux_loss: null buffer");
  AuxParams p;
  memset(&p, 0, sizeof(p));
  p.x[0] = ca; p.x[1] = cm; p.x[2] = aq; p.x[3] = mq;
  p.dx[0] = d_ca; p.dx[1] = d_cm; p.dx[2] = d_aq; p.dx[3] = d_mq;
  p.loss_part = loss_part; p.coef_com = coef_com; p.coef_dep = coef_dep;
  p.B = B; p.N = N; p.D = D; p.chunks = (D + kChunk - 1) / kChunk; p.ws = gram_ws;
  p.x3 = precise ? 1 : 0;
  const size_t smem1 = (size_t)round16(N) * kTS * sizeof(float);
  const size_t smem2 = (size_t)(7 * N * N + 6 * N) * sizeof(float);
  static size_t conf1 = 0, conf2 = 0;
  if (smem1 > 48 * 1024 && smem1 > conf1) {
    cudaError_t e = cudaFuncSetAttribute(aux_gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1);
    if (e != cudaSuccess) return set_error("aux_loss: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    conf1 = smem1;
  }
  if (smem2 > 48 * 1024 && smem2 > conf2) {
    cudaError_t e = cudaFuncSetAttribute(aux_mat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
    if (e != cudaSuccess) return set_error("aux_loss: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    conf2 = smem2;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid(B, 4, p.chunks);
  if (cudaMemsetAsync(gram_ws, 0, sizeof(float) * 4 * (size_t)B * N * N, st) != cudaSuccess)
    return set_error("aux_loss: cudaMemsetAsync failed");
  aux_gram_kernel<<<grid, kLossThreads, smem1, st>>>(p);
  DVGR_CHECK_LAUNCH("aux_gram");
  float* mats = gram_ws + 4LL * B * N * N;                       // [B][4][N][N], behind the Grams
  aux_mat_kernel<<<B, 256, smem2, st>>>(p, mats);
  DVGR_CHECK_LAUNCH("aux_mat");
  if (d_ca || d_cm || d_aq || d_mq) {
    if ((D & 1) != 0) return set_error("aux_loss: D=%d must be even", D);
    const int zc = (D + 255) / 256 < 3 ? (D + 255) / 256 : 3;     // column chunks per (video, tensor)
    const int NPad = N <= 32 ? 32 : 64;
    const size_t smem3 = (size_t)N * NPad * sizeof(float);
    if (N <= 32) aux_apply_kernel<32><<<dim3(B, 4, zc), 128, smem3, st>>>(p, mats);
    else aux_apply_kernel<64><<<dim3(B, 4, zc), 128, smem3, st>>>(p, mats);
    DVGR_CHECK_LAUNCH("aux_apply");
  }
  return 0;
}
