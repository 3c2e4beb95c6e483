// Video-based multi-view graph attention, fused per (video, graph) — forward and backward.
//
// Replaces, for all heads of one punishGAT at once, reference model/GraphNN.py:98-111 (+ the dropouts / concat of
// :175-177): pairwise logits, LeakyReLU, adjacency mask, softmax, attention dropout, gate, neighbour aggregation, ELU,
// head concat, output dropout. The [B,N,N,2*Dh] tensor of GraphNN.py:115-155 never exists: a.[Wh_i;Wh_j] + c is
// evaluated as a1.Wh_i + a2.Wh_j + c.
//
// One CTA per (video b, graph g). The node block Wh_b [N, D] (bf16, produced by the tcgen05 projection GEMM) is staged
// in shared memory once and reused by the logit dots, and by the N x N aggregation of every head.
// HBM traffic per CTA: read N*D bf16 + N gates, write N*D bf16  -> the kernel is HBM-bound by design.
#include <cuda_bf16.h>

#include "capi_internal.h"
#include "ptx.cuh"
#include "rng.cuh"

namespace dvgr {

constexpr int kGatThreads = 384;
constexpr int kMaxHeads = 4;

struct GatGraph {
  const __nv_bfloat16* wh;    // [B*N, ld_wh] projected node features (bias included)
  const float* gate;          // [B, N] query-punishment gate of the stream this graph reads
  const float* avec;          // [heads][2*Dh + 1] : a1 | a2 | c
  __nv_bfloat16* out;         // fwd: [B*N, ld_out] ; bwd: h' of the forward pass (post-dropout)
  float* out_f32;             // fwd, optional: dense [B*N, D] fp32 copy of `out` for the auxiliary losses
  const __nv_bfloat16* dout;  // bwd: gradient of `out`
  const float* dout_f32;      // bwd, optional: extra gradient arriving on the fp32 copy (auxiliary losses)
  __nv_bfloat16* dwh;         // bwd: [B*N, ld_wh] gradient of wh
  float* dgate;               // bwd: [B, N] this graph's contribution to the gate gradient
  float* davec;               // bwd: [B][heads][2*Dh + 1] per-video partial sums (reduced by dvgr_colsum)
  unsigned int drop_stream;   // dropout stream id of this graph (attention: +0, output: +1)
};

struct GatParams {
  GatGraph g[4];
  int B, N, D, heads;
  long long ld_wh, ld_out;
  const float* adj;           // [N, N]
  float slope;                // LeakyReLU negative slope (0.01)
  float p_att, p_out;         // dropout probabilities (0 in eval)
  unsigned long long seed;
  const unsigned long long* seed_off;
};

struct GatSmem {
  __nv_bfloat16* wh;   // [N][D]
  float* P;            // [heads][N][NP]   probabilities (fwd: already gated + dropped)
  float* s;            // [heads][N]
  float* t;            // [heads][N]
  float* gate;         // [N]
  float* avec;         // [heads][2*Dh+1]
  unsigned char* adj;  // [N][N]
};

__host__ __device__ inline int round4(int n) { return (n + 3) & ~3; }

__host__ __device__ inline size_t gat_smem_common(int N, int D, int heads) {
  const int NP = round4(N);
  size_t b = (size_t)N * D * 2;                              // wh
  b += (size_t)heads * N * NP * 4;                           // P
  b += (size_t)heads * N * 4 * 2;                            // s, t
  b += (size_t)round4(N) * 4;                                // gate
  b += (size_t)heads * (2 * (D / heads) + 1) * 4 + 16;       // avec
  b += (size_t)round4(N * N);                                // adj
  return (b + 15) & ~(size_t)15;
}

__device__ __forceinline__ GatSmem carve(unsigned char* base, int N, int D, int heads) {
  const int NP = round4(N);
  GatSmem sm;
  sm.wh = reinterpret_cast<__nv_bfloat16*>(base);
  base += (size_t)N * D * 2;
  sm.P = reinterpret_cast<float*>(base);
  base += (size_t)heads * N * NP * 4;
  sm.s = reinterpret_cast<float*>(base);
  base += (size_t)heads * N * 4;
  sm.t = reinterpret_cast<float*>(base);
  base += (size_t)heads * N * 4;
  sm.gate = reinterpret_cast<float*>(base);
  base += (size_t)round4(N) * 4;
  sm.avec = reinterpret_cast<float*>(base);
  base += ((size_t)heads * (2 * (D / heads) + 1) * 4 + 15) & ~(size_t)15;
  sm.adj = base;
  return sm;
}

// Stage Wh tile, gate, a-vectors, adjacency; compute s_i = a1.Wh_i, t_j = a2.Wh_j per head.
__device__ __forceinline__ void gat_stage(const GatParams& p, const GatGraph& gr, int b, const GatSmem& sm) {
  const int N = p.N, D = p.D, K = p.heads, Dh = D / K;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const __nv_bfloat16* src = gr.wh + (long long)b * N * p.ld_wh;
  const int vec_per_row = D / 8;
  for (int v = tid; v < N * vec_per_row; v += nthr) {
    const int r = v / vec_per_row, c = v - r * vec_per_row;
    reinterpret_cast<uint4*>(sm.wh)[v] = *reinterpret_cast<const uint4*>(src + (long long)r * p.ld_wh + c * 8);
  }
  for (int i = tid; i < N; i += nthr) sm.gate[i] = gr.gate[(long long)b * N + i];
  for (int i = tid; i < K * (2 * Dh + 1); i += nthr) sm.avec[i] = gr.avec[i];
  for (int i = tid; i < N * N; i += nthr) sm.adj[i] = p.adj[i] > 0.f ? 1 : 0;
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31, nwarps = nthr >> 5;
  for (int pr = warp; pr < N * K; pr += nwarps) {
    const int i = pr / K, k = pr - i * K;
    const float* a = sm.avec + k * (2 * Dh + 1);
    const __nv_bfloat16* w = sm.wh + i * D + k * Dh;
    float s = 0.f, t = 0.f;
    for (int c = lane; c < Dh; c += 32) {
      const float x = __bfloat162float(w[c]);
      s += a[c] * x;
      t += a[Dh + c] * x;
    }
    s = warp_sum(s);
    t = warp_sum(t);
    if (lane == 0) {
      sm.s[k * N + i] = s;
      sm.t[k * N + i] = t;
    }
  }
  __syncthreads();
}

// Row softmax of head k, node i (one warp). Leaves P[k][i][j] = softmax_j(e_ij) in smem; padded columns are zero.
__device__ __forceinline__ void gat_softmax_row(const GatParams& p, const GatSmem& sm, int k, int i, int lane) {
  const int N = p.N, NP = round4(N), Dh = p.D / p.heads;
  const float c = sm.avec[k * (2 * Dh + 1) + 2 * Dh];
  const float si = sm.s[k * N + i];
  float e[2];
  float m = -INFINITY;
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int j = lane + 32 * q;
    e[q] = -INFINITY;
    if (j < N) {
      float u = si + sm.t[k * N + j] + c;
      u = u > 0.f ? u : p.slope * u;
      e[q] = sm.adj[i * N + j] ? u : -9e15f;
      m = fmaxf(m, e[q]);
    }
  }
  m = warp_max(m);
  float sum = 0.f;
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int j = lane + 32 * q;
    if (j < N) {
      e[q] = __expf(e[q] - m);
      sum += e[q];
    }
  }
  sum = warp_sum(sum);
  const float inv = 1.f / sum;
  float* Prow = sm.P + ((size_t)k * N + i) * NP;
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int j = lane + 32 * q;
    if (j < N) Prow[j] = e[q] * inv;
    else if (j < NP) Prow[j] = 0.f;
  }
}

// Dropout stream layout (any bijection works as long as forward and backward agree; these make ONE Philox call serve
// eight mask elements where the kernels consume them):
//   attention mask (b, k, i, j)  : murmur-hashed element index ((b*K + k)*N + i)*N + j (rng.cuh: dropout_scale1_hash)
//   output mask    (b, i, c)     : call ((b*(D/2) + c/2) * ceil(N/4) + i/4 , element (i%4)*2 + (c&1)
__device__ __forceinline__ float att_keep(const GatParams& p, const DropoutCfg& cfg, int b, int k, int i, int j) {
  return dropout_scale1_hash(cfg, (((unsigned long long)b * p.heads + k) * p.N + i) * p.N + j);
}
__device__ __forceinline__ void out_keep8(const GatParams& p, const DropoutCfg& cfg, int b, int pair, int rowquad,
                                          float (&sc)[8]) {
  const unsigned long long call = ((unsigned long long)b * (p.D >> 1) + pair) * ((p.N + 3) >> 2) + rowquad;
  dropout_scale8(cfg, call, sc);
}

// ------------------------------------------------------------------------------------------------------ forward
__global__ void __launch_bounds__(kGatThreads) gat_attn_fwd_kernel(const GatParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int b = blockIdx.x;
  const GatGraph& gr = p.g[blockIdx.y];
  const int N = p.N, D = p.D, K = p.heads, Dh = D / K, NP = round4(N);
  const GatSmem sm = carve(smem_raw, N, D, K);
  gat_stage(p, gr, b, sm);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
  const DropoutCfg datt{p.seed, gr.drop_stream, p.p_att, p.seed_off};
  for (int pr = warp; pr < N * K; pr += nwarps) {
    const int k = pr / N, i = pr - k * N;
    gat_softmax_row(p, sm, k, i, lane);
    __syncwarp();
    float* Prow = sm.P + ((size_t)k * N + i) * NP;
    for (int j = lane; j < N; j += 32)
      Prow[j] = Prow[j] * sm.gate[j] * att_keep(p, datt, b, k, i, j);   // gate multiplies the values: fold it into P
  }
  __syncthreads();

  // aggregation: out[i][c] = dropout(ELU(sum_j P[k(c)][i][j] * Wh[j][c])) ; 2 columns x 8 rows per thread and pass
  const DropoutCfg dout{p.seed, gr.drop_stream + 1u, p.p_out, p.seed_off};
  __nv_bfloat16* outp = gr.out + (long long)b * N * p.ld_out;
  const __nv_bfloat162* wh2 = reinterpret_cast<const __nv_bfloat162*>(sm.wh);
  for (int pair = tid; pair < D / 2; pair += blockDim.x) {
    const int c = pair * 2;
    const int k = c / Dh;
    const float* Pk = sm.P + (size_t)k * N * NP;
    for (int i0 = 0; i0 < N; i0 += 8) {
      float acc[8][2];
#pragma unroll
      for (int r = 0; r < 8; ++r) acc[r][0] = acc[r][1] = 0.f;
      for (int j0 = 0; j0 < N; j0 += 4) {
        float2 w[4];
#pragma unroll
        for (int q = 0; q < 4; ++q)
          w[q] = (j0 + q < N) ? __bfloat1622float2(wh2[(size_t)(j0 + q) * (D / 2) + pair]) : make_float2(0.f, 0.f);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          if (i0 + r < N) {
            const float4 pv = *reinterpret_cast<const float4*>(Pk + (size_t)(i0 + r) * NP + j0);
            acc[r][0] += pv.x * w[0].x + pv.y * w[1].x + pv.z * w[2].x + pv.w * w[3].x;
            acc[r][1] += pv.x * w[0].y + pv.y * w[1].y + pv.z * w[2].y + pv.w * w[3].y;
          }
        }
      }
      float keep[2][8];
      if (p.p_out > 0.f) {
        out_keep8(p, dout, b, pair, i0 >> 2, keep[0]);
        out_keep8(p, dout, b, pair, (i0 >> 2) + 1, keep[1]);
      }
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int i = i0 + r;
        if (i < N) {
          float o0 = eluf_(acc[r][0]), o1 = eluf_(acc[r][1]);
          if (p.p_out > 0.f) {
            o0 *= keep[r >> 2][(r & 3) * 2];
            o1 *= keep[r >> 2][(r & 3) * 2 + 1];
          }
          *reinterpret_cast<__nv_bfloat162*>(outp + (long long)i * p.ld_out + c) = __floats2bfloat162_rn(o0, o1);
          if (gr.out_f32 != nullptr)
            *reinterpret_cast<float2*>(gr.out_f32 + ((long long)b * N + i) * D + c) = make_float2(o0, o1);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------ backward
// extra shared memory: dz [N][D] bf16, dP [heads][N][NP] f32 (dP~ -> du -> transposed P~), ds/dt [heads][N], dg [N]
__host__ __device__ inline size_t gat_smem_bwd_extra(int N, int D, int heads) {
  const int NP = round4(N);
  return (size_t)N * D * 2 + (size_t)heads * N * NP * 4 + (size_t)heads * N * 4 * 2 + (size_t)round4(N) * 4 + 16;
}

__global__ void __launch_bounds__(kGatThreads) gat_attn_bwd_kernel(const GatParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int b = blockIdx.x;
  const GatGraph& gr = p.g[blockIdx.y];
  const int N = p.N, D = p.D, K = p.heads, Dh = D / K, NP = round4(N);
  const GatSmem sm = carve(smem_raw, N, D, K);
  unsigned char* ext = smem_raw + gat_smem_common(N, D, K);
  __nv_bfloat16* dz = reinterpret_cast<__nv_bfloat16*>(ext);
  ext += (size_t)N * D * 2;
  float* dP = reinterpret_cast<float*>(ext);
  ext += (size_t)K * N * NP * 4;
  float* ds = reinterpret_cast<float*>(ext);
  ext += (size_t)K * N * 4;
  float* dt = reinterpret_cast<float*>(ext);
  ext += (size_t)K * N * 4;
  float* dg = reinterpret_cast<float*>(ext);
  __shared__ float dc_s[kMaxHeads];

  gat_stage(p, gr, b, sm);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
  const DropoutCfg datt{p.seed, gr.drop_stream, p.p_att, p.seed_off};
  const DropoutCfg dout{p.seed, gr.drop_stream + 1u, p.p_out, p.seed_off};

  // 1. probabilities (pre-dropout, ungated); dz = (dout [+ dout_f32]) * mask * ELU'(z), 4 rows x 2 columns per unit
  for (int pr = warp; pr < N * K; pr += nwarps) gat_softmax_row(p, sm, pr / N, pr % N, lane);
  for (int i = tid; i < N; i += blockDim.x) dg[i] = 0.f;
  {
    const __nv_bfloat16* hp = gr.out + (long long)b * N * p.ld_out;
    const __nv_bfloat16* dop = gr.dout + (long long)b * N * p.ld_out;
    const float keepf = 1.f - p.p_out;
    const int quads = (N + 3) >> 2;
    for (int u = tid; u < quads * (D / 2); u += blockDim.x) {
      const int rq = u / (D / 2), pair = u - rq * (D / 2), c = pair * 2;
      float keep[8];
      if (p.p_out > 0.f) out_keep8(p, dout, b, pair, rq, keep);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int i = rq * 4 + r;
        if (i >= N) break;
        float2 h = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(hp + (long long)i * p.ld_out + c));
        float2 d = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(dop + (long long)i * p.ld_out + c));
        if (gr.dout_f32 != nullptr) {
          const float2 e = *reinterpret_cast<const float2*>(gr.dout_f32 + ((long long)b * N + i) * D + c);
          d.x += e.x;
          d.y += e.y;
        }
        if (p.p_out > 0.f) {
          d.x *= keep[2 * r];
          d.y *= keep[2 * r + 1];
          h.x *= keepf;   // recover ELU(z) of the kept elements (dropped ones have zero gradient anyway)
          h.y *= keepf;
        }
        d.x *= elu_grad_from_out(h.x);
        d.y *= elu_grad_from_out(h.y);
        *reinterpret_cast<__nv_bfloat162*>(dz + (size_t)i * D + c) = __floats2bfloat162_rn(d.x, d.y);
      }
    }
  }
  __syncthreads();

  // 2. dP[k][i][j] = mask_ij * g_j * sum_{c in head k} dz[i][c] * Wh[j][c]   (one warp per (k, i), lanes over c)
  for (int pr = warp; pr < N * K; pr += nwarps) {
    const int k = pr / N, i = pr - k * N;
    float zreg[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int c = lane + 32 * q;
      zreg[q] = c < Dh ? __bfloat162float(dz[(size_t)i * D + k * Dh + c]) : 0.f;
    }
    float mine[2] = {0.f, 0.f};
    for (int j = 0; j < N; ++j) {
      float acc = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int c = lane + 32 * q;
        if (c < Dh) acc += zreg[q] * __bfloat162float(sm.wh[(size_t)j * D + k * Dh + c]);
      }
      acc = warp_sum(acc);
      if ((j & 31) == lane) mine[j >> 5] = acc;
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int j = lane + 32 * q;
      if (j < N) dP[((size_t)k * N + i) * NP + j] = mine[q] * sm.gate[j] * att_keep(p, datt, b, k, i, j);
    }
  }
  __syncthreads();

  // 3. softmax / LeakyReLU backward -> du (in place), ds_i = sum_j du_ij
  for (int pr = warp; pr < N * K; pr += nwarps) {
    const int k = pr / N, i = pr - k * N;
    const float* Prow = sm.P + ((size_t)k * N + i) * NP;
    float* dProw = dP + ((size_t)k * N + i) * NP;
    const float cb = sm.avec[k * (2 * Dh + 1) + 2 * Dh];
    float dot = 0.f;
    for (int j = lane; j < N; j += 32) dot += Prow[j] * dProw[j];
    dot = warp_sum(dot);
    float srow = 0.f;
    for (int j = lane; j < N; j += 32) {
      float de = Prow[j] * (dProw[j] - dot);
      if (!sm.adj[i * N + j]) de = 0.f;
      const float u = sm.s[k * N + i] + sm.t[k * N + j] + cb;
      const float du = de * (u > 0.f ? 1.f : p.slope);
      dProw[j] = du;
      srow += du;
    }
    srow = warp_sum(srow);
    if (lane == 0) ds[k * N + i] = srow;
  }
  __syncthreads();
  for (int pr = tid; pr < N * K; pr += blockDim.x) {      // dt_j = sum_i du_ij
    const int k = pr / N, j = pr - k * N;
    float a = 0.f;
    for (int i = 0; i < N; ++i) a += dP[((size_t)k * N + i) * NP + j];
    dt[k * N + j] = a;
  }
  for (int k = warp; k < K; k += nwarps) {                // dc_k = sum_i ds_i
    float a = 0.f;
    for (int i = lane; i < N; i += 32) a += ds[k * N + i];
    a = warp_sum(a);
    if (lane == 0) dc_s[k] = a;
  }
  __syncthreads();
  // du is consumed: reuse its buffer for the TRANSPOSED P~[k][j][i] = P_ij * mask_ij (ungated), zero padded along i
  for (int e = tid; e < K * N * NP; e += blockDim.x) {
    const int k = e / (N * NP), r = e - k * N * NP, j = r / NP, i = r - j * NP;
    dP[e] = (i < N) ? sm.P[((size_t)k * N + i) * NP + j] * att_keep(p, datt, b, k, i, j) : 0.f;
  }
  __syncthreads();

  // 4. dV_j[c] = sum_i P~_ij dz_i[c] ; dWh_j = g_j dV_j + ds_j a1 + dt_j a2 ; dgate_j ; da1, da2   (8 j x 2 columns / pass)
  const __nv_bfloat162* wh2 = reinterpret_cast<const __nv_bfloat162*>(sm.wh);
  const __nv_bfloat162* dz2 = reinterpret_cast<const __nv_bfloat162*>(dz);
  __nv_bfloat16* dwhp = gr.dwh + (long long)b * N * p.ld_wh;
  float* dav = gr.davec + ((long long)b * K) * (2 * Dh + 1);
  for (int pair0 = 0; pair0 < D / 2; pair0 += blockDim.x) {
    const int pair = pair0 + tid;
    const bool active = pair < D / 2;
    const int c = active ? pair * 2 : 0;
    const int k = c / Dh, cl = c - k * Dh;
    const float* a1 = sm.avec + k * (2 * Dh + 1);
    const float* a2 = a1 + Dh;
    const float* Ptk = dP + (size_t)k * N * NP;
    float da1x = 0.f, da1y = 0.f, da2x = 0.f, da2y = 0.f;
    for (int j0 = 0; j0 < N; j0 += 8) {
      float acc[8][2];
#pragma unroll
      for (int r = 0; r < 8; ++r) acc[r][0] = acc[r][1] = 0.f;
      if (active) {
        for (int i0 = 0; i0 < N; i0 += 4) {
          float2 z[4];
#pragma unroll
          for (int q = 0; q < 4; ++q)
            z[q] = (i0 + q < N) ? __bfloat1622float2(dz2[(size_t)(i0 + q) * (D / 2) + pair]) : make_float2(0.f, 0.f);
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            if (j0 + r < N) {
              const float4 pv = *reinterpret_cast<const float4*>(Ptk + (size_t)(j0 + r) * NP + i0);
              acc[r][0] += pv.x * z[0].x + pv.y * z[1].x + pv.z * z[2].x + pv.w * z[3].x;
              acc[r][1] += pv.x * z[0].y + pv.y * z[1].y + pv.z * z[2].y + pv.w * z[3].y;
            }
          }
        }
      }
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int j = j0 + r;
        if (j >= N) break;                                  // uniform across the block
        const float2 w = active ? __bfloat1622float2(wh2[(size_t)j * (D / 2) + pair]) : make_float2(0.f, 0.f);
        float part = acc[r][0] * w.x + acc[r][1] * w.y;     // gate gradient: dg_j += sum_c dV_j[c] Wh_j[c]
        part = warp_sum(part);
        if (lane == 0) atomicAdd(&dg[j], part);
        if (active) {
          const float gj = sm.gate[j], dsj = ds[k * N + j], dtj = dt[k * N + j];
          const float ox = gj * acc[r][0] + dsj * a1[cl] + dtj * a2[cl];
          const float oy = gj * acc[r][1] + dsj * a1[cl + 1] + dtj * a2[cl + 1];
          *reinterpret_cast<__nv_bfloat162*>(dwhp + (long long)j * p.ld_wh + c) = __floats2bfloat162_rn(ox, oy);
          da1x += dsj * w.x; da1y += dsj * w.y;
          da2x += dtj * w.x; da2y += dtj * w.y;
        }
      }
    }
    if (active) {
      float* o = dav + k * (2 * Dh + 1);
      o[cl] = da1x; o[cl + 1] = da1y;
      o[Dh + cl] = da2x; o[Dh + cl + 1] = da2y;
    }
  }
  __syncthreads();
  for (int i = tid; i < N; i += blockDim.x) gr.dgate[(long long)b * N + i] = dg[i];
  if (tid < K) dav[tid * (2 * Dh + 1) + 2 * Dh] = dc_s[tid];
}

}  // namespace dvgr

using namespace dvgr;

static int fill_gat(GatParams& p, const dvgr_gat_args* a, bool bwd) {
  if (!a) return set_error("gat: null args");
  if (a->n_graphs < 1 || a->n_graphs > 4) return set_error("gat: n_graphs=%d out of [1,4]", a->n_graphs);
  if (a->heads < 1 || a->heads > kMaxHeads) return set_error("gat: heads=%d out of [1,%d]", a->heads, kMaxHeads);
  if (a->B <= 0 || a->N <= 0) return 0;
  if (a->N > 64) return set_error("gat: N=%d > 64 nodes per video is not supported", a->N);
  if (a->D % (a->heads * 2) != 0 || a->D % 8 != 0) return set_error("gat: D=%d must be a multiple of 8 and of 2*heads", a->D);
  if ((a->D / a->heads) > 256) return set_error("gat: head dim %d > 256", a->D / a->heads);
  memset(&p, 0, sizeof(p));
  p.B = a->B; p.N = a->N; p.D = a->D; p.heads = a->heads;
  p.ld_wh = a->ld_wh; p.ld_out = a->ld_out;
  p.adj = a->adj; p.slope = a->slope; p.p_att = a->p_att; p.p_out = a->p_out; p.seed = a->seed;
  p.seed_off = seed_offset_ptr();
  for (int i = 0; i < a->n_graphs; ++i) {
    const dvgr_gat_graph& s = a->graphs[i];
    GatGraph& d = p.g[i];
    if (!s.wh || !s.gate || !s.avec || !s.out) return set_error("gat: graph %d has a null buffer", i);
    d.wh = reinterpret_cast<const __nv_bfloat16*>(s.wh);
    d.gate = s.gate;
    d.avec = s.avec;
    d.out = reinterpret_cast<__nv_bfloat16*>(s.out);
    d.drop_stream = s.drop_stream;
    d.out_f32 = s.out_f32;
    if (bwd) {
      d.dout_f32 = s.dout_f32;
      if (!s.dout || !s.dwh || !s.dgate || !s.davec) return set_error("gat bwd: graph %d has a null gradient buffer", i);
      d.dout = reinterpret_cast<const __nv_bfloat16*>(s.dout);
      d.dwh = reinterpret_cast<__nv_bfloat16*>(s.dwh);
      d.dgate = s.dgate;
      d.davec = s.davec;
    }
  }
  return 0;
}

extern "C" int dvgr_gat_attn_fwd(const dvgr_gat_args* a, void* stream) {
  GatParams p;
  if (int rc = fill_gat(p, a, false)) return rc;
  if (a->B <= 0 || a->N <= 0) return 0;
  const size_t smem = gat_smem_common(p.N, p.D, p.heads);
  if (smem > 227 * 1024) return set_error("gat fwd: %zu bytes of shared memory needed", smem);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(gat_attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error("gat fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured = smem;
  }
  gat_attn_fwd_kernel<<<dim3(p.B, a->n_graphs), kGatThreads, smem, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  DVGR_CHECK_LAUNCH("gat_attn_fwd");
  return 0;
}

extern "C" int dvgr_gat_attn_bwd(const dvgr_gat_args* a, void* stream) {
  GatParams p;
  if (int rc = fill_gat(p, a, true)) return rc;
  if (a->B <= 0 || a->N <= 0) return 0;
  const size_t smem = gat_smem_common(p.N, p.D, p.heads) + gat_smem_bwd_extra(p.N, p.D, p.heads);
  if (smem > 227 * 1024) return set_error("gat bwd: %zu bytes of shared memory needed", smem);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(gat_attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error("gat bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured = smem;
  }
  gat_attn_bwd_kernel<<<dim3(p.B, a->n_graphs), kGatThreads, smem, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  DVGR_CHECK_LAUNCH("gat_attn_bwd");
  return 0;
}
