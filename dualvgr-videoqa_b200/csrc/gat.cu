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
#include <stdlib.h>
#include <string.h>

#include "act.cuh"
#include "capi_internal.h"
#include "ptx.cuh"
#include "rng.cuh"

// (compiled twice, see act.cuh; the tensor-core fast path exists in the bf16 build only)
namespace dvgr {
namespace DVGR_VNS {
using namespace act;
constexpr int kEsz = (int)sizeof(act_t);

constexpr int kGatThreads = 384;
constexpr int kMaxHeads = 4;

struct GatGraph {
  const act_t* wh;    // [B*N, ld_wh] projected node features (bias included)
  const float* gate;          // [B, N] query-punishment gate of the stream this graph reads
  const float* avec;          // [heads][2*Dh + 1] : a1 | a2 | c
  act_t* out;         // fwd: [B*N, ld_out] ; bwd: h' of the forward pass (post-dropout)
  float* out_f32;             // fwd, optional: dense [B*N, D] fp32 copy of `out` for the auxiliary losses
  const act_t* dout;  // bwd: gradient of `out`
  const float* dout_f32;      // bwd, optional: extra gradient arriving on the fp32 copy (auxiliary losses)
  act_t* dwh;         // bwd: [B*N, ld_wh] gradient of wh
  float* dgate;               // bwd: [B, N] this graph's contribution to the gate gradient
  float* davec;               // bwd: [B][heads][2*Dh + 1] per-video partial sums (reduced by dvgr_colsum)
  unsigned int drop_stream;   // dropout stream id of this graph (attention: +0, output: +1)
  // head-split launches (one virtual single-head graph per head, see launch_head_split): position of this slice in the real
  // graph, so that the dropout masks and the per-video partials land where the whole-graph kernel puts them
  int mask_k0, mask_pair0;    // first head / first column pair of this slice
  long long ld_davec;         // elements per video in davec (0: heads * (2*Dh + 1))
  int atomic_dgate;           // accumulate the gate gradient with atomics (the heads of a graph add into one buffer)
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
  int mask_heads, mask_D;     // head count / width of the REAL graph in the dropout-mask indices (0: heads / D)
};

struct GatSmem {
  act_t* wh;   // [N][D]
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
  size_t b = (size_t)N * D * kEsz;                           // wh
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
  sm.wh = reinterpret_cast<act_t*>(base);
  base += (size_t)N * D * kEsz;
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
  const act_t* src = gr.wh + (long long)b * N * p.ld_wh;
  constexpr int kEpv = 16 / kEsz;                            // elements per 16-byte vector
  const int vec_per_row = D / kEpv;
  for (int v = tid; v < N * vec_per_row; v += nthr) {
    const int r = v / vec_per_row, c = v - r * vec_per_row;
    reinterpret_cast<uint4*>(sm.wh)[v] = *reinterpret_cast<const uint4*>(src + (long long)r * p.ld_wh + c * kEpv);
  }
  for (int i = tid; i < N; i += nthr) sm.gate[i] = gr.gate[(long long)b * N + i];
  for (int i = tid; i < K * (2 * Dh + 1); i += nthr) sm.avec[i] = gr.avec[i];
  for (int i = tid; i < N * N; i += nthr) sm.adj[i] = p.adj[i] > 0.f ? 1 : 0;
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31, nwarps = nthr >> 5;
  for (int pr = warp; pr < N * K; pr += nwarps) {
    const int i = pr / K, k = pr - i * K;
    const float* a = sm.avec + k * (2 * Dh + 1);
    const act_t* w = sm.wh + i * D + k * Dh;
    float s = 0.f, t = 0.f;
    for (int c = lane; c < Dh; c += 32) {
      const float x = ld1(w + c);
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

// Dropout stream layout (any bijection works as long as forward and backward agree):
//   attention mask (b, k, i, j)  : murmur-hashed element index ((b*K + k)*N + i)*N + j   (rng.cuh: dropout_scale1_hash)
//   output mask    (b, i, c)     : murmur-hashed column-PAIR index (b*N + i)*(D/2) + c/2, 16 bits per element
//                                  (rng.cuh: dropout_scale2_hash) — a thread of the tensor-core kernels owns column pairs
__device__ __forceinline__ float att_keep(const GatParams& p, const GatGraph& gr, const DropoutCfg& cfg, int b, int k, int i, int j) {
  const int heads = p.mask_heads > 0 ? p.mask_heads : p.heads;
  return dropout_scale1_hash(cfg, (((unsigned long long)b * heads + gr.mask_k0 + k) * p.N + i) * p.N + j);
}
__device__ __forceinline__ float2 out_keep2(const GatParams& p, const GatGraph& gr, const DropoutCfg& cfg, int b, int i, int pair) {
  const int D = p.mask_D > 0 ? p.mask_D : p.D;
  return dropout_scale2_hash(cfg, ((unsigned long long)b * p.N + i) * (D >> 1) + gr.mask_pair0 + pair);
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
      Prow[j] = Prow[j] * sm.gate[j] * att_keep(p, gr, datt, b, k, i, j);   // gate multiplies the values: fold it into P
  }
  __syncthreads();

  // aggregation: out[i][c] = dropout(ELU(sum_j P[k(c)][i][j] * Wh[j][c])) ; 2 columns x 8 rows per thread and pass
  const DropoutCfg dout{p.seed, gr.drop_stream + 1u, p.p_out, p.seed_off};
  act_t* outp = gr.out + (long long)b * N * p.ld_out;
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
          w[q] = (j0 + q < N) ? ld2(sm.wh + (size_t)(j0 + q) * D + 2 * pair) : make_float2(0.f, 0.f);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          if (i0 + r < N) {
            const float4 pv = *reinterpret_cast<const float4*>(Pk + (size_t)(i0 + r) * NP + j0);
            acc[r][0] += pv.x * w[0].x + pv.y * w[1].x + pv.z * w[2].x + pv.w * w[3].x;
            acc[r][1] += pv.x * w[0].y + pv.y * w[1].y + pv.z * w[2].y + pv.w * w[3].y;
          }
        }
      }
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int i = i0 + r;
        if (i < N) {
          float o0 = eluf_(acc[r][0]), o1 = eluf_(acc[r][1]);
          if (p.p_out > 0.f) {
            const float2 keep = out_keep2(p, gr, dout, b, i, pair);
            o0 *= keep.x;
            o1 *= keep.y;
          }
          st2(outp + (long long)i * p.ld_out + c, o0, o1);
          if (gr.out_f32 != nullptr)
            *reinterpret_cast<float2*>(gr.out_f32 + ((long long)b * N + i) * D + c) = make_float2(o0, o1);
        }
      }
    }
  }
}

#ifndef DVGR_F32
// ------------------------------------------------------------------------------------------------------ fast path (forward)
// D = 768, 4 heads — the only configuration DualVGR builds (reference model/models.py:95-100) — and N <= NP nodes.
// The generic kernel above spends ~95 % of its issue slots on integer / predicate work around a 300 k-MAC aggregation
// (ncu: 0.19 M warp instructions per CTA, IPC 2.8, DRAM 6 %). Here the aggregation runs on the warp-level tensor cores
// (mma.sync m16n8k16, bf16 x bf16 -> fp32): A = the gated, masked attention matrix of the head, split into a bf16 high
// and low half so that it keeps 16 mantissa bits (the fp32 copy of the output feeds the ill-conditioned auxiliary
// losses), B = the Wh tile already staged in shared memory, read with ldmatrix.trans. ~15 k warp instructions per CTA:
// the kernel becomes what SURVEY.md §8d says it should be, HBM-bound.
constexpr int kFD = 768, kFK = 4, kFDh = 192;
constexpr int kFWP = kFD + 8;             // Wh row pitch (elements): 1552 B, an odd multiple of 16 B -> conflict-free ldmatrix
constexpr int kFThreads = 256;            // 8 warps x 96 output columns (2 warps per head)

template <int NP> struct GatFast {
  static constexpr int PP = NP + 8;       // attention-matrix row pitch (elements): (NP + 8) * 2 B is an odd multiple of 16 B
  static constexpr size_t WH_BYTES = (size_t)NP * kFWP * 2;
  static constexpr size_t P_BYTES = (size_t)2 * kFK * NP * PP * 2;          // high and low halves
  static constexpr size_t ST_BYTES = (size_t)2 * kFK * NP * 4;
  static constexpr size_t GATE_BYTES = (size_t)NP * 4;
  static constexpr size_t ADJ_BYTES = (size_t)NP * NP;
  static constexpr size_t FWD_BYTES = WH_BYTES + P_BYTES + ST_BYTES + GATE_BYTES + ADJ_BYTES;
};

// stage one [N][768] bf16 tile (row stride ld) into shared memory with the padded pitch; rows >= N are zero.
// cp.async (LDGSTS): every 16-byte piece of the tile is in flight at once and never passes through registers — a plain
// load/store loop exposes one HBM latency per iteration (8 per thread), which is what bounded the first version.
// Completion: fast_stage_wait() + __syncthreads().
template <int NP>
__device__ __forceinline__ void fast_stage_tile(__nv_bfloat16* dst, const __nv_bfloat16* src, long long ld, int N) {
  for (int v = threadIdx.x; v < NP * (kFD / 8); v += kFThreads) {
    const int r = v / (kFD / 8), c = v - r * (kFD / 8);
    const uint32_t d = smem_u32(dst + (size_t)r * kFWP + c * 8);
    if (r < N) {
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src + (long long)r * ld + c * 8) : "memory");
    } else {
      asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(d), "r"(0) : "memory");
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void fast_stage_wait() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// s_i = a1 . Wh_i, t_i = a2 . Wh_i per head (one warp per (node, head))
template <int NP>
__device__ __forceinline__ void fast_logit_dots(const __nv_bfloat16* wh, const float* __restrict__ avec, int N, float* s_s,
                                                float* s_t) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int pr = warp; pr < N * kFK; pr += kFThreads / 32) {
    const int i = pr / kFK, k = pr - i * kFK;
    const float* a = avec + k * (2 * kFDh + 1);
    const __nv_bfloat16* w = wh + (size_t)i * kFWP + k * kFDh;
    float s = 0.f, t = 0.f;
#pragma unroll
    for (int q = 0; q < kFDh / 64; ++q) {
      const int c = 2 * lane + 64 * q;
      const float2 x = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(w + c));
      s += __ldg(a + c) * x.x + __ldg(a + c + 1) * x.y;
      t += __ldg(a + kFDh + c) * x.x + __ldg(a + kFDh + c + 1) * x.y;
    }
    s = warp_sum(s);
    t = warp_sum(t);
    if (lane == 0) {
      s_s[k * NP + i] = s;
      s_t[k * NP + i] = t;
    }
  }
}

// softmax over the neighbours of node i in head k (one warp); returns this lane's probabilities for j = lane (+32)
template <int NP>
__device__ __forceinline__ void fast_softmax_row(const GatParams& p, const float* __restrict__ avec, const float* s_s,
                                                 const float* s_t, const unsigned char* s_adj, int k, int i, int lane,
                                                 float (&prob)[NP / 32]) {
  const int N = p.N;
  const float cb = __ldg(avec + k * (2 * kFDh + 1) + 2 * kFDh);
  const float si = s_s[k * NP + i];
  float m = -INFINITY;
#pragma unroll
  for (int q = 0; q < NP / 32; ++q) {
    const int j = lane + 32 * q;
    prob[q] = -INFINITY;
    if (j < N) {
      float u = si + s_t[k * NP + j] + cb;
      u = u > 0.f ? u : p.slope * u;
      prob[q] = s_adj[i * NP + j] ? u : -9e15f;
      m = fmaxf(m, prob[q]);
    }
  }
  m = warp_max(m);
  float sum = 0.f;
#pragma unroll
  for (int q = 0; q < NP / 32; ++q) {
    const int j = lane + 32 * q;
    prob[q] = j < N ? __expf(prob[q] - m) : 0.f;
    sum += prob[q];
  }
  sum = warp_sum(sum);
  const float inv = 1.f / sum;
#pragma unroll
  for (int q = 0; q < NP / 32; ++q) prob[q] *= inv;
}

__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

template <int NP>
__global__ void __launch_bounds__(kFThreads) gat_attn_fwd_mma_kernel(const GatParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using F = GatFast<NP>;
  constexpr int PP = F::PP;
  __nv_bfloat16* wh = reinterpret_cast<__nv_bfloat16*>(smem_raw);
  __nv_bfloat16* Phi = reinterpret_cast<__nv_bfloat16*>(smem_raw + F::WH_BYTES);     // [K][NP][PP]
  __nv_bfloat16* Plo = Phi + (size_t)kFK * NP * PP;
  float* s_s = reinterpret_cast<float*>(smem_raw + F::WH_BYTES + F::P_BYTES);
  float* s_t = s_s + kFK * NP;
  float* s_gate = s_t + kFK * NP;
  unsigned char* s_adj = reinterpret_cast<unsigned char*>(s_gate + NP);
  const int b = blockIdx.x;
  const GatGraph& gr = p.g[blockIdx.y];
  const int N = p.N;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // 1. stage the node block, the gate and the adjacency
  fast_stage_tile<NP>(wh, gr.wh + (long long)b * N * p.ld_wh, p.ld_wh, N);
  for (int i = tid; i < NP; i += kFThreads) s_gate[i] = i < N ? gr.gate[(long long)b * N + i] : 0.f;
  for (int e = tid; e < NP * NP; e += kFThreads) {
    const int i = e / NP, j = e - i * NP;
    s_adj[e] = (i < N && j < N && p.adj[i * N + j] > 0.f) ? 1 : 0;
  }
  fast_stage_wait();
  __syncthreads();
  // 2. logit halves
  fast_logit_dots<NP>(wh, gr.avec, N, s_s, s_t);
  __syncthreads();
  // 3. attention rows: softmax, attention dropout, gate (multiplies the values: folded into P), bf16 high / low halves
  const HashMask matt(DropoutCfg{p.seed, gr.drop_stream, p.p_att, p.seed_off});
  for (int pr = warp; pr < kFK * NP; pr += kFThreads / 32) {
    const int k = pr / NP, i = pr - k * NP;
    float prob[NP / 32];
    if (i < N) {
      fast_softmax_row<NP>(p, gr.avec, s_s, s_t, s_adj, k, i, lane, prob);
    } else {
#pragma unroll
      for (int q = 0; q < NP / 32; ++q) prob[q] = 0.f;
    }
    const unsigned long long arow = (((unsigned long long)b * kFK + k) * N + i) * N;      // att_keep's element index
#pragma unroll
    for (int q = 0; q < NP / 32; ++q) {
      const int j = lane + 32 * q;
      float v = 0.f;
      if (i < N && j < N) v = prob[q] * s_gate[j] * matt.keep1(arow + j);
      __nv_bfloat16 hi, lo;
      split_bf16(v, hi, lo);
      Phi[((size_t)k * NP + i) * PP + j] = hi;
      Plo[((size_t)k * NP + i) * PP + j] = lo;
    }
  }
  __syncthreads();

  // 4. aggregation on the tensor cores: out[i][c] = dropout(ELU(sum_j P~[k(c)][i][j] Wh[j][c])); warp w owns columns
  //    [96 w, 96 w + 96) of head w / 2
  const HashMask mout(DropoutCfg{p.seed, gr.drop_stream + 1u, p.p_out, p.seed_off});
  __nv_bfloat16* outp = gr.out + (long long)b * N * p.ld_out;
  float* out32 = gr.out_f32 != nullptr ? gr.out_f32 + (long long)b * N * kFD : nullptr;
  const int k = warp >> 1, cb = 96 * warp;
  const int g = lane >> 2, t = lane & 3;
  const int lrow = (lane & 7) + ((lane >> 3) & 1) * 8;      // ldmatrix row within a 16-row block
  const int ksteps = (N + 15) >> 4;
  const uint32_t wh_s = smem_u32(wh), phi_s = smem_u32(Phi), plo_s = smem_u32(Plo);
#pragma unroll 1
  for (int mt = 0; mt < NP / 16; ++mt) {
    if (mt * 16 >= N) break;
    uint32_t ahi[NP / 16][4], alo[NP / 16][4];
#pragma unroll
    for (int ks = 0; ks < NP / 16; ++ks) {
      const uint32_t off = (uint32_t)((((size_t)k * NP + mt * 16 + lrow) * PP + ks * 16 + (lane >> 4) * 8) * 2);
      ldsm_x4(ahi[ks], phi_s + off);
      ldsm_x4(alo[ks], plo_s + off);
    }
    // per-row bases of this thread's two rows (g, g + 8): everything below only adds the compile-time column offset
    const int i0 = mt * 16 + g, i1 = i0 + 8;
    const int c0 = cb + 2 * t;
    __nv_bfloat16* ob0 = outp + (long long)i0 * p.ld_out + c0;
    __nv_bfloat16* ob1 = outp + (long long)i1 * p.ld_out + c0;
    float* of0 = out32 != nullptr ? out32 + (long long)i0 * kFD + c0 : nullptr;
    float* of1 = out32 != nullptr ? out32 + (long long)i1 * kFD + c0 : nullptr;
    const unsigned long long m0 = ((unsigned long long)b * N + i0) * (kFD / 2) + (c0 >> 1);
    const unsigned long long m1 = ((unsigned long long)b * N + i1) * (kFD / 2) + (c0 >> 1);
    const uint32_t bbase = wh_s + (uint32_t)(((size_t)lrow * kFWP + cb) * 2);
#pragma unroll
    for (int nt = 0; nt < 12; ++nt) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int ks = 0; ks < NP / 16; ++ks) {
        if (ks < ksteps) {
          uint32_t b0, b1;
          ldsm_x2_trans(b0, b1, bbase + (uint32_t)(((size_t)ks * 16 * kFWP + nt * 8) * 2));
          mma_bf16(acc, ahi[ks], b0, b1);
          mma_bf16(acc, alo[ks], b0, b1);
        }
      }
      if (i0 < N) {
        const float2 keep = mout.keep2(m0 + nt * 4);
        const float o0 = elu_fast(acc[0]) * keep.x, o1 = elu_fast(acc[1]) * keep.y;
        *reinterpret_cast<__nv_bfloat162*>(ob0 + nt * 8) = __floats2bfloat162_rn(o0, o1);
        if (of0 != nullptr) *reinterpret_cast<float2*>(of0 + nt * 8) = make_float2(o0, o1);
      }
      if (i1 < N) {
        const float2 keep = mout.keep2(m1 + nt * 4);
        const float o0 = elu_fast(acc[2]) * keep.x, o1 = elu_fast(acc[3]) * keep.y;
        *reinterpret_cast<__nv_bfloat162*>(ob1 + nt * 8) = __floats2bfloat162_rn(o0, o1);
        if (of1 != nullptr) *reinterpret_cast<float2*>(of1 + nt * 8) = make_float2(o0, o1);
      }
    }
  }
}

#endif  // DVGR_F32
// ------------------------------------------------------------------------------------------------------ backward
// extra shared memory: dz [N][D] bf16, dP [heads][N][NP] f32 (dP~ -> du -> transposed P~), ds/dt [heads][N], dg [N]
__host__ __device__ inline size_t gat_smem_bwd_extra(int N, int D, int heads) {
  const int NP = round4(N);
  return (size_t)N * D * kEsz + (size_t)heads * N * NP * 4 + (size_t)heads * N * 4 * 2 + (size_t)round4(N) * 4 + 16;
}

__global__ void __launch_bounds__(kGatThreads) gat_attn_bwd_kernel(const GatParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int b = blockIdx.x;
  const GatGraph& gr = p.g[blockIdx.y];
  const int N = p.N, D = p.D, K = p.heads, Dh = D / K, NP = round4(N);
  const GatSmem sm = carve(smem_raw, N, D, K);
  unsigned char* ext = smem_raw + gat_smem_common(N, D, K);
  act_t* dz = reinterpret_cast<act_t*>(ext);
  ext += (size_t)N * D * kEsz;
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
    const act_t* hp = gr.out + (long long)b * N * p.ld_out;
    const act_t* dop = gr.dout + (long long)b * N * p.ld_out;
    const float keepf = 1.f - p.p_out;
    const int quads = (N + 3) >> 2;
    for (int u = tid; u < quads * (D / 2); u += blockDim.x) {
      const int rq = u / (D / 2), pair = u - rq * (D / 2), c = pair * 2;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int i = rq * 4 + r;
        if (i >= N) break;
        const float2 keep = out_keep2(p, gr, dout, b, i, pair);
        float2 h = ld2(hp + (long long)i * p.ld_out + c);
        float2 d = ld2(dop + (long long)i * p.ld_out + c);
        if (gr.dout_f32 != nullptr) {
          const float2 e = *reinterpret_cast<const float2*>(gr.dout_f32 + ((long long)b * N + i) * D + c);
          d.x += e.x;
          d.y += e.y;
        }
        if (p.p_out > 0.f) {
          d.x *= keep.x;
          d.y *= keep.y;
          h.x *= keepf;   // recover ELU(z) of the kept elements (dropped ones have zero gradient anyway)
          h.y *= keepf;
        }
        d.x *= elu_grad_from_out(h.x);
        d.y *= elu_grad_from_out(h.y);
        st2(dz + (size_t)i * D + c, d.x, d.y);
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
      zreg[q] = c < Dh ? ld1(dz + (size_t)i * D + k * Dh + c) : 0.f;
    }
    float mine[2] = {0.f, 0.f};
    for (int j = 0; j < N; ++j) {
      float acc = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int c = lane + 32 * q;
        if (c < Dh) acc += zreg[q] * ld1(sm.wh + (size_t)j * D + k * Dh + c);
      }
      acc = warp_sum(acc);
      if ((j & 31) == lane) mine[j >> 5] = acc;
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int j = lane + 32 * q;
      if (j < N) dP[((size_t)k * N + i) * NP + j] = mine[q] * sm.gate[j] * att_keep(p, gr, datt, b, k, i, j);
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
    dP[e] = (i < N) ? sm.P[((size_t)k * N + i) * NP + j] * att_keep(p, gr, datt, b, k, i, j) : 0.f;
  }
  __syncthreads();

  // 4. dV_j[c] = sum_i P~_ij dz_i[c] ; dWh_j = g_j dV_j + ds_j a1 + dt_j a2 ; dgate_j ; da1, da2   (8 j x 2 columns / pass)
  act_t* dwhp = gr.dwh + (long long)b * N * p.ld_wh;
  float* dav = gr.davec + (long long)b * (gr.ld_davec > 0 ? gr.ld_davec : (long long)K * (2 * Dh + 1));
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
            z[q] = (i0 + q < N) ? ld2(dz + (size_t)(i0 + q) * D + 2 * pair) : make_float2(0.f, 0.f);
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
        const float2 w = active ? ld2(sm.wh + (size_t)j * D + 2 * pair) : make_float2(0.f, 0.f);
        float part = acc[r][0] * w.x + acc[r][1] * w.y;     // gate gradient: dg_j += sum_c dV_j[c] Wh_j[c]
        part = warp_sum(part);
        if (lane == 0) atomicAdd(&dg[j], part);
        if (active) {
          const float gj = sm.gate[j], dsj = ds[k * N + j], dtj = dt[k * N + j];
          const float ox = gj * acc[r][0] + dsj * a1[cl] + dtj * a2[cl];
          const float oy = gj * acc[r][1] + dsj * a1[cl + 1] + dtj * a2[cl + 1];
          st2(dwhp + (long long)j * p.ld_wh + c, ox, oy);
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
  for (int i = tid; i < N; i += blockDim.x) {
    if (gr.atomic_dgate) atomicAdd(gr.dgate + (long long)b * N + i, dg[i]);
    else gr.dgate[(long long)b * N + i] = dg[i];
  }
  if (tid < K) dav[tid * (2 * Dh + 1) + 2 * Dh] = dc_s[tid];
}

#ifndef DVGR_F32
// ------------------------------------------------------------------------------------------------------ fast path (backward)
// Same configuration as the forward fast path, N <= 32. The heads of a graph are independent except for the gate
// gradient, so ONE CTA handles (video, graph, PAIR of heads): 384 of the 768 columns. That halves the two staged tiles
// (Wh and dz) to 25 KB each — 3 CTAs per SM instead of 1 (the whole-graph version measured 12 % occupancy, IPC 1.3,
// "no eligible warp" 66 % of the cycles) — and doubles the grid to 2 B x graphs CTAs. The three N x N x D products run on
// mma.sync:
//   dP~[k][i][j] = sum_{c in head k} dz[i][c] Wh[j][c]        A = dz (row-major), B = Wh (stored [n][k])
//   dV[j][c]     = sum_i P~^T[k][j][i] dz[i][c]               A = P~^T (bf16 high + low halves), B = dz (ldmatrix.trans)
// and everything that hangs off them (gate / attention-vector gradients, dWh) is folded into the fragment epilogues.
// dgate receives the two head pairs' partial sums by atomicAdd: the caller zeroes it (two addends: deterministic).
// NP = 32 nodes: kBH = 2 heads per CTA (97 KB, 2 CTAs / SM); NP = 64 nodes (BASELINE config 5): kBH = 1 head per CTA (119 KB, 1 / SM)
template <int NP, int kBH> struct GatFastBwd {
  static constexpr int BC = kBH * kFDh;            // columns per CTA
  static constexpr int BWP = BC + 8;               // tile row pitch (elements): an odd multiple of 16 B
  static constexpr int PP = NP + 8, NPF = NP + 4, MT = NP / 16;
  static constexpr size_t TILE_BYTES = (size_t)NP * BWP * 2;
  static constexpr size_t PT_BYTES = (size_t)2 * kBH * NP * PP * 2;
  static constexpr size_t DP_BYTES = (size_t)kBH * NP * NPF * 4;
  static constexpr size_t SMALL_BYTES = (size_t)(4 * kBH * NP + 2 * NP + 8) * 4;      // s, t, ds, dt, gate, dg, dc
  static constexpr size_t ADJ_BYTES = (size_t)NP * NP;
  // + a third tile: the LOW half of dz. dV_j = sum_i P_ij dz_i averages dz over the (near-uniformly attended) nodes, and the
  // gradients of the auxiliary losses are centred over the nodes: the true sum is a small residual of large terms, so dz
  // rounded to ONE bf16 turns it into noise (measured: 6x error on the full-loss gradient of config 2); dz = hi + lo does not.
  static constexpr size_t LO_OFFSET = 2 * TILE_BYTES + PT_BYTES + DP_BYTES + SMALL_BYTES + ((ADJ_BYTES + 15) & ~(size_t)15);
  static constexpr size_t BYTES = LO_OFFSET + TILE_BYTES;
};

template <int NP, int kBH>
__global__ void __launch_bounds__(kFThreads, (NP == 32) ? 2 : 1) gat_attn_bwd_mma_kernel(const GatParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using FB = GatFastBwd<NP, kBH>;
  constexpr int PP = FB::PP, NPF = FB::NPF, MT = FB::MT, kBC = FB::BC, kBWP = FB::BWP;
  constexpr int kGroups = kFK / kBH;               // CTAs per (video, graph)
  constexpr int WPH = (kFThreads / 32) / kBH;      // warps per head
  __nv_bfloat16* wh = reinterpret_cast<__nv_bfloat16*>(smem_raw);                                // [NP][kBWP]
  __nv_bfloat16* dz = reinterpret_cast<__nv_bfloat16*>(smem_raw + FB::TILE_BYTES);
  __nv_bfloat16* dzl = reinterpret_cast<__nv_bfloat16*>(smem_raw + FB::LO_OFFSET);                // low half of dz
  __nv_bfloat16* PThi = reinterpret_cast<__nv_bfloat16*>(smem_raw + 2 * FB::TILE_BYTES);         // [kBH][NP (j)][PP (i)]
  __nv_bfloat16* PTlo = PThi + (size_t)kBH * NP * PP;
  float* dP = reinterpret_cast<float*>(smem_raw + 2 * FB::TILE_BYTES + FB::PT_BYTES);            // [kBH][NP (i)][NPF (j)]
  float* s_s = reinterpret_cast<float*>(smem_raw + 2 * FB::TILE_BYTES + FB::PT_BYTES + FB::DP_BYTES);
  float* s_t = s_s + kBH * NP;
  float* ds = s_t + kBH * NP;
  float* dt = ds + kBH * NP;
  float* s_gate = dt + kBH * NP;
  float* dg = s_gate + NP;
  float* dc_s = dg + NP;
  unsigned char* s_adj = reinterpret_cast<unsigned char*>(dc_s + 8);
  const int b = blockIdx.x;
  const GatGraph& gr = p.g[blockIdx.y / kGroups];
  const int hp = blockIdx.y % kGroups, head0 = hp * kBH, col0 = hp * kBC;
  const int N = p.N;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int lrow = (lane & 7) + ((lane >> 3) & 1) * 8;
  const HashMask matt(DropoutCfg{p.seed, gr.drop_stream, p.p_att, p.seed_off});
  const HashMask mout(DropoutCfg{p.seed, gr.drop_stream + 1u, p.p_out, p.seed_off});

  // 1. stage this head pair's Wh columns (cp.async); dz = (dout [+ dout_f32]) * mask * ELU'(z) straight from global memory
  {
    const __nv_bfloat16* src = gr.wh + (long long)b * N * p.ld_wh + col0;
    for (int v = tid; v < NP * (kBC / 8); v += kFThreads) {
      const int r = v / (kBC / 8), c = v - r * (kBC / 8);
      const uint32_t d = smem_u32(wh + (size_t)r * kBWP + c * 8);
      if (r < N) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src + (long long)r * p.ld_wh + c * 8) : "memory");
      } else {
        asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(d), "r"(0) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  for (int i = tid; i < NP; i += kFThreads) {
    s_gate[i] = i < N ? gr.gate[(long long)b * N + i] : 0.f;
    dg[i] = 0.f;
  }
  for (int e = tid; e < NP * NP; e += kFThreads) {
    const int i = e / NP, j = e - i * NP;
    s_adj[e] = (i < N && j < N && p.adj[i * N + j] > 0.f) ? 1 : 0;
  }
  {
    const __nv_bfloat16* hp_ = gr.out + (long long)b * N * p.ld_out + col0;
    const __nv_bfloat16* dop = gr.dout + (long long)b * N * p.ld_out + col0;
    const float* d32 = gr.dout_f32 != nullptr ? gr.dout_f32 + (long long)b * N * kFD + col0 : nullptr;
    const float keepf = 1.f - p.p_out;
    const bool vec = ((p.ld_out & 7) == 0);
    constexpr int kBatch = 4;                       // items per thread whose loads are all issued before any is consumed
    const int items = N * (kBC / 8);                // 8 columns per item
    for (int v0 = tid; v0 < items; v0 += kFThreads * kBatch) {
      uint4 hv[kBatch], dv[kBatch];
      float4 e0[kBatch], e1[kBatch];
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const int v = v0 + u * kFThreads;
        hv[u] = dv[u] = make_uint4(0, 0, 0, 0);
        e0[u] = e1[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (v < items) {
          const int i = v / (kBC / 8), c = (v - i * (kBC / 8)) * 8;
          if (vec) {
            hv[u] = *reinterpret_cast<const uint4*>(hp_ + (long long)i * p.ld_out + c);
            dv[u] = *reinterpret_cast<const uint4*>(dop + (long long)i * p.ld_out + c);
          } else {
            uint32_t* hw_ = reinterpret_cast<uint32_t*>(&hv[u]);
            uint32_t* dw_ = reinterpret_cast<uint32_t*>(&dv[u]);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              hw_[e] = *reinterpret_cast<const uint32_t*>(hp_ + (long long)i * p.ld_out + c + 2 * e);
              dw_[e] = *reinterpret_cast<const uint32_t*>(dop + (long long)i * p.ld_out + c + 2 * e);
            }
          }
          if (d32 != nullptr) {
            e0[u] = *reinterpret_cast<const float4*>(d32 + (long long)i * kFD + c);
            e1[u] = *reinterpret_cast<const float4*>(d32 + (long long)i * kFD + c + 4);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const int v = v0 + u * kFThreads;
        if (v < items) {
          const int i = v / (kBC / 8), c = (v - i * (kBC / 8)) * 8;
          const float ex[8] = {e0[u].x, e0[u].y, e0[u].z, e0[u].w, e1[u].x, e1[u].y, e1[u].z, e1[u].w};
          const uint32_t* hw = reinterpret_cast<const uint32_t*>(&hv[u]);
          const uint32_t* dw = reinterpret_cast<const uint32_t*>(&dv[u]);
          const unsigned long long mrow = ((unsigned long long)b * N + i) * (kFD / 2) + ((col0 + c) >> 1);
          uint4 o, ol;
          uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
          uint32_t* olw = reinterpret_cast<uint32_t*>(&ol);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float2 h = unpack_bf16x2(hw[e]);
            float2 d = unpack_bf16x2(dw[e]);
            d.x += ex[2 * e];
            d.y += ex[2 * e + 1];
            if (mout.on) {
              const float2 keep = mout.keep2(mrow + e);
              d.x *= keep.x;
              d.y *= keep.y;
              h.x *= keepf;     // recover ELU(z) of the kept elements (dropped ones have zero gradient anyway)
              h.y *= keepf;
            }
            const float zx = d.x * elu_grad_from_out(h.x), zy = d.y * elu_grad_from_out(h.y);
            ow[e] = pack_bf16x2(zx, zy);
            const float2 hi = unpack_bf16x2(ow[e]);
            olw[e] = pack_bf16x2(zx - hi.x, zy - hi.y);
          }
          *reinterpret_cast<uint4*>(dz + (size_t)i * kBWP + c) = o;
          *reinterpret_cast<uint4*>(dzl + (size_t)i * kBWP + c) = ol;
        }
      }
    }
    for (int v = items + tid; v < NP * (kBC / 8); v += kFThreads) {      // zero rows N .. NP-1
      const int i = v / (kBC / 8), c = (v - i * (kBC / 8)) * 8;
      *reinterpret_cast<uint4*>(dz + (size_t)i * kBWP + c) = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(dzl + (size_t)i * kBWP + c) = make_uint4(0, 0, 0, 0);
    }
  }
  fast_stage_wait();
  __syncthreads();
  // logit halves s_i = a1 . Wh_i, t_i = a2 . Wh_i of the two heads (one warp per (node, head))
  for (int pr = warp; pr < N * kBH; pr += kFThreads / 32) {
    const int i = pr / kBH, kl = pr - i * kBH;
    const float* a = gr.avec + (head0 + kl) * (2 * kFDh + 1);
    const __nv_bfloat16* w = wh + (size_t)i * kBWP + kl * kFDh;
    float s = 0.f, tt = 0.f;
#pragma unroll
    for (int q = 0; q < kFDh / 64; ++q) {
      const int c = 2 * lane + 64 * q;
      const float2 x = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(w + c));
      s += __ldg(a + c) * x.x + __ldg(a + c + 1) * x.y;
      tt += __ldg(a + kFDh + c) * x.x + __ldg(a + kFDh + c + 1) * x.y;
    }
    s = warp_sum(s);
    tt = warp_sum(tt);
    if (lane == 0) {
      s_s[kl * NP + i] = s;
      s_t[kl * NP + i] = tt;
    }
  }

  // 2. dP~ = dz Wh^T per head on the tensor cores; warp w: head w / WPH, neighbour columns j in [8 (w % WPH), +8)
  {
    static_assert(NP / 8 == WPH, "one 8-wide neighbour tile per warp");
    const int kl = warp / WPH, nt = warp % WPH;
    const uint32_t dz_s = smem_u32(dz), wh_s = smem_u32(wh);
    float acc[MT][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) acc[mt][0] = acc[mt][1] = acc[mt][2] = acc[mt][3] = 0.f;
    const uint32_t bbase = wh_s + (uint32_t)((((size_t)nt * 8 + (lane & 7)) * kBWP + kl * kFDh + ((lane >> 3) & 1) * 8) * 2);
    const uint32_t abase = dz_s + (uint32_t)((((size_t)lrow) * kBWP + kl * kFDh + (lane >> 4) * 8) * 2);
#pragma unroll
    for (int ks = 0; ks < kFDh / 16; ++ks) {
      uint32_t b0, b1;
      ldsm_x2(b0, b1, bbase + ks * 32);
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        uint32_t a[4];
        ldsm_x4(a, abase + (uint32_t)((size_t)mt * 16 * kBWP * 2) + ks * 32);
        mma_bf16(acc[mt], a, b0, b1);
      }
    }
    const int k = head0 + kl;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int i = mt * 16 + g + 8 * h, j = nt * 8 + 2 * t;
        const unsigned long long arow = (((unsigned long long)b * kFK + k) * N + i) * N;
        float v0 = 0.f, v1 = 0.f;
        if (i < N && j < N) v0 = acc[mt][2 * h] * s_gate[j] * matt.keep1(arow + j);
        if (i < N && j + 1 < N) v1 = acc[mt][2 * h + 1] * s_gate[j + 1] * matt.keep1(arow + j + 1);
        *reinterpret_cast<float2*>(dP + ((size_t)kl * NP + i) * NPF + j) = make_float2(v0, v1);
      }
  }
  __syncthreads();

  // 3. softmax / LeakyReLU backward per row: du (in place of dP~), ds_i; and the transposed masked attention matrix
  //    P~^T[k][j][i] = P_ij * mask_ij (ungated) as bf16 high / low halves — the A operand of the dV product
  for (int pr = warp; pr < kBH * NP; pr += kFThreads / 32) {
    const int kl = pr / NP, i = pr - kl * NP, k = head0 + kl;
    const float* av = gr.avec + k * (2 * kFDh + 1);
    constexpr int Q = NP / 32;                       // neighbours per lane
    float prob[Q], uu[Q];
    float srow = 0.f;
#pragma unroll
    for (int q = 0; q < Q; ++q) prob[q] = uu[q] = 0.f;
    if (i < N) {
      const float cb = __ldg(av + 2 * kFDh);
      const float si = s_s[kl * NP + i];
      float m = -INFINITY;
#pragma unroll
      for (int q = 0; q < Q; ++q) {
        const int j = lane + 32 * q;
        prob[q] = -INFINITY;
        if (j < N) {
          uu[q] = si + s_t[kl * NP + j] + cb;
          const float lu = uu[q] > 0.f ? uu[q] : p.slope * uu[q];
          prob[q] = s_adj[i * NP + j] ? lu : -9e15f;
          m = fmaxf(m, prob[q]);
        }
      }
      m = warp_max(m);
      float sum = 0.f;
#pragma unroll
      for (int q = 0; q < Q; ++q) {
        prob[q] = (lane + 32 * q) < N ? __expf(prob[q] - m) : 0.f;
        sum += prob[q];
      }
      sum = warp_sum(sum);
      const float inv = 1.f / sum;
      float* dProw = dP + ((size_t)kl * NP + i) * NPF;
      float dp[Q], dot = 0.f;
#pragma unroll
      for (int q = 0; q < Q; ++q) {
        prob[q] *= inv;
        dp[q] = (lane + 32 * q) < N ? dProw[lane + 32 * q] : 0.f;
        dot += prob[q] * dp[q];
      }
      dot = warp_sum(dot);
#pragma unroll
      for (int q = 0; q < Q; ++q) {
        const int j = lane + 32 * q;
        if (j < N) {
          float de = prob[q] * (dp[q] - dot);
          if (!s_adj[i * NP + j]) de = 0.f;
          const float du = de * (uu[q] > 0.f ? 1.f : p.slope);
          dProw[j] = du;
          srow += du;
        }
      }
      srow = warp_sum(srow);
    }
    if (lane == 0) ds[kl * NP + i] = srow;
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      const int j = lane + 32 * q;
      float v = 0.f;
      if (i < N && j < N) v = prob[q] * matt.keep1((((unsigned long long)b * kFK + k) * N + i) * N + j);
      __nv_bfloat16 hi, lo;
      split_bf16(v, hi, lo);
      PThi[((size_t)kl * NP + j) * PP + i] = hi;
      PTlo[((size_t)kl * NP + j) * PP + i] = lo;
    }
  }
  __syncthreads();
  for (int pr = tid; pr < kBH * NP; pr += kFThreads) {        // dt_j = sum_i du_ij
    const int kl = pr / NP, j = pr - kl * NP;
    float a = 0.f;
    if (j < N)
      for (int i = 0; i < N; ++i) a += dP[((size_t)kl * NP + i) * NPF + j];
    dt[pr] = a;
  }
  if (tid < kBH) {                                            // dc_k = sum_i ds_i
    float a = 0.f;
    for (int i = 0; i < N; ++i) a += ds[tid * NP + i];
    dc_s[tid] = a;
  }
  __syncthreads();

  // 4. dV = P~^T dz on the tensor cores; dWh_j = g_j dV_j + ds_j a1 + dt_j a2 ; dgate_j ; da1, da2 in the fragment epilogue.
  //    warp w owns columns [CW w, CW w + CW) of the CTA's heads (CW = 48 or 24) = head w / WPH
  {
    constexpr int CW = kBC / (kFThreads / 32), NT4 = CW / 8;
    constexpr bool kResident = (MT == 2);            // A fragments of both row blocks fit in registers only for NP = 32
    const int kl = warp / WPH, k = head0 + kl, cbl = CW * warp;      // column base inside the CTA's columns
    const uint32_t dz_s = smem_u32(dz), dzl_s = smem_u32(dzl), phi_s = smem_u32(PThi), plo_s = smem_u32(PTlo);
    const int ksteps = (N + 15) >> 4;
    const float* a1 = gr.avec + k * (2 * kFDh + 1);
    const float* a2 = a1 + kFDh;
    __nv_bfloat16* dwhp = gr.dwh + (long long)b * N * p.ld_wh + col0;
    float* dav = gr.davec + ((long long)b * kFK + k) * (2 * kFDh + 1);
    // A fragments of this head: resident across the column tiles for NP = 32 (32 registers), re-loaded per tile otherwise
    uint32_t ahi[kResident ? MT : 1][MT][4], alo[kResident ? MT : 1][MT][4];
    if (kResident) {
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int ks = 0; ks < MT; ++ks) {
          const uint32_t off = (uint32_t)((((size_t)kl * NP + mt * 16 + lrow) * PP + ks * 16 + (lane >> 4) * 8) * 2);
          ldsm_x4(ahi[kResident ? mt : 0][ks], phi_s + off);
          ldsm_x4(alo[kResident ? mt : 0][ks], plo_s + off);
        }
    }
    float dgp[MT][2];
    float gj[MT][2], dsj[MT][2], dtj[MT][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int j = mt * 16 + g + 8 * h;
        dgp[mt][h] = 0.f;
        gj[mt][h] = s_gate[j]; dsj[mt][h] = ds[kl * NP + j]; dtj[mt][h] = dt[kl * NP + j];
      }
#pragma unroll 1
    for (int nt = 0; nt < NT4; ++nt) {
      const int c = cbl + nt * 8 + 2 * t, cl = c - kl * kFDh;      // column inside the CTA's columns / inside the head
      uint32_t bfr[MT][2], bfl[MT][2];
#pragma unroll
      for (int ks = 0; ks < MT; ++ks) {
        ldsm_x2_trans(bfr[ks][0], bfr[ks][1], dz_s + (uint32_t)((((size_t)ks * 16 + lrow) * kBWP + cbl + nt * 8) * 2));
        ldsm_x2_trans(bfl[ks][0], bfl[ks][1], dzl_s + (uint32_t)((((size_t)ks * 16 + lrow) * kBWP + cbl + nt * 8) * 2));
      }
      const float a1x = __ldg(a1 + cl), a1y = __ldg(a1 + cl + 1), a2x = __ldg(a2 + cl), a2y = __ldg(a2 + cl + 1);
      float da1x = 0.f, da1y = 0.f, da2x = 0.f, da2y = 0.f;
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        if (mt * 16 < N) {
          float acc[4] = {0.f, 0.f, 0.f, 0.f};
          if (!kResident) {
#pragma unroll
            for (int ks = 0; ks < MT; ++ks) {
              const uint32_t off = (uint32_t)((((size_t)kl * NP + mt * 16 + lrow) * PP + ks * 16 + (lane >> 4) * 8) * 2);
              ldsm_x4(ahi[0][ks], phi_s + off);
              ldsm_x4(alo[0][ks], plo_s + off);
            }
          }
#pragma unroll
          for (int ks = 0; ks < MT; ++ks) {
            if (ks < ksteps) {
              mma_bf16(acc, ahi[kResident ? mt : 0][ks], bfl[ks][0], bfl[ks][1]);
              mma_bf16(acc, alo[kResident ? mt : 0][ks], bfr[ks][0], bfr[ks][1]);
              mma_bf16(acc, ahi[kResident ? mt : 0][ks], bfr[ks][0], bfr[ks][1]);
            }
          }
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int j = mt * 16 + g + 8 * h;
            if (j < N) {
              const float2 w = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(wh + (size_t)j * kBWP + c));
              const float ox = gj[mt][h] * acc[2 * h] + dsj[mt][h] * a1x + dtj[mt][h] * a2x;
              const float oy = gj[mt][h] * acc[2 * h + 1] + dsj[mt][h] * a1y + dtj[mt][h] * a2y;
              *reinterpret_cast<__nv_bfloat162*>(dwhp + (long long)j * p.ld_wh + c) = __floats2bfloat162_rn(ox, oy);
              dgp[mt][h] += acc[2 * h] * w.x + acc[2 * h + 1] * w.y;      // dgate_j += sum_c dV_j[c] Wh_j[c]
              da1x += dsj[mt][h] * w.x; da1y += dsj[mt][h] * w.y;
              da2x += dtj[mt][h] * w.x; da2y += dtj[mt][h] * w.y;
            }
          }
        }
      }
      // column sums over the rows held by the 8 lanes that share t
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) {
        da1x += __shfl_xor_sync(0xffffffffu, da1x, o);
        da1y += __shfl_xor_sync(0xffffffffu, da1y, o);
        da2x += __shfl_xor_sync(0xffffffffu, da2x, o);
        da2y += __shfl_xor_sync(0xffffffffu, da2y, o);
      }
      if (g == 0) {
        dav[cl] = da1x; dav[cl + 1] = da1y;
        dav[kFDh + cl] = da2x; dav[kFDh + cl + 1] = da2y;
      }
    }
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float v = dgp[mt][h];
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        const int j = mt * 16 + g + 8 * h;
        if (t == 0 && j < N) atomicAdd(&dg[j], v);
      }
  }
  __syncthreads();
  for (int i = tid; i < N; i += kFThreads) atomicAdd(gr.dgate + (long long)b * N + i, dg[i]);      // + the other head group(s)
  if (tid < kBH) gr.davec[((long long)b * kFK + head0 + tid) * (2 * kFDh + 1) + 2 * kFDh] = dc_s[tid];
}

#endif  // DVGR_F32
}  // namespace DVGR_VNS
}  // namespace dvgr

using namespace dvgr;
using namespace dvgr::DVGR_VNS;

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
    d.wh = reinterpret_cast<const act_t*>(s.wh);
    d.gate = s.gate;
    d.avec = s.avec;
    d.out = reinterpret_cast<act_t*>(s.out);
    d.drop_stream = s.drop_stream;
    d.out_f32 = s.out_f32;
    if (bwd) {
      d.dout_f32 = s.dout_f32;
      if (!s.dout || !s.dwh || !s.dgate || !s.davec) return set_error("gat bwd: graph %d has a null gradient buffer", i);
      d.dout = reinterpret_cast<const act_t*>(s.dout);
      d.dwh = reinterpret_cast<act_t*>(s.dwh);
      d.dgate = s.dgate;
      d.davec = s.davec;
    }
  }
  return 0;
}

// Whole-graph tiles too large for shared memory (fp32 activations: N > 28 at D = 768): run every head of a graph as its own
// single-head "virtual graph" of width Dh (the heads only share the gate gradient, accumulated with atomics). One launch per
// real graph, grid (B, heads). Not available when the fp32 side outputs / gradients are requested (bf16 build only).
static int launch_head_split(const GatParams& p, int n_graphs, bool bwd, cudaStream_t st) {
  const int K = p.heads, Dh = p.D / K;
  const size_t smem = gat_smem_common(p.N, Dh, 1) + (bwd ? gat_smem_bwd_extra(p.N, Dh, 1) : 0);
  if (smem > 227 * 1024) return set_error("gat %s: %zu bytes of shared memory needed even per head", bwd ? "bwd" : "fwd", smem);
  auto kern = bwd ? gat_attn_bwd_kernel : gat_attn_fwd_kernel;
  if (smem > 48 * 1024) {      // (the limit itself: never lowers what a whole-graph launch of the same kernel configured)
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, kern);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - (int)fa.sharedSizeBytes);
    if (e != cudaSuccess) return set_error("gat (head split): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  }
  for (int g = 0; g < n_graphs; ++g) {
    const GatGraph& src = p.g[g];
    if (src.out_f32 != nullptr || src.dout_f32 != nullptr) return set_error("gat: head-split launch does not take fp32 side buffers");
    GatParams q = p;
    q.heads = 1; q.D = Dh; q.mask_heads = K; q.mask_D = p.D;
    for (int k = 0; k < K; ++k) {
      GatGraph& d = q.g[k];
      d = src;
      d.wh = src.wh + k * Dh;
      d.out = src.out + k * Dh;
      d.avec = src.avec + k * (2 * Dh + 1);
      d.mask_k0 = k; d.mask_pair0 = k * (Dh / 2);
      if (bwd) {
        d.dout = src.dout + k * Dh;
        d.dwh = src.dwh + k * Dh;
        d.davec = src.davec + k * (2 * Dh + 1);
        d.ld_davec = (long long)K * (2 * Dh + 1);
        d.atomic_dgate = 1;
      }
    }
    if (bwd && cudaMemsetAsync(src.dgate, 0, sizeof(float) * (size_t)p.B * p.N, st) != cudaSuccess)
      return set_error("gat bwd (head split): cudaMemsetAsync failed");
    kern<<<dim3(p.B, K), kGatThreads, smem, st>>>(q);
    DVGR_CHECK_LAUNCH(bwd ? "gat_attn_bwd (head split)" : "gat_attn_fwd (head split)");
  }
  return 0;
}

#ifndef DVGR_F32
static int gat_fast_knob() {      // read per call (tests flip it to compare the two paths): bit 0 forward, bit 1 backward
  const char* e = getenv("DVGR_GAT_FAST");
  return e ? atoi(e) : 3;
}
static bool gat_fast_ok(const GatParams& p) {
  return p.D == kFD && p.heads == kFK && p.N <= 64 && (p.ld_wh % 8) == 0 && (p.ld_out % 2) == 0;
}

template <int NP>
static int launch_gat_fwd_fast(const GatParams& p, int n_graphs, cudaStream_t st) {
  static bool configured = false;
  const size_t smem = GatFast<NP>::FWD_BYTES;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gat_attn_fwd_mma_kernel<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error("gat fwd (mma): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  gat_attn_fwd_mma_kernel<NP><<<dim3(p.B, n_graphs), kFThreads, smem, st>>>(p);
  DVGR_CHECK_LAUNCH("gat_attn_fwd_mma");
  return 0;
}

#endif  // DVGR_F32

extern "C" int DVGR_FN(dvgr_gat_attn_fwd)(const dvgr_gat_args* a, void* stream) {
  GatParams p;
  if (int rc = fill_gat(p, a, false)) return rc;
  if (a->B <= 0 || a->N <= 0) return 0;
#ifndef DVGR_F32
  if ((gat_fast_knob() & 1) && gat_fast_ok(p)) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    return p.N <= 32 ? launch_gat_fwd_fast<32>(p, a->n_graphs, st) : launch_gat_fwd_fast<64>(p, a->n_graphs, st);
  }
#endif
  const size_t smem = gat_smem_common(p.N, p.D, p.heads);
  if (smem > 227 * 1024 && p.heads > 1) return launch_head_split(p, a->n_graphs, false, reinterpret_cast<cudaStream_t>(stream));
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

#ifndef DVGR_F32
template <int NP, int kBH>
static int launch_gat_bwd_fast(const GatParams& p, int n_graphs, cudaStream_t st) {
  static bool configured = false;
  const size_t smem = GatFastBwd<NP, kBH>::BYTES;
  auto kern = gat_attn_bwd_mma_kernel<NP, kBH>;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error("gat bwd (mma): cudaFuncSetAttribute(%zu): %s", smem, cudaGetErrorString(e));
    configured = true;
  }
  for (int i = 0; i < n_graphs; ++i) {       // the head-group CTAs of a (video, graph) add their partial gate gradients
    cudaError_t e = cudaMemsetAsync(p.g[i].dgate, 0, sizeof(float) * (size_t)p.B * p.N, st);
    if (e != cudaSuccess) return set_error("gat bwd (mma): cudaMemsetAsync: %s", cudaGetErrorString(e));
  }
  kern<<<dim3(p.B, (kFK / kBH) * n_graphs), kFThreads, smem, st>>>(p);
  DVGR_CHECK_LAUNCH("gat_attn_bwd_mma");
  return 0;
}

#endif  // DVGR_F32

extern "C" int DVGR_FN(dvgr_gat_attn_bwd)(const dvgr_gat_args* a, void* stream) {
  GatParams p;
  if (int rc = fill_gat(p, a, true)) return rc;
  if (a->B <= 0 || a->N <= 0) return 0;
#ifndef DVGR_F32
  if ((gat_fast_knob() & 2) && gat_fast_ok(p)) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    return p.N <= 32 ? launch_gat_bwd_fast<32, 2>(p, a->n_graphs, st) : launch_gat_bwd_fast<64, 1>(p, a->n_graphs, st);
  }
#endif
  const size_t smem = gat_smem_common(p.N, p.D, p.heads) + gat_smem_bwd_extra(p.N, p.D, p.heads);
  if (smem > 227 * 1024 && p.heads > 1) return launch_head_split(p, a->n_graphs, true, reinterpret_cast<cudaStream_t>(stream));
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
