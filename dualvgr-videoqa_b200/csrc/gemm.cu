// tcgen05 / TMEM / TMA GEMM family for the DualVGR hot path (sm_100a only).
//
//   D[b][m][n] = epilogue( sum_k A[b][m][k] * B[b][n][k] )          bf16 operands, fp32 accumulation in TMEM
//
// One persistent CTA per SM, warp-specialised:
//   warp 0   : TMA producer  (cp.async.bulk.tensor, SWIZZLE_128B, 4-D tensor maps)
//   warp 1   : MMA issuer    (one thread, tcgen05.mma.cta_group::1.kind::f16, 128 x BN x 16)
//   warp 2   : TMEM allocator
//   warps 4-7: epilogue      (tcgen05.ld 32x32b -> registers -> fused epilogue -> global)
// Two TMEM accumulator stages so the epilogue of tile i overlaps the main loop of tile i+1.
//
// Operands may be K-major (rows x K, K contiguous: activations / nn.Linear weights in the forward GEMM) or
// MN-major (K x rows, rows contiguous: the same tensors seen by dgrad / wgrad), so no transposed copies exist.
//
// Epilogues (gemm.cuh: EpiMode):
//   EPI_LINEAR    bias + {none, ELU, tanh} (+ accumulate), bf16 / fp32 output, optional output-row permutation
//                 -> replaces every nn.Linear on the path (reference model/models.py:46, GraphNN.py:96,
//                    Attention.py:14-18, fusions.py:420-449, AnswerDecoder.py:173-200, model/utils.py:68)
//   EPI_LSTM_FWD  LSTM cell on the recurrent product (reference nn.LSTM, model/Preprocessing.py:227)
//   EPI_LSTM_BWD  LSTM cell backward on dh = dgates * W_hh
#include "gemm.cuh"
#include "ptx.cuh"

#include <limits.h>
#include <stdlib.h>
#include <map>
#include <mutex>
#include <string.h>
#include <tuple>
#include <vector>

#include "capi_internal.h"

namespace dvgr {

// ---- dynamic tile schedule (shared by the LSTM sequence kernels and, optionally, the GEMM): the TMA-producer thread of a CTA
// claims tile indices from a global counter and publishes them to its MMA / epilogue warps through a small shared-memory ring
// guarded by mbarriers. See the discussion at LstmSeqParams.
constexpr int kTileRing = 8;        // LSTM kernels: STAGES (4) + 2 TMEM stages + 2
constexpr int kGemmRing = 16;       // GEMM kernel: up to 6 stages
__device__ __forceinline__ int ring_tile(const int* ring, int slot) {
  int v;
  asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(ring + slot)) : "memory");
  return v;
}
template <int R = kTileRing>
__device__ __forceinline__ void ring_publish(int* ring, uint64_t* bars, int it, int tile) {
  asm volatile("st.shared.s32 [%0], %1;" ::"r"(smem_u32(ring + (it % R))), "r"(tile) : "memory");
  mbar_arrive(&bars[it % R]);       // release: the consumers' try_wait (acquire) sees the slot
}
template <int R = kTileRing>
__device__ __forceinline__ int ring_consume(const int* ring, uint64_t* bars, int it) {
  mbar_wait(&bars[it % R], static_cast<uint32_t>((it / R) & 1));
  return ring_tile(ring, it % R);
}
__device__ __forceinline__ int claim_tile(int* counter, int num_tiles) {
  const int t = atomicAdd(counter, 1);
  return t < num_tiles ? t : -1;
}

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kGemmThreads = 384;   // warps 0-3: TMA / MMA / TMEM alloc / spare ; warps 4-11: epilogue (2 per TMEM lane quarter)
constexpr int kEpiWarps = 8;

// kOcc = CTAs per SM the variant is built for: 1 (deep pipeline, GEMM-bound shapes) or 2 (3 stages, 96 KB: the LSTM
// recurrence steps, whose epilogue (scattered state loads/stores + cell math) is the bottleneck, get twice the
// epilogue warps and memory-level parallelism per SM; 2 x 256 TMEM columns still fit the 512 available).
template <int BN, int kOcc = 1> struct TileCfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (kOcc >= 2) ? 3 : (BN == 256) ? 4 : 6;   // kOcc == 3: one CTA/SM, 3 stages (more L1 left)
  static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static constexpr int STAGING_BYTES = 8 * 32 * 33 * 4;      // per-epilogue-warp transpose tiles (epi_linear)
  static constexpr int BAR_BYTES = 512;                       // pipeline barriers + TMEM slot (256 B) | tile ring (256 B)
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + BAR_BYTES + STAGING_BYTES;
  static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB of shared memory a CTA can opt into");
};

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3)
      : "memory");
}

// multicast variant: the box lands at the same shared-memory offset of every CTA in `mask` and completes bytes on the
// barrier at the same offset in each of them
__device__ __forceinline__ void tma_load_4d_mc(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                               int c3, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap* m, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask)
               : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ------------------------------------------------------------------------------------------------ epilogues
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
constexpr int kStagePitch = 33;                                  // floats per staged row: conflict-free both ways
constexpr int kStageFloats = 32 * kStagePitch;                   // one 32 x 32 fp32 chunk per epilogue warp

// Linear epilogue of one 32-row x 32-column accumulator chunk (this warp's TMEM lane quarter).
// The accumulators arrive one ROW per lane (tcgen05.ld 32x32b); storing them like that issues 16-byte writes to 32
// different rows per instruction (measured: ~12 us per 128x256 tile, the whole GEMM was epilogue-bound). They are
// transposed through a per-warp shared-memory tile instead, so every store instruction writes whole 64 / 128-byte row
// segments; bias, activation, accumulation and the bf16 conversion happen on the transposed side.
__device__ __forceinline__ void epi_linear(const GemmParams& p, int b, int row0, int col0, const uint32_t (&r)[32], int ks,
                                           float* stage, int lane) {
  if (row0 >= p.M || col0 >= p.N) return;      // warp-uniform
  const uint32_t stage_s = smem_u32(stage);      // explicit shared-space accesses (the pointer's provenance is lost to
                                                 // the 1024-byte alignment arithmetic, generic ST.E/LD.E otherwise)
  const int nvalid = min(32, p.N - col0);
  // accumulating bf16 epilogue (C += acc: the dgrad GEMMs that add into an existing gradient): issue the four 16-byte loads of
  // the old values FIRST, so their latency hides under the shared-memory transpose below (loading them inside the store
  // loop exposed one global round trip per iteration: the accumulating dgrad took 80 us against 34 us for the plain one)
  const bool fast_bf16 = nvalid == 32 && p.row_map == nullptr && !p.out_f32 && ((p.ldc & 7) == 0) &&
                         ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) && ((p.c_batch & 7) == 0);
  uint4 oldv[4];
  if (fast_bf16 && p.beta != 0) {
    const __nv_bfloat16* cb0 = reinterpret_cast<const __nv_bfloat16*>(p.C) + (long long)b * p.c_batch + col0 + 8 * (lane & 3);
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int row = row0 + it * 8 + (lane >> 2);
      oldv[it] = row < p.M ? *reinterpret_cast<const uint4*>(cb0 + (long long)row * p.ldc) : make_uint4(0, 0, 0, 0);
    }
  }
#pragma unroll
  for (int j = 0; j < 32; ++j) sts_f32(stage_s + (lane * kStagePitch + j) * 4, __uint_as_float(r[j]));
  __syncwarp();
  const float* bias = (p.bias && ks == 0) ? p.bias + (long long)b * p.bias_batch : nullptr;
  // FAST PATHS (warp-uniform test): full 32-column chunk, no row permutation, 16-byte aligned rows. The generic path
  // below re-tests every runtime option per 16-byte store (~600 instructions per chunk and warp, measured with ncu);
  // these loops are ~5x shorter and cover every GEMM of the train step except ragged edges and the LSTM row maps.
  const int act = p.act;
  if (nvalid == 32 && p.row_map == nullptr) {
    if (fast_bf16) {
      __nv_bfloat16* cbase = reinterpret_cast<__nv_bfloat16*>(p.C) + (long long)b * p.c_batch + col0 + 8 * (lane & 3);
      float bv[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) bv[e] = bias != nullptr ? __ldg(bias + col0 + 8 * (lane & 3) + e) : 0.f;
      const uint32_t sbase = stage_s + ((lane >> 2) * kStagePitch + 8 * (lane & 3)) * 4;
      const bool accum = p.beta != 0;
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int row = row0 + it * 8 + (lane >> 2);
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = lds_f32(sbase + (it * 8 * kStagePitch + e) * 4) + bv[e];
        // the result is rounded to bf16 (ulp 2^-8): single-MUFU activations (2^-11) — the exact tanh (exp + divide per
        // element) made the view-attention projection epilogue-bound: 111 us against 45 us for the same product without it
        if (act == ACT_ELU) {
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = elu_fast(v[e]);
        } else if (act == ACT_TANH) {
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = tanh_fast(v[e]);
        }
        if (row < p.M) {
          uint4* c = reinterpret_cast<uint4*>(cbase + (long long)row * p.ldc);
          if (accum) {
            const uint4 old = oldv[it];
            const float2 a0 = unpack_bf16x2(old.x), a1 = unpack_bf16x2(old.y), a2 = unpack_bf16x2(old.z), a3 = unpack_bf16x2(old.w);
            v[0] += a0.x; v[1] += a0.y; v[2] += a1.x; v[3] += a1.y; v[4] += a2.x; v[5] += a2.y; v[6] += a3.x; v[7] += a3.y;
          }
          uint4 o;
          o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
          o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
          *c = o;
        }
      }
      __syncwarp();
      return;
    }
    if (p.out_f32 && ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) && ((p.c_batch & 3) == 0)) {
      float* cbase = reinterpret_cast<float*>(p.C) + (long long)b * p.c_batch + col0 + 4 * (lane & 7);
      float bv[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) bv[e] = bias != nullptr ? __ldg(bias + col0 + 4 * (lane & 7) + e) : 0.f;
      const uint32_t sbase = stage_s + ((lane >> 3) * kStagePitch + 4 * (lane & 7)) * 4;
      const int beta = p.beta;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int row = row0 + it * 4 + (lane >> 3);
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = lds_f32(sbase + (it * 4 * kStagePitch + e) * 4) + bv[e];
        if (act == ACT_ELU) {
#pragma unroll
          for (int e = 0; e < 4; ++e) v[e] = eluf_(v[e]);
        } else if (act == ACT_TANH) {
#pragma unroll
          for (int e = 0; e < 4; ++e) v[e] = tanhf_(v[e]);
        }
        if (row < p.M) {
          float* c = cbase + (long long)row * p.ldc;
          if (beta == 2) {
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(c), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3])
                         : "memory");
          } else {
            float4 o = make_float4(v[0], v[1], v[2], v[3]);
            if (beta == 1) {
              const float4 old = *reinterpret_cast<const float4*>(c);
              o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
            }
            *reinterpret_cast<float4*>(c) = o;
          }
        }
      }
      __syncwarp();
      return;
    }
  }
  // GENERIC PATH (ragged N, unaligned rows, row permutation). Loops stay rolled on purpose (code size).
  const int per = p.out_f32 ? 4 : 8;                    // elements per lane: one 16-byte store either way
  const int lanes_per_row = 32 / per;                    // 8 (fp32: 128-byte row segment) or 4 (bf16: 64-byte segment)
  const int q = lane % lanes_per_row, cq = per * q;
  const int rows_per_it = 32 / lanes_per_row;
  float bv[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) bv[e] = (bias != nullptr && e < per && cq + e < nvalid) ? __ldg(bias + col0 + cq + e) : 0.f;
#pragma unroll 1
  for (int it = 0; it < lanes_per_row; ++it) {
    const int rr = it * rows_per_it + lane / lanes_per_row, row = row0 + rr;
    if (row >= p.M) continue;
    const long long orow = p.row_map ? p.row_map[row] : row;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = (e < per) ? lds_f32(stage_s + (rr * kStagePitch + cq + e) * 4) + bv[e] : 0.f;
    if (act == ACT_ELU) {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = eluf_(v[e]);
    } else if (act == ACT_TANH) {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = tanhf_(v[e]);
    }
    if (p.out_f32) {
      float* c = reinterpret_cast<float*>(p.C) + (long long)b * p.c_batch + orow * p.ldc + col0 + cq;
      if (cq + 4 <= nvalid && (reinterpret_cast<uintptr_t>(c) & 15) == 0) {
        if (p.beta == 2) {
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(c), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3])
                       : "memory");
        } else {
          float4 o = make_float4(v[0], v[1], v[2], v[3]);
          if (p.beta) {
            const float4 old = *reinterpret_cast<const float4*>(c);
            o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
          }
          *reinterpret_cast<float4*>(c) = o;
        }
      } else {
#pragma unroll 1
        for (int e = 0; e < 4; ++e) {
          if (cq + e >= nvalid) break;
          if (p.beta == 2) atomicAdd(c + e, v[e]);
          else c[e] = p.beta ? c[e] + v[e] : v[e];
        }
      }
    } else {
      __nv_bfloat16* c = reinterpret_cast<__nv_bfloat16*>(p.C) + (long long)b * p.c_batch + orow * p.ldc + col0 + cq;
      if (cq + 8 <= nvalid && (reinterpret_cast<uintptr_t>(c) & 15) == 0) {
        if (p.beta) {
          const uint4 old = *reinterpret_cast<const uint4*>(c);
          const float2 a0 = unpack_bf16x2(old.x), a1 = unpack_bf16x2(old.y), a2 = unpack_bf16x2(old.z), a3 = unpack_bf16x2(old.w);
          v[0] += a0.x; v[1] += a0.y; v[2] += a1.x; v[3] += a1.y; v[4] += a2.x; v[5] += a2.y; v[6] += a3.x; v[7] += a3.y;
        }
        uint4 o;
        o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
        o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
        *reinterpret_cast<uint4*>(c) = o;
      } else {
#pragma unroll 1
        for (int e = 0; e < 8; ++e) {
          if (cq + e >= nvalid) break;
          float x = v[e];
          if (p.beta) x += __bfloat162float(c[e]);
          c[e] = __float2bfloat16_rn(x);
        }
      }
    }
  }
  __syncwarp();      // the staging tile is reused by this warp's next chunk
}

// LSTM cell forward on one row (sequence) and 8 hidden units (32 interleaved gate columns 4*j + {i,f,g,o}).
// kFused = false: the input product x W_ih^T + b was written to `gates` by a separate GEMM (per-step launches, p.s).
// kFused = true : the accumulator already holds x_t W_ih^T + h W_hh^T (lstm_seq_fwd_kernel); `bias` points at the 32 bias
//                 values of this chunk and the previous cell state is read around L1 (it was written by another SM).
template <bool kFused>
__device__ __forceinline__ void epi_lstm_fwd(const GemmParams& p, int dir, int s, int seq, int col0, const uint32_t (&r)[32],
                                             const float* __restrict__ bias) {
  if (seq >= p.M || col0 >= p.N) return;
  const int H = p.N >> 2;
  const int S = p.M;
  const int t = (dir & 1) == 0 ? s : p.T - 1 - s;
  const int j0 = col0 >> 2;
  __nv_bfloat16* g = p.gates + ((long long)t * p.M + seq) * p.gates_ld + (long long)dir * p.gates_dir + col0;
  const long long st = ((long long)dir * (p.T + 1) + s) * S * H + (long long)seq * H + j0;   // slot s
  const long long st1 = st + (long long)S * H;                                                // slot s+1
  const bool live = (p.seq_len == nullptr) || (t < p.seq_len[seq]);
  // blocked layout (fused mode): this warp's tile of gates / cell state, see gemm.cuh
  const int UG = H >> 3;
  __nv_bfloat16* gb = p.gates + lstm_blk_gates(t, dir, p.batch, p.RB, UG, seq >> 5, col0 >> 5) + (seq & 31) * 8;
  float* cb0 = p.c_hist + lstm_blk_c(dir, s, p.T, p.RB, UG, seq >> 5, col0 >> 5) + (seq & 31) * 4;
  float* cb1 = cb0 + (long long)p.RB * UG * 256;

  uint4 gin[4];
  float4 cp0, cp1;
  if (kFused) {
#pragma unroll
    for (int q = 0; q < 4; ++q) gin[q] = make_uint4(0, 0, 0, 0);
    // (whole-sequence kernels: the initial state, slot 0, is implicit zeros and never read — no fill launches per step)
    cp0 = s > 0 ? __ldcg(reinterpret_cast<const float4*>(cb0)) : make_float4(0.f, 0.f, 0.f, 0.f);
    cp1 = s > 0 ? __ldcg(reinterpret_cast<const float4*>(cb0 + 128)) : make_float4(0.f, 0.f, 0.f, 0.f);
  } else {
#pragma unroll
    for (int q = 0; q < 4; ++q) gin[q] = reinterpret_cast<const uint4*>(g)[q];
    cp0 = *reinterpret_cast<const float4*>(p.c_hist + st);
    cp1 = *reinterpret_cast<const float4*>(p.c_hist + st + 4);
  }
  float cprev[8] = {cp0.x, cp0.y, cp0.z, cp0.w, cp1.x, cp1.y, cp1.z, cp1.w};
  float cnew[8], hnew[8];
  uint4 gout[4];
  const uint32_t* gw = reinterpret_cast<const uint32_t*>(gin);
  uint32_t* go = reinterpret_cast<uint32_t*>(gout);
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    float2 a01, a23;
    if (kFused) {
      const float4 bv = __ldg(reinterpret_cast<const float4*>(bias) + u);
      a01 = make_float2(bv.x, bv.y);
      a23 = make_float2(bv.z, bv.w);
    } else {
      a01 = unpack_bf16x2(gw[2 * u]);
      a23 = unpack_bf16x2(gw[2 * u + 1]);
    }
    // single-MUFU tanh / sigmoid: the gates and h are stored as bf16, whose ulp (2^-8) dwarfs the 2^-11 approximation
    float ig = sigmoid_fast(__uint_as_float(r[4 * u + 0]) + a01.x);
    float fg = sigmoid_fast(__uint_as_float(r[4 * u + 1]) + a01.y);
    float gg = tanh_fast(__uint_as_float(r[4 * u + 2]) + a23.x);
    float og = sigmoid_fast(__uint_as_float(r[4 * u + 3]) + a23.y);
    float c = fg * cprev[u] + ig * gg;
    cnew[u] = c;
    hnew[u] = og * tanh_fast(c);
    go[2 * u] = pack_bf16x2(ig, fg);
    go[2 * u + 1] = pack_bf16x2(gg, og);
  }
  if (!live) {
    // padded step: state is carried unchanged, gates are zeroed so the backward pass sees no contribution
    uint4 hp = kFused ? (s > 0 ? __ldcg(reinterpret_cast<const uint4*>(p.h_hist + st)) : make_uint4(0, 0, 0, 0))
                      : *reinterpret_cast<const uint4*>(p.h_hist + st);
    const uint32_t* hw = reinterpret_cast<const uint32_t*>(&hp);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float2 h2 = unpack_bf16x2(hw[u]);
      hnew[2 * u] = h2.x; hnew[2 * u + 1] = h2.y;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) cnew[u] = cprev[u];
#pragma unroll
    for (int q = 0; q < 4; ++q) gout[q] = make_uint4(0, 0, 0, 0);
  }
  if (kFused) {
#pragma unroll
    for (int q = 0; q < 4; ++q) *reinterpret_cast<uint4*>(gb + q * 256) = gout[q];
    *reinterpret_cast<float4*>(cb1) = make_float4(cnew[0], cnew[1], cnew[2], cnew[3]);
    *reinterpret_cast<float4*>(cb1 + 128) = make_float4(cnew[4], cnew[5], cnew[6], cnew[7]);
  } else {
#pragma unroll
    for (int q = 0; q < 4; ++q) reinterpret_cast<uint4*>(g)[q] = gout[q];
    *reinterpret_cast<float4*>(p.c_hist + st1) = make_float4(cnew[0], cnew[1], cnew[2], cnew[3]);
    *reinterpret_cast<float4*>(p.c_hist + st1 + 4) = make_float4(cnew[4], cnew[5], cnew[6], cnew[7]);
  }
  uint4 hv;
  hv.x = pack_bf16x2(hnew[0], hnew[1]); hv.y = pack_bf16x2(hnew[2], hnew[3]);
  hv.z = pack_bf16x2(hnew[4], hnew[5]); hv.w = pack_bf16x2(hnew[6], hnew[7]);
  *reinterpret_cast<uint4*>(p.h_hist + st1) = hv;
  if (p.h_last != nullptr && s == p.T - 1)
    *reinterpret_cast<uint4*>(p.h_last + (long long)seq * p.h_last_ld + (long long)dir * H + j0) = hv;
  if (p.seq_out != nullptr) {
    uint4 ov = live ? hv : make_uint4(0, 0, 0, 0);
    *reinterpret_cast<uint4*>(p.seq_out + ((long long)seq * p.T + t) * p.seq_out_ld + (long long)dir * H + j0) = ov;
  }
}

// LSTM cell backward for 8 hidden units of one sequence at processed step s.
//   dh[8]      : gradient w.r.t. h after step s (everything already summed)
//   reads  gates (activated i,f,g,o), c_hist[s], c_hist[s+1], dc (running)
//   writes dgates (pre-activation) in place, dc <- dc_total * f
// kCoherent: the running buffers (dc, dh_carry) were written by ANOTHER CTA of the same launch (lstm_seq_bwd_kernel), so
//            they are read around L1 (an earlier step may have left a stale line of the same address there).
template <bool kCoherent>
__device__ __forceinline__ float4 ld_run4(const float* ptr) {
  return kCoherent ? __ldcg(reinterpret_cast<const float4*>(ptr)) : *reinterpret_cast<const float4*>(ptr);
}
// kSeq = false: per-step launches — row-major gates / c_hist / dc, dgates written in place of the gates.
// kSeq = true : whole-sequence kernels — blocked gates / c_hist / dc (gemm.cuh), running buffers read around L1, and the
//               gate gradients are RETURNED in gout (zeros for a padded step) for the caller to store in the standard
//               [T][S][D*4H] layout the TMA / wgrad operands need.
template <bool kSeq>
__device__ __forceinline__ void lstm_cell_bwd8(const GemmParams& p, int dir, int s, int seq, int j0, float (&dh)[8],
                                               uint4 (&gout)[4]) {
  const int H = p.N;   // bwd GEMM has N = H
  const int S = p.M;
  const int t = (dir & 1) == 0 ? s : p.T - 1 - s;
  const int UG = H >> 3;
  const __nv_bfloat16* g;
  const float *cprev_p, *cc_p;
  float* dcp;
  long long gq, cq;          // element stride between the 16-byte pieces of this thread
  if (kSeq) {
    g = p.gates + lstm_blk_gates(t, dir, p.batch, p.RB, UG, seq >> 5, j0 >> 3) + (seq & 31) * 8;
    cprev_p = p.c_hist + lstm_blk_c(dir, s, p.T, p.RB, UG, seq >> 5, j0 >> 3) + (seq & 31) * 4;
    cc_p = cprev_p + (long long)p.RB * UG * 256;
    dcp = p.dc + lstm_blk_dc(dir, p.RB, UG, seq >> 5, j0 >> 3) + (seq & 31) * 4;
    gq = 256; cq = 128;
  } else {
    g = p.gates + ((long long)t * p.M + seq) * p.gates_ld + (long long)dir * p.gates_dir + 4 * j0;
    const long long st = ((long long)dir * (p.T + 1) + s) * S * H + (long long)seq * H + j0;
    cprev_p = p.c_hist + st;
    cc_p = cprev_p + (long long)S * H;
    dcp = p.dc + ((long long)dir * S + seq) * H + j0;
    gq = 8; cq = 4;
  }
  const bool live = (p.seq_len == nullptr) || (t < p.seq_len[seq]);
  // issue every load of this group up front (one latency round)
  uint4 gin[4];
  float4 a0, a1, b0, b1, d0, d1;
  if (live) {
#pragma unroll
    for (int q = 0; q < 4; ++q) gin[q] = *reinterpret_cast<const uint4*>(g + q * gq);
    if (kSeq && s == 0) {       // slot 0 of the whole-sequence kernels is implicit zeros
      a0 = a1 = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
      a0 = *reinterpret_cast<const float4*>(cprev_p); a1 = *reinterpret_cast<const float4*>(cprev_p + cq);
    }
    b0 = *reinterpret_cast<const float4*>(cc_p); b1 = *reinterpret_cast<const float4*>(cc_p + cq);
    d0 = ld_run4<kSeq>(dcp); d1 = ld_run4<kSeq>(dcp + cq);
  }
  if (live && p.dh_ext != nullptr) {
    // gradient arriving on the per-step hidden output; padded steps emit constant zeros.
    // per-step path: [S][T][ld] (column dir*H); whole-sequence path: blocked [T][D][RB][H/8][32 rows][8] (coalesced)
    const __nv_bfloat16* ep = kSeq ? p.dh_ext + (((((long long)t * p.batch + dir) * p.RB + (seq >> 5)) * UG + (j0 >> 3)) * 32 + (seq & 31)) * 8
                                   : p.dh_ext + ((long long)seq * p.T + t) * p.seq_out_ld + (long long)dir * H + j0;
    uint4 ev = *reinterpret_cast<const uint4*>(ep);
    const uint32_t* ew = reinterpret_cast<const uint32_t*>(&ev);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float2 e2 = unpack_bf16x2(ew[u]);
      dh[2 * u] += e2.x; dh[2 * u + 1] += e2.y;
    }
  }
  if (p.dh_carry != nullptr) {
    // padded steps carry the state forward, so their incoming dh must reach the last live step unchanged
    // (running buffer: row-major [D][S][H] per step launch, blocked like dc in the whole-sequence kernels)
    float* cp = kSeq ? p.dh_carry + lstm_blk_dc(dir, p.RB, UG, seq >> 5, j0 >> 3) + (seq & 31) * 4
                     : p.dh_carry + ((long long)dir * S + seq) * H + j0;
    float4 c0 = ld_run4<kSeq>(cp), c1 = ld_run4<kSeq>(cp + cq);
    dh[0] += c0.x; dh[1] += c0.y; dh[2] += c0.z; dh[3] += c0.w;
    dh[4] += c1.x; dh[5] += c1.y; dh[6] += c1.z; dh[7] += c1.w;
    if (live) {
      *reinterpret_cast<float4*>(cp) = make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(cp + cq) = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
      *reinterpret_cast<float4*>(cp) = make_float4(dh[0], dh[1], dh[2], dh[3]);
      *reinterpret_cast<float4*>(cp + cq) = make_float4(dh[4], dh[5], dh[6], dh[7]);
    }
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) gout[q] = make_uint4(0, 0, 0, 0);
  if (!live) return;   // padded step: zero gate gradients (in place they already are: the forward wrote zeros); dc untouched
  float cprev[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
  float cc[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
  float dc[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
  const uint32_t* gw = reinterpret_cast<const uint32_t*>(gin);
  uint32_t* go = reinterpret_cast<uint32_t*>(gout);
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    float2 i_f = unpack_bf16x2(gw[2 * u]);
    float2 g_o = unpack_bf16x2(gw[2 * u + 1]);
    float ig = i_f.x, fg = i_f.y, gg = g_o.x, og = g_o.y;
    float tc = tanh_fast(cc[u]);
    float dct = dc[u] + dh[u] * og * (1.f - tc * tc);
    float d_o = dh[u] * tc * og * (1.f - og);
    float d_i = dct * gg * ig * (1.f - ig);
    float d_f = dct * cprev[u] * fg * (1.f - fg);
    float d_g = dct * ig * (1.f - gg * gg);
    dc[u] = dct * fg;
    go[2 * u] = pack_bf16x2(d_i, d_f);
    go[2 * u + 1] = pack_bf16x2(d_g, d_o);
  }
  if (!kSeq) {
    __nv_bfloat16* gw_ = const_cast<__nv_bfloat16*>(g);
#pragma unroll
    for (int q = 0; q < 4; ++q) reinterpret_cast<uint4*>(gw_)[q] = gout[q];
  }
  *reinterpret_cast<float4*>(dcp) = make_float4(dc[0], dc[1], dc[2], dc[3]);
  *reinterpret_cast<float4*>(dcp + cq) = make_float4(dc[4], dc[5], dc[6], dc[7]);
}

__device__ __forceinline__ void epi_lstm_bwd(const GemmParams& p, int dir, int seq, int col0, const uint32_t (&r)[32]) {
  if (seq >= p.M || col0 >= p.N) return;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float dh[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) dh[u] = __uint_as_float(r[8 * q + u]);
    uint4 gout[4];
    lstm_cell_bwd8<false>(p, dir, p.s, seq, col0 + 8 * q, dh, gout);
  }
}

// ------------------------------------------------------------------------------------------------ kernel
// kCluster == 2: CTA pairs (a thread-block cluster of 2 along M) share every B tile — each CTA fetches HALF of it and
// TMA-multicasts it into both shared memories, which cuts the L2 -> SMEM operand traffic per MMA by a third (the
// single-CTA kernel measured 36 % tensor-pipe activity at the ~6.3 KB/clk L2 throughput cap).
template <bool A_MN, bool B_MN, int BN, int kOcc, int kCluster>
__global__ void __launch_bounds__(kGemmThreads, (kOcc == 2) ? 2 : 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const GemmParams p) {
  using Cfg = TileCfg<BN, kOcc>;
  const uint32_t crank = (kCluster > 1) ? cluster_ctarank() : 0u;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  uint64_t* tile_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES + 256);   // [kGemmRing]
  int* tile_ring = reinterpret_cast<int*>(tile_bar + kGemmRing);                               // [kGemmRing]
  float* stage_base = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES + Cfg::BAR_BYTES);
  static_assert(STAGES + 4 <= kGemmRing, "tile ring too shallow");
  // dynamic schedule (p.tile_counter, zero on entry): tiles are claimed from a global counter instead of dealt round-robin, so
  // a CTA that gets its SM late (another kernel is still running there) simply takes fewer tiles — with the static deal the
  // whole launch waits for the slowest CTA's full share (measured: the W_ih weight gradient went from 0.70 to 1.13 ms when
  // 24 of its 148 CTAs started 0.5 ms late behind the question encoder's backward)
  const bool dyn = (kCluster == 1) && p.tile_counter != nullptr;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m_blocks = ((p.M + BM - 1) / BM + kCluster - 1) / kCluster;   // per-cluster M blocks (pairs when kCluster == 2)
  const int n_blocks = (p.N + BN - 1) / BN;
  const int k_blocks = (p.K + BK - 1) / BK;
  const int ksplit = p.ksplit > 1 ? p.ksplit : 1;
  const int tile0 = blockIdx.x / kCluster, tile_step = gridDim.x / kCluster;
  const int kb_per = (k_blocks + ksplit - 1) / ksplit;
  const int tiles_per_batch = m_blocks * n_blocks;
  const int num_tiles = tiles_per_batch * p.batch * ksplit;     // tile = ((ks * batch + b) * m_blocks + m) * n_blocks + n

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], kCluster);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], kEpiWarps);
    }
    for (int i = 0; i < kGemmRing; ++i) mbar_init(&tile_bar[i], 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  if (kCluster > 1) cluster_sync_all();     // the peer's barriers must be initialised before any multicast touches them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor prefetch) touched no global
  // memory and may overlap the tail of the previous kernel in the stream; from here on we read / write tensors, so wait
  // until the prerequisite grid has completed and flushed (a no-op when the launch carried no programmatic dependency)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");     // lets the next PDL kernel start ITS prologue early

  if (warp == 0 && lane == 0) {
    // ===================================================== TMA producer
    int stage = 0;
    uint32_t phase = 0;
    int next = dyn ? claim_tile(p.tile_counter, num_tiles) : 0;
    for (int it = 0;; ++it) {
      int tile;
      if (dyn) {
        tile = next;
        ring_publish<kGemmRing>(tile_ring, tile_bar, it, tile);
        if (tile < 0) break;
        next = claim_tile(p.tile_counter, num_tiles);       // one claim ahead: the atomic's latency hides under the loads
      } else {
        tile = tile0 + it * tile_step;
        if (tile >= num_tiles) break;
      }
      const int ks = tile / (tiles_per_batch * p.batch);
      const int t2 = tile - ks * tiles_per_batch * p.batch;
      const int b = t2 / tiles_per_batch;
      const int rem = t2 - b * tiles_per_batch;
      const int m_blk = (rem / n_blocks) * kCluster + (int)crank;
      const int n_blk = rem - (rem / n_blocks) * n_blocks;
      const int kb0 = ks * kb_per, kb1 = min(k_blocks, kb0 + kb_per);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        uint8_t* sA = smem + stage * Cfg::STAGE_BYTES;
        uint8_t* sB = sA + Cfg::A_BYTES;
        mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
        const int seg = kb / p.k_inner;
        const int kin = (kb - seg * p.k_inner) * BK;
        if (!A_MN) {
          // (K-major operands may be segmented too: k-block kb -> column kin of plane a_c2 + seg * step — the three bf16
          //  planes of a split fp32 operand, dvgr_split3; unsegmented: seg = 0, kin = kb * BK)
          tma_load_4d(sA, &tmA, &full_bar[stage], p.a_c0[b] + kin, m_blk * BM, p.a_c2[b] + seg * p.a_c2_step[b], p.a_c3[b]);
          if (p.prefetch_a) {
            // pull the A block this CTA needs for its NEXT tile from HBM into L2 one tile ahead (it is first-touch there)
            const int nxt = dyn ? next : tile + tile_step;
            if (nxt >= 0 && nxt < num_tiles) {
              const int nks = nxt / (tiles_per_batch * p.batch);
              const int nt2 = nxt - nks * tiles_per_batch * p.batch;
              const int nb = nt2 / tiles_per_batch;
              const int nm = ((nt2 - nb * tiles_per_batch) / n_blocks) * kCluster + (int)crank;
              if (nm != m_blk || nb != b || nks != ks)
                tma_prefetch_4d(&tmA, p.a_c0[nb] + (nks * kb_per + (kb - kb0)) * BK, nm * BM, p.a_c2[nb], p.a_c3[nb]);
            }
          }
        } else {
#pragma unroll
          for (int c = 0; c < BM / 64; ++c)
            tma_load_4d(sA + c * (64 * BK * 2), &tmA, &full_bar[stage], p.a_c0[b] + m_blk * BM + c * 64, kin,
                        p.a_c2[b] + seg * p.a_c2_step[b], p.a_c3[b]);
        }
        if (kCluster == 1) {
          if (!B_MN) {
            tma_load_4d(sB, &tmB, &full_bar[stage], p.b_c0[b] + kin, n_blk * BN, p.b_c2[b] + seg * p.b_c2_step[b], p.b_c3[b]);
          } else {
#pragma unroll
            for (int c = 0; c < BN / 64; ++c)
              tma_load_4d(sB + c * (64 * BK * 2), &tmB, &full_bar[stage], p.b_c0[b] + n_blk * BN + c * 64, kin,
                          p.b_c2[b] + seg * p.b_c2_step[b], p.b_c3[b]);
          }
        } else {   // this CTA fetches its half of the B tile and multicasts it to the pair
          constexpr uint16_t kMask = (1u << kCluster) - 1u;
          if (!B_MN) {
            const int half_rows = BN / kCluster;
            tma_load_4d_mc(sB + crank * (half_rows * BK * 2), &tmB, &full_bar[stage], p.b_c0[b] + kb * BK,
                           n_blk * BN + (int)crank * half_rows, p.b_c2[b], p.b_c3[b], kMask);
          } else {
#pragma unroll
            for (int c = 0; c < BN / 64 / kCluster; ++c) {
              const int cc = (int)crank * (BN / 64 / kCluster) + c;
              tma_load_4d_mc(sB + cc * (64 * BK * 2), &tmB, &full_bar[stage], p.b_c0[b] + n_blk * BN + cc * 64, kin,
                             p.b_c2[b] + seg * p.b_c2_step[b], p.b_c3[b], kMask);
            }
          }
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ===================================================== MMA issuer (single thread)
    constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, A_MN, B_MN);
    int stage = 0;
    uint32_t phase = 0;
    int as = 0;
    uint32_t aphase = 0;
    for (int it = 0;; ++it) {
      const int tile = dyn ? ring_consume<kGemmRing>(tile_ring, tile_bar, it) : tile0 + it * tile_step;
      if (tile < 0 || tile >= num_tiles) break;
      const int ks = tile / (tiles_per_batch * p.batch);
      const int kb0 = ks * kb_per, kb1 = min(k_blocks, kb0 + kb_per);
      mbar_wait(&tempty_bar[as], aphase ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * BN);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sA = smem_u32(smem + stage * Cfg::STAGE_BYTES);
        const uint32_t sB = sA + Cfg::A_BYTES;
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          const uint64_t da = A_MN ? umma_smem_desc(sA + k * 2048, 64 * BK * 2, 1024) : umma_smem_desc(sA + k * 32, 16, 1024);
          const uint64_t db = B_MN ? umma_smem_desc(sB + k * 2048, 64 * BK * 2, 1024) : umma_smem_desc(sB + k * 32, 16, 1024);
          umma_bf16(d_tmem, da, db, idesc, (kb > kb0 || k != 0) ? 1u : 0u);
        }
        if (kCluster == 1) umma_commit(&empty_bar[stage]);
        else umma_commit_mc(&empty_bar[stage], (1u << kCluster) - 1u);   // the slot is reusable once BOTH CTAs drained it
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
      umma_commit(&tfull_bar[as]);
      if (++as == 2) { as = 0; aphase ^= 1u; }
    }
  } else if (warp >= 4) {
    // ===================================================== epilogue: 8 warps, warp w reads TMEM lane quarter (w & 3)
    // (the hardware restricts a warp to lanes [32*(warp%4), +32)) and the even / odd 32-column chunks
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    int as = 0;
    uint32_t aphase = 0;
    for (int it = 0;; ++it) {
      const int tile = dyn ? ring_consume<kGemmRing>(tile_ring, tile_bar, it) : tile0 + it * tile_step;
      if (tile < 0 || tile >= num_tiles) break;
      const int ks = tile / (tiles_per_batch * p.batch);
      const int t2 = tile - ks * tiles_per_batch * p.batch;
      const int b = t2 / tiles_per_batch;
      const int rem = t2 - b * tiles_per_batch;
      const int m_blk = (rem / n_blocks) * kCluster + (int)crank;
      const int n_blk = rem - (rem / n_blocks) * n_blocks;
      const int kb0 = ks * kb_per, kb1 = min(k_blocks, kb0 + kb_per);
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const int row = m_blk * BM + q * 32 + lane;
      float* stage_w = stage_base + (warp - 4) * kStageFloats;
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(as * BN);
#pragma unroll 1
      for (int c = half; c < BN / 32; c += 2) {
        uint32_t r[32];
        tmem_ld_32x32(t_addr + c * 32, r);
        tmem_ld_wait();
        if (c + 2 >= BN / 32) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty_bar[as]);
        }
        const int col0 = n_blk * BN + c * 32;
        if (p.mode == EPI_LINEAR) epi_linear(p, b, m_blk * BM + q * 32, col0, r, ks, stage_w, lane);
        else if (p.mode == EPI_LSTM_FWD) epi_lstm_fwd<false>(p, b, p.s, row, col0, r, nullptr);
        else epi_lstm_bwd(p, b, row, col0, r);
      }
      if (++as == 2) { as = 0; aphase ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (kCluster > 1) cluster_sync_all();     // no CTA may retire while its peer can still multicast into its shared memory
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------ grouped weight gradients
// dW_i[rows_i][cols_i] += dy_i[M_i][rows_i]^T x_i[M_i][cols_i] for up to kMaxGroup independent problems in ONE persistent
// launch. The ~30 small weight-gradient GEMMs of a train step (768 x 768 outputs, reduction over 5-10 k rows) are off the
// critical path of the backward pass — nothing consumes them before the optimizer — but as separate launches each pays
// launch latency, prologue, pipeline fill and a partially filled last wave (36-144 tiles on 148 SMs): ~20 us apiece for
// ~5 us of tensor work. Queued and flushed together they become one tile stream that keeps every SM busy.
// Both operands are read MN-major (no transposed copies), fp32 output accumulated with vector reductions (split-K safe).
constexpr int kMaxGroup = 32;
struct GroupProb {
  int M, N;                     // output rows (= columns of dy), output columns (= columns of x)
  int m_blocks, n_blocks;       // 128-row / BN-column tiles
  int k_blocks, ksplit, kb_per; // 64-row blocks of the reduction, splits, blocks per split
  float* C;
  long long ldc;
};
struct GroupedParams {
  CUtensorMap ta[kMaxGroup], tb[kMaxGroup];
  GroupProb prob[kMaxGroup];
  int tile_start[kMaxGroup + 1];      // prefix sums of m_blocks * n_blocks * ksplit
  int n;
};

template <int BN>
__global__ void __launch_bounds__(kGemmThreads, 1) gemm_grouped_wgrad_kernel(const __grid_constant__ GroupedParams G) {
  using Cfg = TileCfg<BN, 1>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* stage_base = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES + 256);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total = G.tile_start[G.n];

  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], kEpiWarps); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  // tile g -> (problem, split, m block, n block); every role decodes the same way
  auto decode = [&](int g, int& pi, int& ks, int& m_blk, int& n_blk) {
    pi = 0;
    while (pi + 1 < G.n && g >= G.tile_start[pi + 1]) ++pi;
    const GroupProb& P = G.prob[pi];
    const int t = g - G.tile_start[pi];
    const int per = P.m_blocks * P.n_blocks;
    ks = t / per;
    const int rem = t - ks * per;
    m_blk = rem / P.n_blocks;
    n_blk = rem - m_blk * P.n_blocks;
  };

  if (warp == 0 && lane == 0) {
    // ===================================================== TMA producer
    int stage = 0;
    uint32_t phase = 0;
    for (int g = blockIdx.x; g < total; g += gridDim.x) {
      int pi, ks, m_blk, n_blk;
      decode(g, pi, ks, m_blk, n_blk);
      const GroupProb& P = G.prob[pi];
      const int kb0 = ks * P.kb_per, kb1 = min(P.k_blocks, kb0 + P.kb_per);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        uint8_t* sA = smem + stage * Cfg::STAGE_BYTES;
        uint8_t* sB = sA + Cfg::A_BYTES;
        mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
#pragma unroll
        for (int c = 0; c < BM / 64; ++c)
          tma_load_4d(sA + c * (64 * BK * 2), &G.ta[pi], &full_bar[stage], m_blk * BM + c * 64, kb * BK, 0, 0);
#pragma unroll
        for (int c = 0; c < BN / 64; ++c)
          tma_load_4d(sB + c * (64 * BK * 2), &G.tb[pi], &full_bar[stage], n_blk * BN + c * 64, kb * BK, 0, 0);
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ===================================================== MMA issuer (single thread)
    constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, true, true);
    int stage = 0;
    uint32_t phase = 0;
    int as = 0;
    uint32_t aphase = 0;
    for (int g = blockIdx.x; g < total; g += gridDim.x) {
      int pi, ks, m_blk, n_blk;
      decode(g, pi, ks, m_blk, n_blk);
      const GroupProb& P = G.prob[pi];
      const int kb0 = ks * P.kb_per, kb1 = min(P.k_blocks, kb0 + P.kb_per);
      mbar_wait(&tempty_bar[as], aphase ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * BN);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sA = smem_u32(smem + stage * Cfg::STAGE_BYTES);
        const uint32_t sB = sA + Cfg::A_BYTES;
#pragma unroll
        for (int k = 0; k < BK / 16; ++k)
          umma_bf16(d_tmem, umma_smem_desc(sA + k * 2048, 64 * BK * 2, 1024), umma_smem_desc(sB + k * 2048, 64 * BK * 2, 1024),
                    idesc, (kb > kb0 || k != 0) ? 1u : 0u);
        umma_commit(&empty_bar[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
      umma_commit(&tfull_bar[as]);
      if (++as == 2) { as = 0; aphase ^= 1u; }
    }
  } else if (warp >= 4) {
    // ===================================================== epilogue: fp32 vector reductions into the gradient buffers
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    int as = 0;
    uint32_t aphase = 0;
    for (int g = blockIdx.x; g < total; g += gridDim.x) {
      int pi, ks, m_blk, n_blk;
      decode(g, pi, ks, m_blk, n_blk);
      const GroupProb& P = G.prob[pi];
      GemmParams lp;                       // only the fields the linear epilogue reads
      lp.M = P.M; lp.N = P.N; lp.C = P.C; lp.ldc = P.ldc; lp.c_batch = 0; lp.out_f32 = 1; lp.act = ACT_NONE; lp.beta = 2;
      lp.bias = nullptr; lp.bias_batch = 0; lp.row_map = nullptr;
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      float* stage_w = stage_base + (warp - 4) * kStageFloats;
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(as * BN);
#pragma unroll 1
      for (int c = half; c < BN / 32; c += 2) {
        uint32_t r[32];
        tmem_ld_32x32(t_addr + c * 32, r);
        tmem_ld_wait();
        if (c + 2 >= BN / 32) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty_bar[as]);
        }
        epi_linear(lp, 0, m_blk * BM + q * 32, n_blk * BN + c * 32, r, ks, stage_w, lane);
      }
      if (++as == 2) { as = 0; aphase ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------ whole-sequence LSTM
// One persistent launch runs ALL T steps of every direction with the input projection fused in:
//   tile (s, d, m, n):  acc[128 x BN] = x_t[m] W_ih[d][n]^T  (K1, no dependency)  +  h_{s}[d][m] W_hh[d][n]^T  (H)
//   epilogue          :  + bias, LSTM cell, writes activated gates (kept for backward), c_{s+1}, h_{s+1}
// so the [T][S][D*4H] pre-activation tensor never exists in HBM (the two-kernel path wrote and re-read 503 MB of it at
// the SVQA shape) and the cell epilogue hides under a K = K1 + H main loop instead of a K = H one.
// Tiles are ordered step-major and dealt round-robin to the resident CTAs; the only cross-CTA dependency is
// "h_{s} of my 128 sequences is complete", tracked by one counter per (direction, m-block): every epilogue warp adds 1
// after its stores (release), the TMA producer of a dependent tile polls it (acquire) before loading h. A dependency is
// always ~one step (hundreds of tiles) behind the tile that needs it, so the poll is practically never taken; since every
// CTA walks its tiles in increasing order and all CTAs are co-resident (the launcher caps the grid), it cannot deadlock.
struct LstmSeqParams {
  int kb1, kb2;              // 64-wide k-blocks of the input part (ceil(K1 / 64)) and of the recurrent part (H / 64)
  int m_blocks, n_blocks;    // ceil(S / 128), 4H / BN
  const float* bias;         // [D][4H] fp32, gate-interleaved (b_ih + b_hh)
  int* flags;                // [D][m_blocks] zero at launch
  int* counter;              // zero at launch: tiles are CLAIMED from it in index order (dynamic schedule, see below)
  int* error;                // sticky: set when a dependency poll gave up (never in a healthy run)
};

// Dynamic tile schedule. Tiles are numbered step-major and a tile depends only on tiles with a SMALLER index. Instead of a
// static round-robin deal (which needs every CTA of the grid co-resident: a tile owned by a CTA that never got an SM would
// block its dependents forever), the TMA-producer thread of each CTA claims the next tile index from a global counter and
// publishes it to its MMA / epilogue warps through a small shared-memory ring. A claimed tile belongs to a CTA that is
// running by construction, every CTA works through its claims in increasing order, and a CTA claims at most one tile ahead of
// the one it is loading — so the smallest unfinished tile can always make progress, whatever else shares the GPU (a
// concurrent NCCL kernel, a second LSTM launch on another stream, a partially resident grid).
// Ring depth: the producer is at most STAGES tiles ahead of the MMA thread (every tile has >= 1 k-block), which is at most 2
// tiles (the TMEM stages) ahead of the slowest epilogue warp: kTileRing = 8 >= 4 + 2 + 2 slots are never overwritten early.
__device__ __forceinline__ int ld_acquire_gpu(const int* ptr) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(ptr) : "memory");
  return v;
}
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

// bounded acquire-poll: gives up (and says so in *error) instead of hanging the GPU if the protocol were ever violated
__device__ __forceinline__ void wait_flag(const int* flag, int need, int* error) {
  for (int it = 0; ld_acquire_gpu(flag) < need; ++it) {
    if ((it & 255) == 255) {
      if (*reinterpret_cast<volatile int*>(error) != 0) break;
      if (it > (1 << 21)) { atomicExch(error, 1); break; }
    }
  }
}

constexpr int kSeqEpiWarps = 16;                        // one per (TMEM lane quarter, column-chunk residue mod 4)
constexpr int kSeqThreads = 128 + 32 * kSeqEpiWarps;

template <int BN>
__global__ void __launch_bounds__(kSeqThreads, 1)
lstm_seq_fwd_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmWih,
                    const __grid_constant__ CUtensorMap tmH, const __grid_constant__ CUtensorMap tmWhh,
                    const GemmParams p, const LstmSeqParams q) {
  using Cfg = TileCfg<BN, 1>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  uint64_t* tile_bar = tempty_bar + 3;                          // [kTileRing]
  int* tile_ring = reinterpret_cast<int*>(tile_bar + kTileRing); // [kTileRing]
  static_assert(STAGES + 4 <= kTileRing, "tile ring too shallow for this pipeline depth");

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int H = p.N >> 2;
  const int ndir = p.batch;
  const int tiles_per_dir = q.m_blocks * q.n_blocks;
  const int tiles_per_step = tiles_per_dir * ndir;
  const int num_tiles = tiles_per_step * p.T;          // tile = ((s * D + d) * m_blocks + m) * n_blocks + n

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmWih); tma_prefetch_desc(&tmH); tma_prefetch_desc(&tmWhh);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], kSeqEpiWarps); }
    for (int i = 0; i < kTileRing; ++i) mbar_init(&tile_bar[i], 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && lane == 0) {
    // ===================================================== TMA producer
    int stage = 0;
    uint32_t phase = 0;
    int next = claim_tile(q.counter, num_tiles);
    for (int it = 0;; ++it) {
      const int tile = next;
      ring_publish(tile_ring, tile_bar, it, tile);
      if (tile < 0) break;
      next = claim_tile(q.counter, num_tiles);      // one claim ahead: the atomic's latency hides under this tile's loads
      const int s = tile / tiles_per_step;
      const int r0 = tile - s * tiles_per_step;
      const int d = r0 / tiles_per_dir;
      const int r1 = r0 - d * tiles_per_dir;
      const int m_blk = r1 / q.n_blocks, n_blk = r1 - m_blk * q.n_blocks;
      const int t = (d & 1) == 0 ? s : p.T - 1 - s;
      for (int kb = 0; kb < q.kb1; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        uint8_t* sA = smem + stage * Cfg::STAGE_BYTES;
        mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
        tma_load_4d(sA, &tmX, &full_bar[stage], kb * BK, m_blk * BM, t, 0);
        tma_load_4d(sA + Cfg::A_BYTES, &tmWih, &full_bar[stage], kb * BK, d * p.N + n_blk * BN, 0, 0);
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
      if (s > 0) {      // h_0 = 0: step 0 has no recurrent part
        wait_flag(q.flags + d * q.m_blocks + m_blk, kSeqEpiWarps * q.n_blocks * s, q.error);
        fence_proxy_async_global();     // the h rows were written through the generic proxy, TMA reads through the async one
        for (int kb = 0; kb < q.kb2; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* sA = smem + stage * Cfg::STAGE_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          tma_load_4d(sA, &tmH, &full_bar[stage], kb * BK, m_blk * BM, s, d);
          tma_load_4d(sA + Cfg::A_BYTES, &tmWhh, &full_bar[stage], kb * BK, n_blk * BN, d, 0);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ===================================================== MMA issuer (single thread)
    constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, false, false);
    int stage = 0;
    uint32_t phase = 0;
    int as = 0;
    uint32_t aphase = 0;
    for (int it = 0;; ++it) {
      const int tile = ring_consume(tile_ring, tile_bar, it);
      if (tile < 0) break;
      const int s = tile / tiles_per_step;
      const int nkb = q.kb1 + (s > 0 ? q.kb2 : 0);
      mbar_wait(&tempty_bar[as], aphase ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * BN);
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sA = smem_u32(smem + stage * Cfg::STAGE_BYTES);
        const uint32_t sB = sA + Cfg::A_BYTES;
#pragma unroll
        for (int k = 0; k < BK / 16; ++k)
          umma_bf16(d_tmem, umma_smem_desc(sA + k * 32, 16, 1024), umma_smem_desc(sB + k * 32, 16, 1024), idesc,
                    (kb > 0 || k != 0) ? 1u : 0u);
        umma_commit(&empty_bar[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
      umma_commit(&tfull_bar[as]);
      if (++as == 2) { as = 0; aphase ^= 1u; }
    }
  } else if (warp >= 4) {
    // ===================================================== epilogue: LSTM cell on the finished accumulator
    const int qd = warp & 3;
    const int c4 = (warp - 4) >> 2;          // this warp's 32-column chunks: c4, c4 + 4, ...
    int as = 0;
    uint32_t aphase = 0;
    for (int it = 0;; ++it) {
      const int tile = ring_consume(tile_ring, tile_bar, it);
      if (tile < 0) break;
      const int s = tile / tiles_per_step;
      const int r0 = tile - s * tiles_per_step;
      const int d = r0 / tiles_per_dir;
      const int r1 = r0 - d * tiles_per_dir;
      const int m_blk = r1 / q.n_blocks, n_blk = r1 - m_blk * q.n_blocks;
      int* flag = q.flags + d * q.m_blocks + m_blk;
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      if (s > 0) {      // order this warp's reads of c_{s} / h_{s} (written by other CTAs) after their release
        if (lane == 0) (void)ld_acquire_gpu(flag);
        __syncwarp();
      }
      const int row = m_blk * BM + qd * 32 + lane;
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(qd * 32) << 16) + static_cast<uint32_t>(as * BN);
#pragma unroll 1
      for (int c = c4; c < BN / 32; c += 4) {
        uint32_t r[32];
        tmem_ld_32x32(t_addr + c * 32, r);
        tmem_ld_wait();
        if (c + 4 >= BN / 32) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty_bar[as]);
        }
        const int col0 = n_blk * BN + c * 32;
        epi_lstm_fwd<true>(p, d, s, row, col0, r, q.bias + (long long)d * p.N + col0);
      }
      // publish: h_{s+1} / c_{s+1} of this warp's rows and columns are in memory
      fence_proxy_async_global();
      __syncwarp();
      if (lane == 0) {
        __threadfence();
        atomicAdd(flag, 1);
      }
      if (++as == 2) { as = 0; aphase ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
  (void)H;
}

// ------------------------------------------------------------------------------------------------ whole-sequence LSTM backward
// Steps s = T-2 ... 0 of every direction in ONE persistent launch (step T-1 has no recurrent product: lstm_bwd_first_kernel):
//   tile (k, d, m, n), s = T-1-k:  dh_s[128 x 128] = dgates_{s+1}[d][m] (128 x 4H)  W_hh[d][:, n]  on tensor cores,
//   epilogue: + external dh, LSTM cell backward -> dgates_s written in place of the activated gates, running dc.
// Same dependency protocol as lstm_seq_fwd_kernel (a tile needs ALL gate columns of its 128 sequences from the step
// processed before), same step-major round-robin tile order. The per-step launches spent ~2/3 of their time in the
// memory-latency-bound cell epilogue with 8 warps per SM and lost another ~20 % to wave quantisation (240 tiles on 148
// SMs); here 16 epilogue warps (one per TMEM lane quarter x 32-column chunk) keep twice the loads in flight and the tile
// stream never drains between steps.
constexpr int kBwdEpiWarps = 16;
constexpr int kBwdThreads = 128 + 32 * kBwdEpiWarps;
constexpr int kBwdStages = 4;
constexpr int kBwdBN = 128;
constexpr int kBwdStageBytes = BM * BK * 2 + kBwdBN * BK * 2;
constexpr int kBwdOutBytes = kBwdEpiWarps * 32 * 128;      // per epilogue warp: 32 rows x 128 B of gate gradients (2 unit groups)
constexpr int kBwdSmemBytes = kBwdStages * kBwdStageBytes + 1024 + 256 + kBwdOutBytes;

__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}

__global__ void __launch_bounds__(kBwdThreads, 1)
lstm_seq_bwd_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmWhh,
                    const GemmParams p, const LstmSeqParams q) {
  constexpr int BN = kBwdBN;
  constexpr int STAGES = kBwdStages;
  constexpr int A_BYTES = BM * BK * 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * kBwdStageBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  uint64_t* tile_bar = tempty_bar + 3;                          // [kTileRing]
  int* tile_ring = reinterpret_cast<int*>(tile_bar + kTileRing); // [kTileRing]
  static_assert(STAGES + 4 <= kTileRing, "tile ring too shallow for this pipeline depth");
  uint8_t* out_stage = smem + STAGES * kBwdStageBytes + 256;      // [16 warps][32 rows][128 B], 16-byte pieces XOR-swizzled

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int H = p.N;                                   // backward product: N = H, K = 4H
  const int ndir = p.batch;
  const int tiles_per_dir = q.m_blocks * q.n_blocks;
  const int tiles_per_step = tiles_per_dir * ndir;
  const int num_tiles = tiles_per_step * (p.T - 1);    // tile = (((k-1) * D + d) * m_blocks + m) * n_blocks + n, k = 1..T-1

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmG); tma_prefetch_desc(&tmWhh); }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], kBwdEpiWarps); }
    for (int i = 0; i < kTileRing; ++i) mbar_init(&tile_bar[i], 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<2 * BN>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && lane == 0) {
    // ===================================================== TMA producer
    int stage = 0;
    uint32_t phase = 0;
    int next = claim_tile(q.counter, num_tiles);
    for (int it = 0;; ++it) {
      const int tile = next;
      ring_publish(tile_ring, tile_bar, it, tile);
      if (tile < 0) break;
      next = claim_tile(q.counter, num_tiles);
      const int k1 = tile / tiles_per_step;              // k - 1
      const int r0 = tile - k1 * tiles_per_step;
      const int d = r0 / tiles_per_dir;
      const int r1 = r0 - d * tiles_per_dir;
      const int m_blk = r1 / q.n_blocks, n_blk = r1 - m_blk * q.n_blocks;
      const int s1 = p.T - 1 - k1;                       // the step processed before this one (its dgates are the A operand)
      const int t1 = (d & 1) == 0 ? s1 : p.T - 1 - s1;
      if (k1 > 0) {
        wait_flag(q.flags + d * q.m_blocks + m_blk, kBwdEpiWarps * q.n_blocks * k1, q.error);
        fence_proxy_async_global();
      }
      for (int kb = 0; kb < q.kb1; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        uint8_t* sA = smem + stage * kBwdStageBytes;
        uint8_t* sB = sA + A_BYTES;
        mbar_arrive_expect_tx(&full_bar[stage], kBwdStageBytes);
        tma_load_4d(sA, &tmG, &full_bar[stage], d * 4 * H + kb * BK, m_blk * BM, t1, 0);
#pragma unroll
        for (int c = 0; c < BN / 64; ++c)
          tma_load_4d(sB + c * (64 * BK * 2), &tmWhh, &full_bar[stage], n_blk * BN + c * 64, kb * BK, d, 0);
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ===================================================== MMA issuer (single thread)
    constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, false, true);
    int stage = 0;
    uint32_t phase = 0;
    int as = 0;
    uint32_t aphase = 0;
    for (int it = 0;; ++it) {
      if (ring_consume(tile_ring, tile_bar, it) < 0) break;
      mbar_wait(&tempty_bar[as], aphase ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * BN);
      for (int kb = 0; kb < q.kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sA = smem_u32(smem + stage * kBwdStageBytes);
        const uint32_t sB = sA + A_BYTES;
#pragma unroll
        for (int k = 0; k < BK / 16; ++k)
          umma_bf16(d_tmem, umma_smem_desc(sA + k * 32, 16, 1024), umma_smem_desc(sB + k * 2048, 64 * BK * 2, 1024), idesc,
                    (kb > 0 || k != 0) ? 1u : 0u);
        umma_commit(&empty_bar[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
      umma_commit(&tfull_bar[as]);
      if (++as == 2) { as = 0; aphase ^= 1u; }
    }
  } else if (warp >= 4) {
    // ===================================================== epilogue: warp e = (TMEM lane quarter, 32-column chunk)
    const int e = warp - 4;
    const int qd = warp & 3;                 // hardware: a warp may only touch TMEM lanes [32 * (warp % 4), +32)
    const int chunk = e >> 2;                // 0..3 -> hidden units [chunk * 32, +32) of the tile
    int as = 0;
    uint32_t aphase = 0;
    for (int it = 0;; ++it) {
      const int tile = ring_consume(tile_ring, tile_bar, it);
      if (tile < 0) break;
      const int k1 = tile / tiles_per_step;
      const int r0 = tile - k1 * tiles_per_step;
      const int d = r0 / tiles_per_dir;
      const int r1 = r0 - d * tiles_per_dir;
      const int m_blk = r1 / q.n_blocks, n_blk = r1 - m_blk * q.n_blocks;
      const int s = p.T - 2 - k1;
      int* flag = q.flags + d * q.m_blocks + m_blk;
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      if (k1 > 0) {     // order this warp's reads of dc / dh_carry (written by other CTAs) after their release
        if (lane == 0) (void)ld_acquire_gpu(flag);
        __syncwarp();
      }
      const int seq = m_blk * BM + qd * 32 + lane;
      const int j_base = n_blk * BN + chunk * 32;
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(qd * 32) << 16) + static_cast<uint32_t>(as * BN + chunk * 32);
      // one 8-unit group at a time (8 accumulator registers live instead of 32); the accumulator stage is handed back to
      // the MMA warp after the last group's load — the cell epilogue, not the MMA, is this kernel's critical resource.
      // Gate gradients go to the standard [T][S][D*4H] layout (TMA / wgrad operand): two groups (128 B per row) are
      // collected in this warp's swizzled shared-memory tile and written as whole 128-byte lines, 4 rows per instruction.
      const int t = (d & 1) == 0 ? s : p.T - 1 - s;
      const uint32_t ost = smem_u32(out_stage + e * (32 * 128));
      __nv_bfloat16* dgrow = p.dgates + ((long long)t * p.M + m_blk * BM + qd * 32) * p.gates_ld + (long long)d * p.gates_dir;
#pragma unroll 1
      for (int g = 0; g < 4; ++g) {
        uint32_t r[8];
        tmem_ld_32x8(t_addr + g * 8, r);
        tmem_ld_wait();
        if (g == 3) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty_bar[as]);
        }
        uint4 gout[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) gout[u] = make_uint4(0, 0, 0, 0);
        if (seq < p.M && j_base + g * 8 < H) {
          float dh[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) dh[u] = __uint_as_float(r[u]);
          lstm_cell_bwd8<true>(p, d, s, seq, j_base + g * 8, dh, gout);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int piece = (g & 1) * 4 + u;
          const uint32_t a = ost + lane * 128 + ((piece ^ (lane & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(gout[u].x), "r"(gout[u].y), "r"(gout[u].z), "r"(gout[u].w) : "memory");
        }
        if (g & 1) {
          __syncwarp();
          const int col = 4 * (j_base + (g - 1) * 8);            // first gate column of the group pair
          if (j_base + (g - 1) * 8 < H) {
            const bool second = j_base + g * 8 < H;               // H % 16 == 8 never happens (H % 64 == 0), kept for safety
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              const int row = it * 4 + (lane >> 3), piece = lane & 7;
              uint4 v;
              const uint32_t a = ost + row * 128 + ((piece ^ (row & 7)) << 4);
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
              if (m_blk * BM + qd * 32 + row < p.M && (piece < 4 || second))
                *reinterpret_cast<uint4*>(dgrow + (long long)row * p.gates_ld + col + piece * 8) = v;
            }
          }
          __syncwarp();
        }
      }
      fence_proxy_async_global();
      __syncwarp();
      if (lane == 0) {
        __threadfence();
        atomicAdd(flag, 1);
      }
      if (++as == 2) { as = 0; aphase ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<2 * BN>(tmem_base);
  }
}

// First (s == T-1) backward step of the LSTM: no recurrent product yet, dh comes from the encoder output gradient.
// kSeq: whole-sequence path — blocked gates / c_hist / dc in, gate gradients out to p.dgates (standard layout).
template <bool kSeq>
__global__ void lstm_bwd_first_kernel(const GemmParams p, const __nv_bfloat16* __restrict__ dh_last, long long dh_ld) {
  const int H = p.N, S = p.M;
  const int groups = H / 8;
  const long long total = (long long)p.batch * S * groups;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    // per-step path: unit groups fastest (row-major tensors); whole-sequence path: sequences fastest (blocked tensors)
    int jg, seq, dir;
    if (kSeq) {
      seq = (int)(i % S);
      const long long rest = i / S;
      jg = (int)(rest % groups);
      dir = (int)(rest / groups);
    } else {
      jg = (int)(i % groups);
      const long long rest = i / groups;
      seq = (int)(rest % S);
      dir = (int)(rest / S);
    }
    float dh[8];
    if (dh_last != nullptr) {
      uint4 v = *reinterpret_cast<const uint4*>(dh_last + (long long)seq * dh_ld + (long long)dir * H + jg * 8);
      const uint32_t* w = reinterpret_cast<const uint32_t*>(&v);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float2 f = unpack_bf16x2(w[u]);
        dh[2 * u] = f.x; dh[2 * u + 1] = f.y;
      }
    } else {
#pragma unroll
      for (int u = 0; u < 8; ++u) dh[u] = 0.f;
    }
    uint4 gout[4];
    lstm_cell_bwd8<kSeq>(p, dir, p.s, seq, jg * 8, dh, gout);
    if (kSeq) {
      const int t = (dir & 1) == 0 ? p.s : p.T - 1 - p.s;
      __nv_bfloat16* o = p.dgates + ((long long)t * S + seq) * p.gates_ld + (long long)dir * p.gates_dir + 32 * jg;
#pragma unroll
      for (int q = 0; q < 4; ++q) reinterpret_cast<uint4*>(o)[q] = gout[q];
    }
  }
}

// Plain SIMT reference GEMM (fp32 accumulate) used ONLY by the test-suite to check the tcgen05 path on the device
// at sizes the CPU oracle cannot reach. Arbitrary element strides.
__global__ void gemm_ref_kernel(const __nv_bfloat16* A, long long a_rs, long long a_ks, const __nv_bfloat16* B,
                                long long b_rs, long long b_ks, float* C, long long ldc, int M, int N, int K) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = blockIdx.y * blockDim.y + threadIdx.y;
  if (m >= M || n >= N) return;
  float acc = 0.f;
  for (int k = 0; k < K; ++k)
    acc += __bfloat162float(A[m * a_rs + k * a_ks]) * __bfloat162float(B[n * b_rs + k * b_ks]);
  C[m * ldc + n] = acc;
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  });
  return fn;
}

struct TmKey {
  const void* ptr;
  long long dims[4];
  long long strides[4];
  int box0, box1;
  bool operator<(const TmKey& o) const { return memcmp(this, &o, sizeof(TmKey)) < 0; }
};

static int make_tensor_map(CUtensorMap* out, const dvgr_operand& op, int box0, int box1) {
  static std::map<TmKey, CUtensorMap> cache;
  static std::mutex mu;
  TmKey key;
  memset(&key, 0, sizeof(key));
  key.ptr = op.ptr;
  for (int i = 0; i < 4; ++i) {
    key.dims[i] = i < op.ndim ? op.dims[i] : 1;
    key.strides[i] = i < op.ndim ? op.strides[i] : 0;
  }
  key.box0 = box0;
  key.box1 = box1;
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return 0; }
  }
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return set_error("cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t gdim[4];
  cuuint64_t gstride[3];
  cuuint32_t box[4] = {(cuuint32_t)box0, (cuuint32_t)box1, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  for (int i = 0; i < 4; ++i) gdim[i] = (cuuint64_t)key.dims[i];
  long long str[4];
  str[0] = 1;
  for (int i = 1; i < 4; ++i) {
    long long s;
    if (i < op.ndim) {
      s = op.strides[i];
    } else {   // padded unit dimension: any valid 16-byte-multiple stride (its coordinate is always 0)
      s = str[i - 1] * key.dims[i - 1];
      if (s < 8) s = 8;
      s = (s + 7) / 8 * 8;
    }
    if (s <= 0 || (s * 2) % 16 != 0)
      return set_error("tensor-map stride (dim %d = %lld elements) is not a positive multiple of 16 bytes", i, s);
    str[i] = s;
    gstride[i - 1] = (cuuint64_t)s * 2;
  }
  if ((reinterpret_cast<uintptr_t>(op.ptr) & 15) != 0) return set_error("tensor-map base pointer not 16-byte aligned");
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(op.ptr), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  {
    std::lock_guard<std::mutex> lk(mu);
    if (cache.size() > 4096) cache.clear();
    cache[key] = *out;
  }
  return 0;
}

static int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

template <bool A_MN, bool B_MN, int BN, int kOcc = 1, int kCluster = 1>
static int launch_variant(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int max_ctas,
                          cudaStream_t stream) {
  using Cfg = TileCfg<BN, kOcc>;
  static bool attr_set = false;
  auto kern = gemm_tcgen05_kernel<A_MN, B_MN, BN, kOcc, kCluster>;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return set_error("cudaFuncSetAttribute(smem=%d): %s", Cfg::SMEM_BYTES, cudaGetErrorString(e));
    attr_set = true;
  }
  // the transpose tiles are only used by the linear epilogue; the LSTM epilogues keep that shared memory as L1 (their
  // row-per-thread state accesses depend on it: +0.5 ms per step when it was carved away)
  const int smem_bytes = Cfg::SMEM_BYTES - (p.mode == EPI_LINEAR ? 0 : Cfg::STAGING_BYTES);
  const int m_blocks = ((p.M + BM - 1) / BM + kCluster - 1) / kCluster, n_blocks = (p.N + BN - 1) / BN;
  const long long tiles = (long long)m_blocks * n_blocks * p.batch * (p.ksplit > 1 ? p.ksplit : 1);
  long long grid = std::min<long long>(tiles * kCluster, max_ctas > 0 ? max_ctas : num_sms() * (kOcc == 2 ? 2 : 1));
  grid = grid / kCluster * kCluster;
  if (grid <= 0) return 0;
  static const int pdl = getenv("DVGR_PDL") ? atoi(getenv("DVGR_PDL")) : 1;
  if (kCluster == 1 && pdl) {
    // programmatic stream serialisation: this kernel may begin (its prologue) before the previous one has drained
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kGemmThreads);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ta, tb, p);
    if (e != cudaSuccess) return set_error("gemm PDL launch failed: %s", cudaGetErrorString(e));
  } else if (kCluster == 1) {
    kern<<<(int)grid, kGemmThreads, smem_bytes, stream>>>(ta, tb, p);
  } else {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kGemmThreads);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kCluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ta, tb, p);
    if (e != cudaSuccess) return set_error("gemm cluster launch failed: %s", cudaGetErrorString(e));
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error("gemm launch failed: %s", cudaGetErrorString(e));
  return 0;
}

int gemm_dispatch(const dvgr_operand& A, const dvgr_operand& B, GemmParams p, int bn, int max_ctas, cudaStream_t stream) {
  if (p.M <= 0 || p.N <= 0 || p.batch <= 0) return 0;
  if (p.K <= 0) return set_error("gemm: K must be positive");
  if (p.batch > kMaxBatch) return set_error("gemm: batch %d exceeds %d", p.batch, kMaxBatch);
  if (bn != 128 && bn != 256) bn = (p.N > 128) ? 256 : 128;
  if (p.mode != EPI_LINEAR) bn = 128;
  if (p.k_inner <= 0) p.k_inner = INT_MAX;
  static const int prefetch = getenv("DVGR_GEMM_PREFETCH") ? atoi(getenv("DVGR_GEMM_PREFETCH")) : 0;
  p.prefetch_a = prefetch;
  static const int dbg = getenv("DVGR_GEMM_DEBUG") ? atoi(getenv("DVGR_GEMM_DEBUG")) : 0;
  p.debug = dbg;
  if (p.ksplit > 1) {
    if (p.mode != EPI_LINEAR || !p.out_f32 || p.beta != 2) return set_error("gemm: ksplit needs the fp32 atomic-accumulate epilogue (beta = 2)");
    const int k_blocks = (p.K + BK - 1) / BK;
    if (p.ksplit > k_blocks) p.ksplit = k_blocks;
    const int per = (k_blocks + p.ksplit - 1) / p.ksplit;
    p.ksplit = (k_blocks + per - 1) / per;       // no empty splits: an empty split would publish an unwritten accumulator
  }
  if (p.beta == 2 && !p.out_f32) return set_error("gemm: atomic accumulation needs an fp32 output");
  const bool a_mn = A.major != 0, b_mn = B.major != 0;
  CUtensorMap ta, tb;
  int rc = make_tensor_map(&ta, A, 64, a_mn ? 64 : BM);
  if (rc) return rc;
  static const int use_cluster = getenv("DVGR_GEMM_CLUSTER") ? atoi(getenv("DVGR_GEMM_CLUSTER")) : 0;   // measured r1: multicast at cluster size 2 does not raise throughput (L2 broadcast ~ unicast below cluster size 8); kept as an opt-in
  static const int lstm_occ = getenv("DVGR_LSTM_OCC") ? atoi(getenv("DVGR_LSTM_OCC")) : 8;   // tuning knob (bit0: fwd, bit1: bwd use 2 CTAs/SM)
  const bool lstm2 = (p.mode == EPI_LSTM_FWD && !a_mn && !b_mn && (lstm_occ & 1)) ||
                     (p.mode == EPI_LSTM_BWD && !a_mn && b_mn && (lstm_occ & 2));
  const bool cluster = use_cluster && !lstm2 && ((p.M + BM - 1) / BM >= 2);
  rc = make_tensor_map(&tb, B, 64, b_mn ? 64 : (cluster ? bn / 2 : bn));
  if (rc) return rc;
  if (p.mode == EPI_LSTM_FWD && !a_mn && !b_mn && (lstm_occ & 1)) return launch_variant<false, false, 128, 2, 1>(ta, tb, p, max_ctas, stream);
  if (p.mode == EPI_LSTM_BWD && !a_mn && b_mn && (lstm_occ & 2)) return launch_variant<false, true, 128, 2, 1>(ta, tb, p, max_ctas, stream);
  if (p.mode == EPI_LSTM_FWD && !a_mn && !b_mn && (lstm_occ & 4)) return launch_variant<false, false, 128, 3, 1>(ta, tb, p, max_ctas, stream);
  // (the 3-stage variant trades pipeline depth for L1: right for the appearance encoder's 5120 sequences, wrong for the
  //  question encoder's 256, whose 24 CTAs are bound by the latency of their 24 serial k-blocks -> deep pipeline there)
  if (p.mode == EPI_LSTM_BWD && !a_mn && b_mn && (lstm_occ & 8) && p.M >= 2048) return launch_variant<false, true, 128, 3, 1>(ta, tb, p, max_ctas, stream);
#define DVGR_LAUNCH(AM, BMJ, BNV)                                                         \
  do {                                                                                    \
    if (cluster) return launch_variant<AM, BMJ, BNV, 1, 2>(ta, tb, p, max_ctas, stream);  \
    return launch_variant<AM, BMJ, BNV, 1, 1>(ta, tb, p, max_ctas, stream);               \
  } while (0)
  if (!a_mn && !b_mn) { if (bn == 256) DVGR_LAUNCH(false, false, 256); else DVGR_LAUNCH(false, false, 128); }
  if (!a_mn && b_mn) { if (bn == 256) DVGR_LAUNCH(false, true, 256); else DVGR_LAUNCH(false, true, 128); }
  if (a_mn && b_mn) { if (bn == 256) DVGR_LAUNCH(true, true, 256); else DVGR_LAUNCH(true, true, 128); }
#undef DVGR_LAUNCH
  return set_error("gemm: operand layout combination (A MN-major, B K-major) is not instantiated");
}

// Host side of the grouped weight-gradient launch. probs: dy_i [M_i][rows_i] (row stride ld_dy), x_i [M_i][cols_i] bf16.
int gemm_grouped_wgrad(const dvgr_wgrad_problem* probs, int n, cudaStream_t stream) {
  constexpr int BN = 256;
  using Cfg = TileCfg<BN, 1>;
  auto kern = gemm_grouped_wgrad_kernel<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return set_error("cudaFuncSetAttribute(grouped wgrad, smem=%d): %s", Cfg::SMEM_BYTES, cudaGetErrorString(e));
    attr_set = true;
  }
  for (int base = 0; base < n; base += kMaxGroup) {
    const int cnt = std::min(kMaxGroup, n - base);
    GroupedParams G;
    memset(&G, 0, sizeof(G));
    G.n = cnt;
    int tiles = 0;
    for (int i = 0; i < cnt; ++i) {
      const dvgr_wgrad_problem& w = probs[base + i];
      if (!w.dy || !w.x || !w.out) return set_error("wgrad_grouped: problem %d has a null pointer", base + i);
      if (w.M <= 0 || w.rows <= 0 || w.cols <= 0) return set_error("wgrad_grouped: problem %d has an empty extent", base + i);
      dvgr_operand A, B;
      memset(&A, 0, sizeof(A)); memset(&B, 0, sizeof(B));
      A.ptr = w.dy; A.major = 1; A.ndim = 2; A.dims[0] = w.rows; A.dims[1] = w.M; A.strides[0] = 1; A.strides[1] = w.ld_dy;
      B.ptr = w.x; B.major = 1; B.ndim = 2; B.dims[0] = w.cols; B.dims[1] = w.M; B.strides[0] = 1; B.strides[1] = w.ld_x;
      int rc = make_tensor_map(&G.ta[i], A, 64, 64);
      if (!rc) rc = make_tensor_map(&G.tb[i], B, 64, 64);
      if (rc) return rc;
      GroupProb& P = G.prob[i];
      P.M = w.rows; P.N = w.cols; P.C = w.out; P.ldc = w.ldc;
      P.m_blocks = (w.rows + BM - 1) / BM;
      P.n_blocks = (w.cols + BN - 1) / BN;
      P.k_blocks = (w.M + BK - 1) / BK;
      // split long reductions so that one tile-unit is at most ~40 k-blocks (~18 us): even load over the 148 CTAs
      int ks = (P.k_blocks + 39) / 40;
      if (ks < 1) ks = 1;
      P.kb_per = (P.k_blocks + ks - 1) / ks;
      P.ksplit = (P.k_blocks + P.kb_per - 1) / P.kb_per;       // no empty splits
      G.tile_start[i] = tiles;
      tiles += P.m_blocks * P.n_blocks * P.ksplit;
    }
    G.tile_start[cnt] = tiles;
    const int grid = std::min(tiles, num_sms());
    if (grid <= 0) continue;
    kern<<<grid, kGemmThreads, Cfg::SMEM_BYTES, stream>>>(G);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error("grouped wgrad launch failed: %s", cudaGetErrorString(e));
    count_launch();
  }
  return 0;
}

// Whole-sequence fused LSTM forward (lstm_seq_fwd_kernel). p carries the EPI_LSTM fields with M = S, N = 4H, batch = D.
int lstm_seq_fwd_launch(const dvgr_operand& X, const dvgr_operand& Wih, const dvgr_operand& Hh, const dvgr_operand& Whh,
                        GemmParams p, int K1, const float* bias, int* sync, int max_ctas, cudaStream_t stream) {
  constexpr int BN = 256;
  using Cfg = TileCfg<BN, 1>;
  const int H = p.N / 4;
  if (p.N % BN != 0) return set_error("lstm_seq_fwd: 4H = %d must be a multiple of %d", p.N, BN);
  if (K1 <= 0 || K1 % 8 != 0) return set_error("lstm_seq_fwd: K1 = %d must be a positive multiple of 8", K1);
  CUtensorMap tx, twih, th, twhh;
  int rc = make_tensor_map(&tx, X, 64, BM);
  if (!rc) rc = make_tensor_map(&twih, Wih, 64, BN);
  if (!rc) rc = make_tensor_map(&th, Hh, 64, BM);
  if (!rc) rc = make_tensor_map(&twhh, Whh, 64, BN);
  if (rc) return rc;
  p.RB = (p.M + 31) / 32;
  LstmSeqParams q;
  q.kb1 = (K1 + BK - 1) / BK;
  q.kb2 = H / BK;
  q.m_blocks = (p.M + BM - 1) / BM;
  q.n_blocks = p.N / BN;
  q.bias = bias;
  q.flags = sync;
  q.counter = sync + (long long)p.batch * q.m_blocks;
  q.error = q.counter + 1;
  auto kern = lstm_seq_fwd_kernel<BN>;
  const int smem_bytes = Cfg::SMEM_BYTES - Cfg::STAGING_BYTES;
  static int max_resident = 0;
  if (max_resident == 0) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return set_error("cudaFuncSetAttribute(lstm_seq_fwd, smem=%d): %s", smem_bytes, cudaGetErrorString(e));
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kSeqThreads, smem_bytes);
    if (e != cudaSuccess || per_sm < 1) return set_error("lstm_seq_fwd: kernel cannot be resident (%s)", cudaGetErrorString(e));
    max_resident = per_sm * num_sms();     // the dependency protocol needs every CTA of the grid co-resident
  }
  const long long tiles = (long long)q.m_blocks * q.n_blocks * p.batch * p.T;
  // at most one step's tiles can run at the same time (each depends on the step before): more CTAs than that only take SMs
  // away from whatever runs next to this launch (the question encoder next to the appearance encoder: 48 of 148 SMs)
  const long long per_step = (long long)q.m_blocks * q.n_blocks * p.batch;
  int grid = (int)std::min<long long>(std::min(tiles, per_step), std::min(max_resident, num_sms()));
  if (max_ctas > 0) grid = std::min(grid, max_ctas);
  if (grid <= 0) return 0;
  kern<<<grid, kSeqThreads, smem_bytes, stream>>>(tx, twih, th, twhh, p, q);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error("lstm_seq_fwd launch failed: %s", cudaGetErrorString(e));
  return 0;
}

// Whole-sequence LSTM backward, steps T-2 ... 0 (lstm_seq_bwd_kernel). p: EPI_LSTM fields with M = S, N = H, batch = D.
int lstm_seq_bwd_launch(const dvgr_operand& G, const dvgr_operand& Whh, GemmParams p, int* sync, int max_ctas, cudaStream_t stream) {
  const int H = p.N;
  if (H % 64 != 0) return set_error("lstm_seq_bwd: H = %d must be a multiple of 64", H);
  if (p.T < 2) return 0;
  CUtensorMap tg, tw;
  int rc = make_tensor_map(&tg, G, 64, BM);
  if (!rc) rc = make_tensor_map(&tw, Whh, 64, 64);
  if (rc) return rc;
  LstmSeqParams q;
  q.kb1 = 4 * H / BK;
  q.kb2 = 0;
  q.m_blocks = (p.M + BM - 1) / BM;
  q.n_blocks = (H + kBwdBN - 1) / kBwdBN;
  q.bias = nullptr;
  q.flags = sync;
  q.counter = sync + (long long)p.batch * q.m_blocks;
  q.error = q.counter + 1;
  static int max_resident = 0;
  if (max_resident == 0) {
    cudaError_t e = cudaFuncSetAttribute(lstm_seq_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmemBytes);
    if (e != cudaSuccess) return set_error("cudaFuncSetAttribute(lstm_seq_bwd, smem=%d): %s", kBwdSmemBytes, cudaGetErrorString(e));
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lstm_seq_bwd_kernel, kBwdThreads, kBwdSmemBytes);
    if (e != cudaSuccess || per_sm < 1) return set_error("lstm_seq_bwd: kernel cannot be resident (%s)", cudaGetErrorString(e));
    max_resident = per_sm * num_sms();
  }
  const long long tiles = (long long)q.m_blocks * q.n_blocks * p.batch * (p.T - 1);
  const long long per_step = (long long)q.m_blocks * q.n_blocks * p.batch;      // see lstm_seq_fwd_launch
  int grid = (int)std::min<long long>(std::min(tiles, per_step), std::min(max_resident, num_sms()));
  if (max_ctas > 0) grid = std::min(grid, max_ctas);
  if (grid <= 0) return 0;
  lstm_seq_bwd_kernel<<<grid, kBwdThreads, kBwdSmemBytes, stream>>>(tg, tw, p, q);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error("lstm_seq_bwd launch failed: %s", cudaGetErrorString(e));
  return 0;
}

int lstm_bwd_first(const GemmParams& p, const void* dh_last, long long dh_ld, cudaStream_t stream) {
  const long long total = (long long)p.batch * p.M * (p.N / 8);
  if (total <= 0) return 0;
  int blocks = (int)std::min<long long>((total + 255) / 256, 148 * 8);
  if (p.RB > 0)
    lstm_bwd_first_kernel<true><<<blocks, 256, 0, stream>>>(p, reinterpret_cast<const __nv_bfloat16*>(dh_last), dh_ld);
  else
    lstm_bwd_first_kernel<false><<<blocks, 256, 0, stream>>>(p, reinterpret_cast<const __nv_bfloat16*>(dh_last), dh_ld);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error("lstm_bwd_first launch failed: %s", cudaGetErrorString(e));
  return 0;
}

int gemm_ref(const void* A, long long a_rs, long long a_ks, const void* B, long long b_rs, long long b_ks, float* C,
             long long ldc, int M, int N, int K, cudaStream_t stream) {
  dim3 blk(32, 8), grd((N + 31) / 32, (M + 7) / 8);
  gemm_ref_kernel<<<grd, blk, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(A), a_rs, a_ks,
                                            reinterpret_cast<const __nv_bfloat16*>(B), b_rs, b_ks, C, ldc, M, N, K);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error("gemm_ref launch failed: %s", cudaGetErrorString(e));
  return 0;
}

}  // namespace dvgr
