// Shared declarations for the tcgen05 GEMM family (see gemm.cu).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace dvgr {

constexpr int kMaxBatch = 4;

enum EpiMode : int {
  EPI_LINEAR = 0,    // C = act(acc + bias[col]) (+ C if beta), bf16 or fp32 out, optional row map
  EPI_LSTM_FWD = 1,  // LSTM cell forward fused on the recurrent GEMM  (columns interleaved 4*j+gate)
  EPI_LSTM_BWD = 2,  // LSTM cell backward fused on the dh = dgates * W_hh GEMM
};
enum ActMode : int { ACT_NONE = 0, ACT_ELU = 1, ACT_TANH = 2 };

// Plain-old-data parameter block, passed by value to the kernel.
struct GemmParams {
  int M, N, K, batch;
  // per-batch TMA coordinate offsets of A and B: dim 0 (contiguous), dim 2, dim 3
  int a_c0[kMaxBatch], a_c2[kMaxBatch], a_c3[kMaxBatch];
  int b_c0[kMaxBatch], b_c2[kMaxBatch], b_c3[kMaxBatch];
  // MN-major reduction segmentation: k-block kb -> (row = (kb % k_inner) * 64, c2 += (kb / k_inner) * step[b])
  int k_inner;   // number of 64-row k-blocks per segment (INT_MAX when unsegmented)
  int a_c2_step[kMaxBatch], b_c2_step[kMaxBatch];

  int debug;        // profiling aid (DVGR_GEMM_DEBUG): 1 = skip global stores, 2 = skip the whole transposed phase
  int prefetch_a;   // producer prefetches the next tile's A blocks into L2 (tuning knob)
  int ksplit;    // K is split over `ksplit` CTAs per output tile (requires beta == 2 on a zero-initialised / accumulating C)
  int* tile_counter;   // optional, zero on entry: dynamic tile schedule (gemm.cu)

  int mode;
  // ---- EPI_LINEAR
  void* C;
  long long ldc;        // elements
  long long c_batch;    // elements between batches
  int out_f32;          // 1: float output, 0: bf16
  int act;
  int beta;             // 1: accumulate into C (read-modify-write) ; 2: atomic add (fp32 output, split-K safe)
  const float* bias;    // [N] or null ; bias + batch * bias_batch
  long long bias_batch;
  const int* row_map;   // optional: output row = row_map[row]

  // ---- EPI_LSTM_*  (H = N/4 for fwd, N for bwd; S = M sequences; T steps; step index s)
  __nv_bfloat16* gates;     // [T][S][gates_ld] pre-activations in, activated gates out (fwd); dgates out (bwd)
  long long gates_ld;       // elements per (t,seq) row
  long long gates_dir;      // column offset between directions
  float* c_hist;            // [dir][T+1][S][H]
  __nv_bfloat16* h_hist;    // [dir][T+1][S][H]
  __nv_bfloat16* h_last;    // optional [S, h_last_ld]: final hidden, written at column dir*H when s == T-1
  long long h_last_ld;
  float* dc;                // [dir][S][H] running cell gradient (bwd)
  float* dh_carry;          // optional [dir][S][H]: dh passed through padded steps (bwd, with seq_len)
  const __nv_bfloat16* dh_ext;  // optional extra dh for this step (unused for the appearance encoder)
  const int* seq_len;       // optional [S]: steps at t >= len keep the state (question encoder)
  __nv_bfloat16* seq_out;   // optional [S][T][seq_out_ld] per-step hidden output (zero at padded steps)
  long long seq_out_ld;
  int T, s;
  // ---- whole-sequence kernels (lstm_seq_*): tensors that only the cell epilogues touch use the blocked layout below
  int RB;                   // ceil(S / 32) row blocks; 0 = standard layouts (per-step launches)
  __nv_bfloat16* dgates;    // backward output [T][S][gates_ld] (standard layout: TMA / wgrad operand); null = in place
};

// Blocked ("warp tile") layout of the activated gates, the cell states and the running cell gradient in the whole-sequence
// kernels. A warp of a cell epilogue owns 32 consecutive sequences (TMEM lanes) x 8 hidden units; in row-major storage its
// 16-byte pieces are 12 KB apart, so every warp load / store touched 32 separate sectors (ncu: 32 sectors per request,
// L1/LSU wavefronts the top unit of both kernels). Here the pieces of such a tile are stored [piece][row][16 B]: one warp
// instruction = 512 contiguous bytes = 4 full lines.
//   gates_blk [T][D][RB][H/8][4 pieces][32 rows][8 bf16]     piece q = units 2q, 2q+1 x (i, f, g, o)
//   c_blk     [D][T+1][RB][H/8][2 pieces][32 rows][4 f32]    piece q = units 4q .. 4q+3
//   dc_blk    [D][RB][H/8][2 pieces][32 rows][4 f32]
__host__ __device__ inline long long lstm_blk_gates(int t, int d, int D, int RB, int UG, int rb, int ug) {
  return ((((long long)t * D + d) * RB + rb) * UG + ug) * 1024;       // elements (bf16); piece q at + q * 256, row r at + r * 8
}
__host__ __device__ inline long long lstm_blk_c(int d, int slot, int T, int RB, int UG, int rb, int ug) {
  return ((((long long)d * (T + 1) + slot) * RB + rb) * UG + ug) * 256;  // elements (f32); piece q at + q * 128, row r at + r * 4
}
__host__ __device__ inline long long lstm_blk_dc(int d, int RB, int UG, int rb, int ug) {
  return (((long long)d * RB + rb) * UG + ug) * 256;
}

}  // namespace dvgr
