/* dualvgr_b200.h — C ABI of libdualvgr_b200.so: the sm_100a kernels behind the DualVGR reasoning core.
 *
 * The reference (NJUPT-MCC/DualVGR-VideoQA) has no FFI layer: its boundary is the PyTorch nn.Module contract of
 * model/models.py (DualVGR, DualVGRUnit_multiple). Each entry point below replaces the stock ATen call sequence of one
 * reference call site (cited per function as <file>:<lines>, paths relative to the reference root) and is what a
 * ctypes / cffi binding on the reference side would bind (INTEGRATION.md shows the stub).
 *
 * Conventions
 *   - plain pointers + extents; every pointer is a DEVICE pointer unless stated otherwise
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous on that stream, never allocates device
 *     memory, never synchronises; outputs and workspaces are allocated by the caller
 *   - activations are bf16 (`uint16_t` storage) unless the name says f32; parameters and their gradients are fp32;
 *     bf16 copies of the weights are produced by dvgr_cast_* once per optimizer step
 *   - return value: 0 on success, non-zero on error; dvgr_last_error() returns a thread-local message
 *   - stateless and re-entrant (an internal cache of TMA descriptors is mutex-protected)
 */
#ifndef DUALVGR_B200_H_
#define DUALVGR_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DVGR_ABI_VERSION 1

const char* dvgr_last_error(void);
int dvgr_abi_version(void);
/* Number of kernels launched by this library in the calling process so far (bench.py reports the delta). */
long long dvgr_launch_count(void);

/* ------------------------------------------------------------------------------------------------------------------
 * GEMM family (tcgen05 + TMEM + TMA).  D[b][m][n] = epi( sum_k A[b][m][k] * B[b][n][k] ), bf16 operands, fp32 accumulate.
 * An operand is described by up to 4 dims, dims[0] contiguous:
 *   major == 0 (K-major):  dims = {K, rows, d2, d3}     the tensor as nn.Linear sees it (x[M,K], W[N,K])
 *   major == 1 (MN-major): dims = {rows, K, d2, d3}     the same storage seen by dgrad / wgrad (no transposed copies)
 * strides are in ELEMENTS (strides[0] must be 1, all others multiples of 8).
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct dvgr_operand {
  const void* ptr;
  int major;
  int ndim;
  long long dims[4];
  long long strides[4];
} dvgr_operand;

enum { DVGR_ACT_NONE = 0, DVGR_ACT_ELU = 1, DVGR_ACT_TANH = 2 };

typedef struct dvgr_gemm_args {
  dvgr_operand A, B;
  int M, N, K, batch;                 /* batch <= 4 */
  int a_c0[4], a_c2[4], a_c3[4];      /* per-batch coordinate offsets into dims 0, 2, 3 of A */
  int b_c0[4], b_c2[4], b_c3[4];
  int k_inner;                        /* MN-major reductions over [segments][rows]: 64-row blocks per segment, 0 = flat */
  int a_c2_step[4], b_c2_step[4];     /* dims-2 coordinate increment per segment */
  /* linear epilogue:  C = act(acc + bias) (+ C when beta)                                  */
  void* C;
  long long ldc, c_batch;             /* elements */
  int out_f32, act, beta;
  const float* bias;
  long long bias_batch;
  const int* row_map;                 /* optional output-row permutation (wgrad of gate-interleaved LSTM weights) */
  int bn;                             /* N tile: 128, 256, or 0 = choose */
  int max_ctas;                       /* 0 = one CTA per SM */
} dvgr_gemm_args;

/* Replaces every nn.Linear forward / dgrad / wgrad on the path: model/models.py:46,74 (motion projection),
 * model/GraphNN.py:96 (GAT head projections), model/Attention.py:14-18 (view-attention MLP),
 * model/fusions/fusions.py:420-449 (MFB), model/AnswerDecoder.py:173-200 (read-out, classifier),
 * model/utils.py:68 (QueryAttn.feat_enhance), and the W_ih product of nn.LSTM at model/Preprocessing.py:227. */
int dvgr_gemm(const dvgr_gemm_args* args, void* stream);

/* Test-only SIMT reference product (fp32 out), arbitrary element strides:  C[m][n] = sum_k A[m*a_rs + k*a_ks] * B[n*b_rs + k*b_ks] */
int dvgr_gemm_reference(const void* A, long long a_rs, long long a_ks, const void* B, long long b_rs, long long b_ks,
                        float* C, long long ldc, int M, int N, int K, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Fused LSTM recurrence (nn.LSTM(2048, 384, bidirectional) at model/Preprocessing.py:201,227; also usable for the two
 * question BiLSTMs at model/Preprocessing.py:97-101).  One call = one time step of BOTH directions:
 *   recurrent GEMM h_{s} W_hh^T on tensor cores, cell update in the epilogue.
 * Layouts (H hidden, S sequences, T steps, D = ndir directions, gate columns interleaved 4*j + {i,f,g,o}):
 *   gates  [T][S][D*4H] bf16 : in = x_t W_ih^T + b (from dvgr_gemm), out = activated gates (kept for backward);
 *                              after dvgr_lstm_step_bwd it holds the pre-activation gate gradients
 *   h_hist [D][T+1][S][H] bf16, c_hist [D][T+1][S][H] f32 : slot 0 = initial state (zeros), slot s+1 = state after step s
 *   whh    [D][4H][H] bf16 (rows interleaved like the gate columns)
 *   direction d processes time t = s (d = 0) or T-1-s (d = 1) at step s
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct dvgr_lstm_args {
  int S, H, T, ndir, s;
  void* gates;
  const void* whh;
  void* h_hist;
  float* c_hist;
  void* h_last;             /* optional [S][h_last_ld] bf16: final hidden of direction d at column d*H (written at s = T-1) */
  long long h_last_ld;
  const int* seq_len;       /* optional [S] int32: steps with t >= len carry the state and emit zeros */
  void* seq_out;            /* optional [S][T][seq_out_ld] bf16 per-step hidden (column d*H) */
  long long seq_out_ld;
  /* backward only */
  float* dc;                /* [D][S][H] f32 running cell-state gradient, zero before the first backward step */
  const void* dh_last;      /* [S][dh_last_ld] bf16 gradient of h_last (used at s = T-1) */
  long long dh_last_ld;
  const void* dh_seq;       /* optional [S][T][seq_out_ld] bf16 gradient of seq_out */
} dvgr_lstm_args;

int dvgr_lstm_step_fwd(const dvgr_lstm_args* args, void* stream);
/* Backward of step s; must be called for s = T-1, T-2, ..., 0. */
int dvgr_lstm_step_bwd(const dvgr_lstm_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DUALVGR_B200_H_ */
