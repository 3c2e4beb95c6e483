// extern "C" entry points of libdualvgr_b200.so (declared in include/dualvgr_b200.h).
#include <atomic>
#include <limits.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "capi_internal.h"

namespace dvgr {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

int set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return 1;
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
static const unsigned long long* g_seed_off = nullptr;
const unsigned long long* seed_offset_ptr() { return g_seed_off; }
void g_seed_off_set(const unsigned long long* p) { g_seed_off = p; }

static void fill_lstm_params(GemmParams& p, const dvgr_lstm_args& a) {
  p.gates = reinterpret_cast<__nv_bfloat16*>(a.gates);
  p.gates_ld = (long long)a.ndir * 4 * a.H;
  p.gates_dir = 4LL * a.H;
  p.c_hist = a.c_hist;
  p.h_hist = reinterpret_cast<__nv_bfloat16*>(a.h_hist);
  p.h_last = reinterpret_cast<__nv_bfloat16*>(a.h_last);
  p.h_last_ld = a.h_last_ld;
  p.dc = a.dc;
  p.dh_carry = a.dh_carry;
  p.seq_len = a.seq_len;
  p.seq_out = reinterpret_cast<__nv_bfloat16*>(a.seq_out);
  p.seq_out_ld = a.seq_out_ld;
  p.T = a.T;
  p.s = a.s;
  p.batch = a.ndir;
  p.k_inner = INT_MAX;
}

static int check_lstm(const dvgr_lstm_args& a) {
  if (a.S <= 0 || a.T <= 0) return set_error("lstm: S and T must be positive");
  if (a.H <= 0 || a.H % 64 != 0) return set_error("lstm: H=%d must be a multiple of 64", a.H);
  if (a.ndir < 1 || a.ndir > kMaxBatch) return set_error("lstm: ndir=%d out of range", a.ndir);
  if (a.s < 0 || a.s >= a.T) return set_error("lstm: step %d out of range [0,%d)", a.s, a.T);
  if (!a.gates || !a.whh || !a.h_hist || !a.c_hist) return set_error("lstm: null buffer");
  return 0;
}

}  // namespace dvgr

using namespace dvgr;

extern "C" {

const char* dvgr_last_error(void) { return g_err; }
int dvgr_abi_version(void) { return DVGR_ABI_VERSION; }
long long dvgr_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
void dvgr_set_seed_offset(const unsigned long long* dev_ptr) { dvgr::g_seed_off_set(dev_ptr); }

int dvgr_gemm(const dvgr_gemm_args* a, void* stream) {
  if (!a) return set_error("dvgr_gemm: null args");
  if (!a->A.ptr || !a->B.ptr || !a->C) return set_error("dvgr_gemm: null operand");
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = a->M; p.N = a->N; p.K = a->K; p.batch = a->batch > 0 ? a->batch : 1;
  if (p.batch > kMaxBatch) return set_error("dvgr_gemm: batch %d > %d", p.batch, kMaxBatch);
  for (int i = 0; i < kMaxBatch; ++i) {
    p.a_c0[i] = a->a_c0[i]; p.a_c2[i] = a->a_c2[i]; p.a_c3[i] = a->a_c3[i];
    p.b_c0[i] = a->b_c0[i]; p.b_c2[i] = a->b_c2[i]; p.b_c3[i] = a->b_c3[i];
    p.a_c2_step[i] = a->a_c2_step[i]; p.b_c2_step[i] = a->b_c2_step[i];
  }
  p.k_inner = a->k_inner > 0 ? a->k_inner : INT_MAX;
  p.ksplit = a->ksplit;
  p.tile_counter = a->tile_counter;
  p.mode = EPI_LINEAR;
  p.C = a->C; p.ldc = a->ldc; p.c_batch = a->c_batch;
  p.out_f32 = a->out_f32; p.act = a->act; p.beta = a->beta;
  p.bias = a->bias; p.bias_batch = a->bias_batch; p.row_map = a->row_map;
  int rc = gemm_dispatch(a->A, a->B, p, a->bn, a->max_ctas, reinterpret_cast<cudaStream_t>(stream));
  if (rc == 0) count_launch();
  return rc;
}

int dvgr_gemm_reference(const void* A, long long a_rs, long long a_ks, const void* B, long long b_rs, long long b_ks,
                        float* C, long long ldc, int M, int N, int K, void* stream) {
  int rc = gemm_ref(A, a_rs, a_ks, B, b_rs, b_ks, C, ldc, M, N, K, reinterpret_cast<cudaStream_t>(stream));
  if (rc == 0) count_launch();
  return rc;
}

int dvgr_lstm_step_fwd(const dvgr_lstm_args* a, void* stream) {
  if (!a) return set_error("dvgr_lstm_step_fwd: null args");
  if (int rc = check_lstm(*a)) return rc;
  GemmParams p;
  memset(&p, 0, sizeof(p));
  fill_lstm_params(p, *a);
  p.mode = EPI_LSTM_FWD;
  p.M = a->S; p.N = 4 * a->H; p.K = a->H;
  dvgr_operand A, B;
  memset(&A, 0, sizeof(A)); memset(&B, 0, sizeof(B));
  // A = h_hist[d][s] : dims {H, S, T+1, D}
  A.ptr = a->h_hist; A.major = 0; A.ndim = 4;
  A.dims[0] = a->H; A.dims[1] = a->S; A.dims[2] = a->T + 1; A.dims[3] = a->ndir;
  A.strides[0] = 1; A.strides[1] = a->H; A.strides[2] = (long long)a->S * a->H; A.strides[3] = (long long)(a->T + 1) * a->S * a->H;
  // B = whh[d] : dims {H, 4H, D}
  B.ptr = a->whh; B.major = 0; B.ndim = 3;
  B.dims[0] = a->H; B.dims[1] = 4LL * a->H; B.dims[2] = a->ndir;
  B.strides[0] = 1; B.strides[1] = a->H; B.strides[2] = 4LL * a->H * a->H;
  for (int d = 0; d < a->ndir; ++d) { p.a_c2[d] = a->s; p.a_c3[d] = d; p.b_c2[d] = d; }
  int rc = gemm_dispatch(A, B, p, 128, 0, reinterpret_cast<cudaStream_t>(stream));
  if (rc == 0) count_launch();
  return rc;
}

// per-(direction, 128-sequence block) completion counters, the tile-claim counter, the sticky error word (last)
int dvgr_lstm_seq_sync_words(int S, int ndir) { return ndir * ((S + 127) / 128) + 2; }

int dvgr_lstm_seq_fwd(const dvgr_lstm_seq_args* a, void* stream) {
  if (!a) return set_error("dvgr_lstm_seq_fwd: null args");
  dvgr_lstm_args b = a->lstm;
  b.s = 0;
  if (int rc = check_lstm(b)) return rc;
  if (!a->x || !a->wih || !a->bias || !a->sync) return set_error("dvgr_lstm_seq_fwd: null buffer");
  if (a->x_ld % 8 != 0 || a->wih_ld % 8 != 0 || a->x_ld < a->K1 || a->wih_ld < a->K1)
    return set_error("dvgr_lstm_seq_fwd: row strides must be multiples of 8 elements and >= K1");
  GemmParams p;
  memset(&p, 0, sizeof(p));
  fill_lstm_params(p, b);
  p.mode = EPI_LSTM_FWD;
  p.M = b.S; p.N = 4 * b.H; p.K = b.H;
  dvgr_operand X, W, A, B;
  memset(&X, 0, sizeof(X)); memset(&W, 0, sizeof(W)); memset(&A, 0, sizeof(A)); memset(&B, 0, sizeof(B));
  // x [T][S][x_ld] : dims {K1, S, T}
  X.ptr = a->x; X.major = 0; X.ndim = 3;
  X.dims[0] = a->K1; X.dims[1] = b.S; X.dims[2] = b.T;
  X.strides[0] = 1; X.strides[1] = a->x_ld; X.strides[2] = (long long)b.S * a->x_ld;
  // W_ih [D*4H][wih_ld] : dims {K1, D*4H}
  W.ptr = a->wih; W.major = 0; W.ndim = 2;
  W.dims[0] = a->K1; W.dims[1] = (long long)b.ndir * 4 * b.H;
  W.strides[0] = 1; W.strides[1] = a->wih_ld;
  // h_hist [D][T+1][S][H] : dims {H, S, T+1, D}
  A.ptr = b.h_hist; A.major = 0; A.ndim = 4;
  A.dims[0] = b.H; A.dims[1] = b.S; A.dims[2] = b.T + 1; A.dims[3] = b.ndir;
  A.strides[0] = 1; A.strides[1] = b.H; A.strides[2] = (long long)b.S * b.H; A.strides[3] = (long long)(b.T + 1) * b.S * b.H;
  // W_hh [D][4H][H] : dims {H, 4H, D}
  B.ptr = b.whh; B.major = 0; B.ndim = 3;
  B.dims[0] = b.H; B.dims[1] = 4LL * b.H; B.dims[2] = b.ndir;
  B.strides[0] = 1; B.strides[1] = b.H; B.strides[2] = 4LL * b.H * b.H;
  int rc = lstm_seq_fwd_launch(X, W, A, B, p, a->K1, a->bias, a->sync, b.max_ctas, reinterpret_cast<cudaStream_t>(stream));
  if (rc == 0) count_launch();
  return rc;
}

int dvgr_lstm_step_bwd(const dvgr_lstm_args* a, void* stream) {
  if (!a) return set_error("dvgr_lstm_step_bwd: null args");
  if (int rc = check_lstm(*a)) return rc;
  if (!a->dc) return set_error("dvgr_lstm_step_bwd: dc is null");
  if (a->seq_len && !a->dh_carry) return set_error("dvgr_lstm_step_bwd: dh_carry is required with seq_len");
  GemmParams p;
  memset(&p, 0, sizeof(p));
  fill_lstm_params(p, *a);
  p.mode = EPI_LSTM_BWD;
  p.dh_ext = reinterpret_cast<const __nv_bfloat16*>(a->dh_seq);
  p.M = a->S; p.N = a->H; p.K = 4 * a->H;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (a->s == a->T - 1) {
    int rc = lstm_bwd_first(p, a->dh_last, a->dh_last_ld, st);
    if (rc == 0) count_launch();
    return rc;
  }
  // dh_s = dgates[s+1] * W_hh  : A = gates rows of the step processed after s, K = 4H inside the direction's column block
  dvgr_operand A, B;
  memset(&A, 0, sizeof(A)); memset(&B, 0, sizeof(B));
  A.ptr = a->gates; A.major = 0; A.ndim = 3;
  A.dims[0] = (long long)a->ndir * 4 * a->H; A.dims[1] = a->S; A.dims[2] = a->T;
  A.strides[0] = 1; A.strides[1] = A.dims[0]; A.strides[2] = A.dims[0] * a->S;
  // B = whh[d] as [K = 4H][N = H], N contiguous -> MN-major
  B.ptr = a->whh; B.major = 1; B.ndim = 3;
  B.dims[0] = a->H; B.dims[1] = 4LL * a->H; B.dims[2] = a->ndir;
  B.strides[0] = 1; B.strides[1] = a->H; B.strides[2] = 4LL * a->H * a->H;
  for (int d = 0; d < a->ndir; ++d) {
    const int s1 = a->s + 1;
    p.a_c0[d] = d * 4 * a->H;
    p.a_c2[d] = ((d & 1) == 0) ? s1 : a->T - 1 - s1;
    p.b_c2[d] = d;
  }
  int rc = gemm_dispatch(A, B, p, 128, 0, st);
  if (rc == 0) count_launch();
  return rc;
}

int dvgr_lstm_seq_bwd(const dvgr_lstm_args* a, void* dgates, int* sync, void* stream) {
  if (!a) return set_error("dvgr_lstm_seq_bwd: null args");
  dvgr_lstm_args b = *a;
  b.s = b.T - 1;
  if (int rc = check_lstm(b)) return rc;
  if (!b.dc || !sync || !dgates) return set_error("dvgr_lstm_seq_bwd: dc / sync / dgates is null");
  if (b.seq_len && !b.dh_carry) return set_error("dvgr_lstm_seq_bwd: dh_carry is required with seq_len");
  GemmParams p;
  memset(&p, 0, sizeof(p));
  fill_lstm_params(p, b);
  p.mode = EPI_LSTM_BWD;
  p.dh_ext = reinterpret_cast<const __nv_bfloat16*>(b.dh_seq);
  p.M = b.S; p.N = b.H; p.K = 4 * b.H;
  p.RB = (b.S + 31) / 32;                                        // blocked gates / c_hist / dc (gemm.cuh)
  p.dgates = reinterpret_cast<__nv_bfloat16*>(dgates);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int rc = lstm_bwd_first(p, b.dh_last, b.dh_last_ld, st);      // step T-1: no recurrent product
  if (rc) return rc;
  count_launch();
  if (b.T < 2) return 0;
  dvgr_operand A, B;
  memset(&A, 0, sizeof(A)); memset(&B, 0, sizeof(B));
  A.ptr = dgates; A.major = 0; A.ndim = 3;                        // the gate gradients of the step processed before
  A.dims[0] = (long long)b.ndir * 4 * b.H; A.dims[1] = b.S; A.dims[2] = b.T;
  A.strides[0] = 1; A.strides[1] = A.dims[0]; A.strides[2] = A.dims[0] * b.S;
  B.ptr = b.whh; B.major = 1; B.ndim = 3;
  B.dims[0] = b.H; B.dims[1] = 4LL * b.H; B.dims[2] = b.ndir;
  B.strides[0] = 1; B.strides[1] = b.H; B.strides[2] = 4LL * b.H * b.H;
  rc = lstm_seq_bwd_launch(A, B, p, sync, b.max_ctas, st);
  if (rc == 0) count_launch();
  return rc;
}

int dvgr_wgrad_grouped(const dvgr_wgrad_problem* probs, int n, void* stream) {
  if (n <= 0) return 0;
  if (!probs) return set_error("dvgr_wgrad_grouped: null problem list");
  return gemm_grouped_wgrad(probs, n, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
