// fp32 mode: the recurrent cells of the three BiLSTM encoders (reference model/Preprocessing.py:97-101,202; nn.LSTM's cell
// equations, PyTorch docs) in fp32, one launch per time step of ALL directions. The matrix products around the cell —
// x W_ih^T for the whole sequence, h_{t-1} W_hh^T per step, and their gradients — run on the tcgen05 GEMM through three bf16
// planes per operand (dvgr_split3 / k_inner plane segmentation), so the only precision loss of the recurrence is the
// < 2e-5 of a split product; this file keeps the state, the gate non-linearities and the length masking in fp32.
//
// Gate order is nn.LSTM's (i | f | g | o blocks of H columns), directions with an odd index run the sequence backwards.
// Sequences shorter than T (question encoder) behave like a packed sequence: a step at a padded position neither changes
// the state nor emits an output (seq_out is zero there), and receives no gradient.
#include <cuda_bf16.h>

#include "capi_internal.h"

namespace dvgr {
namespace {

__device__ __forceinline__ float sigmoid_exact(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ void store_planes(__nv_bfloat16* p, long long plane, float v) {
  const __nv_bfloat16 hi = __float2bfloat16_rn(v);
  p[0] = __float2bfloat16_rn(v - __bfloat162float(hi));
  p[plane] = hi;
  p[2 * plane] = hi;
}

__global__ void __launch_bounds__(256)
lstm32_cell_fwd_kernel(dvgr_lstm32_args a) {
  const long long per_dir = (long long)a.S * a.H;
  const long long total = per_dir * a.ndir;
  const int G = a.ndir * 4 * a.H;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int d = (int)(i / per_dir);
    const long long r = i - d * per_dir;
    const int row = (int)(r / a.H), j = (int)(r - (long long)row * a.H);
    const int t = (d & 1) ? a.T - 1 - a.s : a.s;
    const bool valid = a.seq_len == nullptr || t < a.seq_len[row];
    const float h_prev = a.s == 0 ? 0.f : a.h[i];
    const float c_prev = a.s == 0 ? 0.f : a.c_hist[(long long)a.s * total + i];
    float* g = a.gates + ((long long)t * a.S + row) * G + (long long)d * 4 * a.H + j;
    float h_new = h_prev, c_new = c_prev;
    if (valid) {
      float pi = g[0], pf = g[a.H], pg = g[2 * a.H], po = g[3 * a.H];
      if (a.s > 0) {
        const float* rc = a.rec + ((long long)d * a.S + row) * 4 * a.H + j;
        pi += rc[0]; pf += rc[a.H]; pg += rc[2 * a.H]; po += rc[3 * a.H];
      }
      const float gi = sigmoid_exact(pi), gf = sigmoid_exact(pf), gg = tanhf(pg), go = sigmoid_exact(po);
      c_new = gf * c_prev + gi * gg;
      h_new = go * tanhf(c_new);
      g[0] = gi; g[a.H] = gf; g[2 * a.H] = gg; g[3 * a.H] = go;
    }
    a.c_hist[(long long)(a.s + 1) * total + i] = c_new;
    a.h[i] = h_new;
    a.hprev_t[((long long)d * a.T + t) * per_dir + r] = h_prev;
    store_planes(static_cast<__nv_bfloat16*>(a.h_planes) + (long long)d * 3 * per_dir + r, per_dir, h_new);
    if (a.seq_out != nullptr)
      a.seq_out[((long long)row * a.T + t) * a.seq_out_ld + (long long)d * a.H + j] = valid ? h_new : 0.f;
    if (a.s == a.T - 1 && a.h_last != nullptr) a.h_last[(long long)row * a.h_last_ld + (long long)d * a.H + j] = h_new;
  }
}

// one backward step: the incoming dh of this step is dh[i] (the carry written by the previous call, plus the recurrent
// product the caller accumulated into it) + dh_seq at this position; writes the pre-activation gate gradients over the
// activated gates (fp32) and as bf16 planes (operand of dh_{t-1} = dgates W_hh), and re-initialises dh[i] with the part of
// the state gradient that does NOT flow through the gates (all of it at a padded position, none otherwise).
__global__ void __launch_bounds__(256)
lstm32_cell_bwd_kernel(dvgr_lstm32_args a) {
  const long long per_dir = (long long)a.S * a.H;
  const long long total = per_dir * a.ndir;
  const int G = a.ndir * 4 * a.H;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int d = (int)(i / per_dir);
    const long long r = i - d * per_dir;
    const int row = (int)(r / a.H), j = (int)(r - (long long)row * a.H);
    const int t = (d & 1) ? a.T - 1 - a.s : a.s;
    const bool valid = a.seq_len == nullptr || t < a.seq_len[row];
    float dh;
    if (a.s == a.T - 1) {
      dh = a.dh_last != nullptr ? a.dh_last[(long long)row * a.dh_last_ld + (long long)d * a.H + j] : 0.f;
    } else {
      dh = a.dh[i];
    }
    float dc = a.s == a.T - 1 ? 0.f : a.dc[i];
    float* g = a.gates + ((long long)t * a.S + row) * G + (long long)d * 4 * a.H + j;
    __nv_bfloat16* gp = static_cast<__nv_bfloat16*>(a.dgate_planes) + ((long long)d * 3 * a.S + row) * 4 * a.H + j;
    const long long plane = (long long)a.S * 4 * a.H;
    float di = 0.f, df = 0.f, dg = 0.f, dov = 0.f;
    if (valid) {
      if (a.dh_seq != nullptr) dh += a.dh_seq[((long long)row * a.T + t) * a.dh_seq_ld + (long long)d * a.H + j];
      const float gi = g[0], gf = g[a.H], gg = g[2 * a.H], go = g[3 * a.H];
      const float c_prev = a.s == 0 ? 0.f : a.c_hist[(long long)a.s * total + i];
      const float tc = tanhf(a.c_hist[(long long)(a.s + 1) * total + i]);
      const float dct = dc + dh * go * (1.f - tc * tc);
      dov = dh * tc * go * (1.f - go);
      di = dct * gg * gi * (1.f - gi);
      dg = dct * gi * (1.f - gg * gg);
      df = dct * c_prev * gf * (1.f - gf);
      dc = dct * gf;
      dh = 0.f;
    }
    float* go_ = a.dgates != nullptr ? a.dgates + (g - a.gates) : g;
    go_[0] = di; go_[a.H] = df; go_[2 * a.H] = dg; go_[3 * a.H] = dov;
    store_planes(gp, plane, di);
    store_planes(gp + a.H, plane, df);
    store_planes(gp + 2 * a.H, plane, dg);
    store_planes(gp + 3 * a.H, plane, dov);
    a.dc[i] = dc;
    a.dh[i] = dh;
  }
}

int check(const dvgr_lstm32_args* a, bool bwd) {
  if (!a) return set_error("lstm32: null args");
  if (a->S <= 0 || a->H <= 0 || a->T <= 0 || a->ndir <= 0 || a->s < 0 || a->s >= a->T)
    return set_error("lstm32: bad shape S=%d H=%d T=%d ndir=%d s=%d", a->S, a->H, a->T, a->ndir, a->s);
  if (!a->gates || !a->c_hist) return set_error("lstm32: null gates / c_hist");
  if (!bwd && (!a->h || !a->h_planes || !a->hprev_t || (a->s > 0 && !a->rec))) return set_error("lstm32 fwd: null state buffer");
  if (bwd && (!a->dh || !a->dc || !a->dgate_planes)) return set_error("lstm32 bwd: null gradient buffer");
  return 0;
}

int grid(const dvgr_lstm32_args* a) {
  const long long total = (long long)a->S * a->H * a->ndir;
  const long long blocks = (total + 255) / 256;
  return (int)(blocks < 148 * 8 ? blocks : 148 * 8);
}

}  // namespace
}  // namespace dvgr

extern "C" int dvgr_lstm32_cell_fwd(const dvgr_lstm32_args* a, void* stream) {
  if (int rc = dvgr::check(a, false)) return rc;
  dvgr::lstm32_cell_fwd_kernel<<<dvgr::grid(a), 256, 0, static_cast<cudaStream_t>(stream)>>>(*a);
  DVGR_CHECK_LAUNCH("lstm32_cell_fwd");
  return 0;
}

extern "C" int dvgr_lstm32_cell_bwd(const dvgr_lstm32_args* a, void* stream) {
  if (int rc = dvgr::check(a, true)) return rc;
  dvgr::lstm32_cell_bwd_kernel<<<dvgr::grid(a), 256, 0, static_cast<cudaStream_t>(stream)>>>(*a);
  DVGR_CHECK_LAUNCH("lstm32_cell_bwd");
  return 0;
}
