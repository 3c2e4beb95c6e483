// Internal glue shared by the .cu translation units of libdualvgr_b200.so.
#pragma once
#include <cuda_runtime.h>

#include "../../include/dualvgr_b200.h"
#include "gemm.cuh"

namespace dvgr {

// printf-style thread-local error message; returns a non-zero status for `return set_error(...)`.
int set_error(const char* fmt, ...);
void count_launch(int n = 1);
const unsigned long long* seed_offset_ptr();
void g_seed_off_set(const unsigned long long* p);

int gemm_dispatch(const dvgr_operand& A, const dvgr_operand& B, GemmParams p, int bn, int max_ctas, cudaStream_t stream);
int lstm_bwd_first(const GemmParams& p, const void* dh_last, long long dh_ld, cudaStream_t stream);
int gemm_grouped_wgrad(const dvgr_wgrad_problem* probs, int n, cudaStream_t stream);
int lstm_seq_bwd_launch(const dvgr_operand& G, const dvgr_operand& Whh, GemmParams p, int* sync, int max_ctas, cudaStream_t stream);
int lstm_seq_fwd_launch(const dvgr_operand& X, const dvgr_operand& Wih, const dvgr_operand& Hh, const dvgr_operand& Whh,
                        GemmParams p, int K1, const float* bias, int* sync, int max_ctas, cudaStream_t stream);
int gemm_ref(const void* A, long long a_rs, long long a_ks, const void* B, long long b_rs, long long b_ks, float* C,
             long long ldc, int M, int N, int K, cudaStream_t stream);

#define DVGR_CHECK_LAUNCH(name)                                                            \
  do {                                                                                     \
    cudaError_t e__ = cudaGetLastError();                                                  \
    if (e__ != cudaSuccess) return ::dvgr::set_error("%s launch failed: %s", name, cudaGetErrorString(e__)); \
    ::dvgr::count_launch();                                                                \
  } while (0)

}  // namespace dvgr
