// Activation storage type of the fused (non-GEMM) kernels. qpm.cu, fused.cu and gat.cu are compiled TWICE:
//   default      act_t = bf16   -> the entry points declared in include/dualvgr_b200.h           (dvgr_xxx)
//   -DDVGR_F32   act_t = float  -> the same entry points with the suffix _f32                     (dvgr_xxx_f32)
// The _f32 build is the "fp32 mode" of the north star (1e-4 parity with the reference's fp32 path): every activation tensor
// (`void*` in the C ABI) is stored as float, GEMMs run as 3 x bf16 split products (dvgr_split3). Kernels are written against
// the helpers below; each variant lives in its own inner namespace so the two object files link into one library.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

#ifdef DVGR_F32
typedef float act_t;
#define DVGR_FN(name) name##_f32
#define DVGR_VNS v_f32
#else
typedef __nv_bfloat16 act_t;
#define DVGR_FN(name) name
#define DVGR_VNS v_bf16
#endif

namespace dvgr {
namespace act {

__device__ __forceinline__ void ld8(const __nv_bfloat16* p, float (&f)[8]) {
  const uint4 v = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float2 t = __bfloat1622float2(h[q]);
    f[2 * q] = t.x;
    f[2 * q + 1] = t.y;
  }
}
__device__ __forceinline__ void ld8(const float* p, float (&f)[8]) {
  const float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
__device__ __forceinline__ void st8(__nv_bfloat16* p, const float (&f)[8]) {
  uint4 o;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
  for (int q = 0; q < 4; ++q) h[q] = __floats2bfloat162_rn(f[2 * q], f[2 * q + 1]);
  *reinterpret_cast<uint4*>(p) = o;
}
__device__ __forceinline__ void st8(float* p, const float (&f)[8]) {
  reinterpret_cast<float4*>(p)[0] = make_float4(f[0], f[1], f[2], f[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(f[4], f[5], f[6], f[7]);
}
__device__ __forceinline__ float2 ld2(const __nv_bfloat16* p) {
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p));
}
__device__ __forceinline__ float2 ld2(const float* p) { return *reinterpret_cast<const float2*>(p); }
__device__ __forceinline__ void st2(__nv_bfloat16* p, float a, float b) {
  *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b);
}
__device__ __forceinline__ void st2(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }
__device__ __forceinline__ float ld1(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ float ld1(const float* p) { return *p; }
__device__ __forceinline__ void st1(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
__device__ __forceinline__ void st1(float* p, float v) { *p = v; }

}  // namespace act
}  // namespace dvgr
