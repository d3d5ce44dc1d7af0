// Shared device/host helpers for the ddg_b200 kernels (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ddg_b200.h"

namespace ddg {

// ---------------------------------------------------------------- errors
// The C ABI never throws: every entry point returns 0 or a negative code and
// leaves a message for ddg_last_error() (thread-local).
void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
int sm_count();

#define DDG_REQUIRE(cond, code, ...)        \
  do {                                      \
    if (!(cond)) {                          \
      ::ddg::set_error(__VA_ARGS__);        \
      return (code);                        \
    }                                       \
  } while (0)

#define DDG_CUDA(call)                                        \
  do {                                                        \
    int _rc = ::ddg::check_cuda((call), #call);               \
    if (_rc) return _rc;                                      \
  } while (0)

// ---------------------------------------------------------------- numerics
__device__ __forceinline__ float bf16_round(float x) {
  return __bfloat162float(__float2bfloat16_rn(x));
}
__device__ __forceinline__ float fp16_round(float x) {
  return __half2float(__float2half_rn(x));
}

template <typename T> struct Cvt;
template <> struct Cvt<float> {
  static __device__ __forceinline__ float to_f32(float v) { return v; }
  static __device__ __forceinline__ float from_f32(float v) { return v; }
  static __device__ __forceinline__ float round(float v) { return v; }
};
template <> struct Cvt<__half> {
  static __device__ __forceinline__ float to_f32(__half v) { return __half2float(v); }
  static __device__ __forceinline__ __half from_f32(float v) { return __float2half_rn(v); }
  static __device__ __forceinline__ float round(float v) { return fp16_round(v); }
};
template <> struct Cvt<__nv_bfloat16> {
  static __device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
  static __device__ __forceinline__ __nv_bfloat16 from_f32(float v) { return __float2bfloat16_rn(v); }
  static __device__ __forceinline__ float round(float v) { return bf16_round(v); }
};

// d/dv [0.5 v (1 + tanh u)],  u = sqrt(2/pi) (v + 0.044715 v^3)   (models/dit.py:243, the backward of
// nn.GELU(approximate='tanh')).  With s = sigmoid(2u) = (1 + tanh u) / 2:  1 - tanh^2 u = 4 s (1 - s), so
//   gelu'(v) = s (1 + v (2 du) (1 - s)),   2 du = 2 k0 (1 + 3 k1 v^2),   s = 1 / (1 + 2^(-2 u log2 e))
// -- one ex2 and one rcp (MUFU; abs. error ~1e-7, far below the 16-bit rounding of the product this
// feeds), exact limits for large |v|.  Every step is an explicitly rounded operation so that the
// packed f32x2 twin in the GEMM's GELU_BWD epilogue (gemm_tcgen05.cu) gives bit-identical results.
constexpr float kGeluK0 = 0.7978845608028654f;                 // sqrt(2 / pi)
constexpr float kGeluK1 = 0.044715f;
constexpr float kGeluExp = -2.f * 0.7978845608028654f * 1.4426950408889634f;   // -2 k0 log2 e
constexpr float kGeluD0 = 2.f * 0.7978845608028654f;           // 2 k0
constexpr float kGeluD1 = 6.f * 0.7978845608028654f * 0.044715f;   // 6 k0 k1
__device__ __forceinline__ float gelu_tanh_grad(float v) {
  const float v2 = __fmul_rn(v, v);
  const float arg = __fmul_rn(v, __fmaf_rn(kGeluK1, v2, 1.f));
  float e, s;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(__fmul_rn(arg, kGeluExp)));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(__fadd_rn(e, 1.f)));
  const float w = __fmul_rn(v, __fmaf_rn(kGeluD1, v2, kGeluD0));
  const float oms = __fmaf_rn(s, -1.f, 1.f);
  return __fmul_rn(s, __fmaf_rn(w, oms, 1.f));
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------- Philox
// Philox4x32-10 counter-based generator (Salmon et al., SC'11).  Stream
// layout used by the sampler (documented in DESIGN.md so the tests can
// regenerate it on the CPU):
//   key     = (seed_lo, seed_hi)
//   counter = (j / 4, row_lo, row_hi, step)      -> 4 words for j..j+3
//   U       = (word >> 8) * 2^-24   in [0, 1)
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  constexpr uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    const uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += W0;
    k.y += W1;
  }
  return c;
}
__device__ __forceinline__ float u01_from_bits(uint32_t w) {
  return (float)(w >> 8) * 5.9604644775390625e-08f;  // 2^-24
}

}  // namespace ddg
