// Whole-sequence softmax attention for the DiT blocks (reference
// models/dit.py:303-307: F.scaled_dot_product_attention on (B, H, L, 64) 16-bit
// tensors, scale 1/sqrt(64), causal only for parameterization=ar).
//
// At L <= 256, d = 64 one (batch, head) problem is 48 KB of q/k/v and 4 MFLOP:
// HBM-bound (65 FLOP/B), so v1 keeps the math on warp-level mma.sync m16n8k16
// with fp32 accumulation and spends the effort on one coalesced pass over qkv:
// cp.async into XOR-swizzled shared memory, ldmatrix fragments, online softmax
// over 64-key blocks, output staged through shared memory for 16-byte stores.
// One CTA per (batch, head); warp w owns query rows 16w..16w+15.
#include "ddg_common.cuh"

namespace ddg {
namespace {

template <typename T> struct Mma;
template <> struct Mma<__half> {
  static __device__ __forceinline__ void run(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
  static __device__ __forceinline__ uint32_t pack(float x, float y) {
    __half2 h = __floats2half2_rn(x, y);
    return *reinterpret_cast<uint32_t*>(&h);
  }
};
template <> struct Mma<__nv_bfloat16> {
  static __device__ __forceinline__ void run(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
  static __device__ __forceinline__ uint32_t pack(float x, float y) {
    __nv_bfloat162 h = __floats2bfloat162_rn(x, y);
    return *reinterpret_cast<uint32_t*>(&h);
  }
};

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

// byte offset of 16-byte chunk `ch` (0..7) of row `r` inside one [L x 64] tile
__device__ __forceinline__ uint32_t sw_off(int r, int ch) { return (uint32_t)(r * 128 + ((ch ^ (r & 7)) << 4)); }

constexpr float kLog2e = 1.4426950408889634f;

template <typename T, bool CAUSAL>
__global__ void __launch_bounds__(512)
attention_kernel(const T* __restrict__ qkv, T* __restrict__ out, int L, int H) {
  extern __shared__ __align__(128) uint8_t sm[];
  const uint32_t sm_base = static_cast<uint32_t>(__cvta_generic_to_shared(sm));
  const uint32_t tile_bytes = (uint32_t)L * 128u;
  const uint32_t sq = sm_base, sk = sm_base + tile_bytes, sv = sm_base + 2 * tile_bytes;

  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int64_t row_stride = (int64_t)3 * H * 64;
  const T* base = qkv + (int64_t)b * L * row_stride + (int64_t)h * 64;

  // ---- q, k, v of this (batch, head) -> shared memory
  const int chunks = 3 * L * 8;
  for (int c = threadIdx.x; c < chunks; c += blockDim.x) {
    const int which = c / (L * 8);
    const int r = (c / 8) % L;
    const int ch = c & 7;
    const T* src = base + (int64_t)r * row_stride + (int64_t)which * H * 64 + ch * 8;
    cp_async16(sm_base + which * tile_bytes + sw_off(r, ch), src);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int q0 = warp * 16;

  // ---- Q fragments (A operand), 4 k-steps over d
  uint32_t qf[4][4];
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    const int r = q0 + (lane & 7) + ((lane >> 3) & 1) * 8;
    const int ch = 2 * ks + (lane >> 4);
    ldsm_x4(qf[ks], sq + sw_off(r, ch));
  }

  float o[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
  float m_run[2] = {-INFINITY, -INFINITY};
  float l_run[2] = {0.f, 0.f};

  const int kv_end = CAUSAL ? (q0 + 16) : L;
  for (int kb0 = 0; kb0 < kv_end; kb0 += 64) {
    const int nvalid = min(64, kv_end - kb0);       // keys in this block (multiple of 16)
    float s[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
      if (nt * 8 < nvalid) {
#pragma unroll
        for (int kp = 0; kp < 2; ++kp) {
          uint32_t kf[4];
          const int r = kb0 + nt * 8 + (lane & 7);
          const int ch = 4 * kp + (lane >> 3);
          ldsm_x4(kf, sk + sw_off(r, ch));
          Mma<T>::run(s[nt], qf[2 * kp], kf[0], kf[1]);
          Mma<T>::run(s[nt], qf[2 * kp + 1], kf[2], kf[3]);
        }
      }
    }
    // ---- scale, mask, block row-max
    float bm[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int key = kb0 + nt * 8 + 2 * t + (e & 1);
        const int qrow = q0 + g + (e >> 1) * 8;
        float v = s[nt][e] * 0.125f;
        const bool dead = (nt * 8 >= nvalid) || (CAUSAL && key > qrow);
        v = dead ? -INFINITY : v;
        s[nt][e] = v;
        bm[e >> 1] = fmaxf(bm[e >> 1], v);
      }
    }
    float alpha[2], mnew[2];
#pragma unroll
    for (int hrow = 0; hrow < 2; ++hrow) {
      float v = bm[hrow];
      v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
      v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
      mnew[hrow] = fmaxf(m_run[hrow], v);
      alpha[hrow] = exp2f((m_run[hrow] - mnew[hrow]) * kLog2e);   // m_run = -inf -> 0
      m_run[hrow] = mnew[hrow];
      l_run[hrow] *= alpha[hrow];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      o[i][0] *= alpha[0]; o[i][1] *= alpha[0];
      o[i][2] *= alpha[1]; o[i][3] *= alpha[1];
    }
    // ---- p = exp(s - m), O += P V
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float p = exp2f((s[nt][e] - mnew[e >> 1]) * kLog2e);
        s[nt][e] = p;
        l_run[e >> 1] += p;
      }
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      if (kk * 16 < nvalid) {
        uint32_t pf[4];
        pf[0] = Mma<T>::pack(s[2 * kk][0], s[2 * kk][1]);
        pf[1] = Mma<T>::pack(s[2 * kk][2], s[2 * kk][3]);
        pf[2] = Mma<T>::pack(s[2 * kk + 1][0], s[2 * kk + 1][1]);
        pf[3] = Mma<T>::pack(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
        for (int dp = 0; dp < 4; ++dp) {
          uint32_t vf[4];
          const int r = kb0 + kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
          const int ch = 2 * dp + (lane >> 4);
          ldsm_x4_trans(vf, sv + sw_off(r, ch));
          Mma<T>::run(o[2 * dp], pf, vf[0], vf[1]);
          Mma<T>::run(o[2 * dp + 1], pf, vf[2], vf[3]);
        }
      }
    }
  }
  // ---- normalise, stage this warp's 16 x 64 tile over its own Q rows, store
#pragma unroll
  for (int hrow = 0; hrow < 2; ++hrow) {
    float v = l_run[hrow];
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    l_run[hrow] = 1.f / v;
  }
  __syncwarp();
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    // element (row g [+8], d = 8 nt + 2t, +1): chunk nt, byte 4t within the chunk
    const uint32_t w0 = Mma<T>::pack(o[nt][0] * l_run[0], o[nt][1] * l_run[0]);
    const uint32_t w1 = Mma<T>::pack(o[nt][2] * l_run[1], o[nt][3] * l_run[1]);
    *reinterpret_cast<uint32_t*>(sm + sw_off(q0 + g, nt) + 4 * t) = w0;
    *reinterpret_cast<uint32_t*>(sm + sw_off(q0 + g + 8, nt) + 4 * t) = w1;
  }
  __syncwarp();
  T* obase = out + ((int64_t)b * L + q0) * ((int64_t)H * 64) + (int64_t)h * 64;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = i * 32 + lane;     // 128 chunks: 16 rows x 8
    const int r = c >> 3, ch = c & 7;
    const uint4 v = *reinterpret_cast<const uint4*>(sm + sw_off(q0 + r, ch));
    *reinterpret_cast<uint4*>(obase + (int64_t)r * H * 64 + ch * 8) = v;
  }
}

template <typename T, bool CAUSAL>
int launch_attention(const void* qkv, void* out, int batch, int L, int H, cudaStream_t s) {
  const int smem = 3 * L * 128;
  static bool attr_done[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (smem > 48 * 1024 && dev >= 0 && dev < 64 && !attr_done[dev]) {
    DDG_CUDA(cudaFuncSetAttribute(attention_kernel<T, CAUSAL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  3 * 256 * 128));
    attr_done[dev] = true;
  }
  attention_kernel<T, CAUSAL><<<batch * H, (L / 16) * 32, smem, s>>>(
    reinterpret_cast<const T*>(qkv), reinterpret_cast<T*>(out), L, H);
  return check_cuda(cudaGetLastError(), "attention launch");
}

}  // namespace
}  // namespace ddg

extern "C" int ddg_attention(const void* qkv, void* out, int32_t batch, int32_t seq_len, int32_t n_heads,
                             int32_t head_dim, int32_t causal, int32_t dtype, ddg_stream_t stream) {
  using namespace ddg;
  int rc = ddg_device_check();
  if (rc) return rc;
  DDG_REQUIRE(qkv && out && batch >= 0 && n_heads > 0, DDG_ERR_INVALID, "attention: bad args");
  DDG_REQUIRE(head_dim == 64, DDG_ERR_UNSUPPORTED, "attention: head_dim must be 64 (got %d)", head_dim);
  DDG_REQUIRE(seq_len >= 16 && seq_len <= 256 && seq_len % 16 == 0, DDG_ERR_UNSUPPORTED,
              "attention: seq_len must be a multiple of 16 in [16, 256] (got %d)", seq_len);
  DDG_REQUIRE(reinterpret_cast<uintptr_t>(qkv) % 16 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0,
              DDG_ERR_INVALID, "attention: qkv/out must be 16-byte aligned");
  if (batch == 0) return DDG_OK;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == DDG_F16)
    return causal ? launch_attention<__half, true>(qkv, out, batch, seq_len, n_heads, s)
                  : launch_attention<__half, false>(qkv, out, batch, seq_len, n_heads, s);
  if (dtype == DDG_BF16)
    return causal ? launch_attention<__nv_bfloat16, true>(qkv, out, batch, seq_len, n_heads, s)
                  : launch_attention<__nv_bfloat16, false>(qkv, out, batch, seq_len, n_heads, s);
  set_error("attention: dtype must be f16 or bf16");
  return DDG_ERR_INVALID;
}
