// Attention input-gradient for the classifier backward of first-order D-CBG
// (reference diffusion.py:1532-1548 back-propagates through models/dit.py:284-307:
// rotary -> softmax(q k^T / 8) v).  Sequences of up to 128 tokens, head_dim 64.
//
// One CTA per (sequence, head), one warp per 16 rows.  q, k (post-rotary), v and dO
// of the head sit in shared memory as 128B-swizzled [Lp x 64] tiles; all five
// products run on mma.sync.m16n8k16 with fp32 accumulators, nothing L x L ever
// leaves the registers:
//   pass A (warp = 16 query rows): S = q k^T, dP = dO v^T -> row max / sum / D,
//          dS = P o (dP - D) / 8, dq = dS k;  row stats go to shared memory.
//   pass B (warp = 16 key rows):   S^T = k q^T, dP^T = v dO^T, P^T and dS^T from the
//          saved row stats, dv = P^T dO, dk = dS^T q.
// P and dS are rounded to the 16-bit type for the second product of each pass, as the
// reference's own (flash-attention, autocast) backward does.  dq / dk go back through
// the rotary (the transpose of the rotation), so g_qkv is the gradient of the QKV
// GEMM output.  Each warp stages its 16 x 64 result tiles in shared memory and stores
// whole 128-byte rows.
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
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// byte offset of 16-byte chunk `ch` (0..7) of row `r` inside one [Lp x 64] tile
__device__ __forceinline__ uint32_t sw_off(int r, int ch) { return (uint32_t)(r * 128 + ((ch ^ (r & 7)) << 4)); }

constexpr float kScale = 0.125f;                       // 1 / sqrt(64)
constexpr float kScale2 = 0.125f * 1.4426950408889634f;

// X (16 x 64, A fragments xa) times Y^T for the row tiles of Y: acc[nt] (16 x 8) for nt < ntl.
template <typename T, int NT>
__device__ __forceinline__ void rows_times_rows_t(float (&acc)[NT][4], const uint32_t (&xa)[4][4], uint32_t sy,
                                                  const uint32_t (&off_b)[2], int ntl) {
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
    if (nt < ntl) {
#pragma unroll
      for (int kp = 0; kp < 2; ++kp) {
        uint32_t f[4];
        ldsm_x4(f, sy + (uint32_t)(nt * 8) * 128u + off_b[kp]);
        Mma<T>::run(acc[nt], xa[2 * kp], f[0], f[1]);
        Mma<T>::run(acc[nt], xa[2 * kp + 1], f[2], f[3]);
      }
    }
  }
}

// acc (16 x 64) = W (16 x Lp, packed from w[nt][4]) times Y (Lp x 64 tile at sy).
template <typename T, int NT>
__device__ __forceinline__ void weights_times_rows(float (&acc)[8][4], const float (&w)[NT][4], uint32_t sy,
                                                   const uint32_t (&off_t)[4], int ntl) {
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
#pragma unroll
  for (int kk = 0; kk < NT / 2; ++kk) {
    if (2 * kk < ntl) {
      uint32_t a[4];
      a[0] = Mma<T>::pack(w[2 * kk][0], w[2 * kk][1]);
      a[1] = Mma<T>::pack(w[2 * kk][2], w[2 * kk][3]);
      a[2] = Mma<T>::pack(w[2 * kk + 1][0], w[2 * kk + 1][1]);
      a[3] = Mma<T>::pack(w[2 * kk + 1][2], w[2 * kk + 1][3]);
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {
        uint32_t f[4];
        ldsm_x4_trans(f, sy + (uint32_t)(kk * 16) * 128u + off_t[dp]);
        Mma<T>::run(acc[2 * dp], a, f[0], f[1]);
        Mma<T>::run(acc[2 * dp + 1], a, f[2], f[3]);
      }
    }
  }
}

// Gradient w.r.t. the rotary's input from the gradient w.r.t. its output, in place on a
// 16 x 64 accumulator tile: dims d < 32 pair with d + 32 (n-tiles nt and nt + 4 of the
// same thread).  x1' = x1 cos - x2 sin, x2' = x2 cos + x1 sin  (models/dit.py:60-75).
template <typename T>
__device__ __forceinline__ void unrotate_tile(float (&acc)[8][4], const T* __restrict__ rope_cos,
                                              const T* __restrict__ rope_sin, int row0, int L, int g, int t) {
#pragma unroll
  for (int hrow = 0; hrow < 2; ++hrow) {
    const int pos = min(row0 + g + 8 * hrow, L - 1);
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int idx = pos * 32 + nt * 8 + 2 * t + e;
        const float c = Cvt<T>::to_f32(rope_cos[idx]), s = Cvt<T>::to_f32(rope_sin[idx]);
        const float g1 = acc[nt][2 * hrow + e], g2 = acc[nt + 4][2 * hrow + e];
        acc[nt][2 * hrow + e] = g1 * c + g2 * s;
        acc[nt + 4][2 * hrow + e] = g2 * c - g1 * s;
      }
    }
  }
}

// Stage a warp's 16 x 64 tile (rounded to T) and store rows row0.. (< L) at dst (+ row * ld).
template <typename T>
__device__ __forceinline__ void store_tile(const float (&acc)[8][4], uint8_t* stage, T* dst, int64_t ld, int row0,
                                           int L, int lane) {
  const int g = lane >> 2, t = lane & 3;
  __syncwarp();
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    *reinterpret_cast<uint32_t*>(stage + sw_off(g, nt) + 4 * t) = Mma<T>::pack(acc[nt][0], acc[nt][1]);
    *reinterpret_cast<uint32_t*>(stage + sw_off(g + 8, nt) + 4 * t) = Mma<T>::pack(acc[nt][2], acc[nt][3]);
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = i * 32 + lane;               // 128 chunks: 16 rows x 8
    const int r = c >> 3, ch = c & 7;
    if (row0 + r < L)
      *reinterpret_cast<uint4*>(dst + (int64_t)(row0 + r) * ld + ch * 8) =
        *reinterpret_cast<const uint4*>(stage + sw_off(r, ch));
  }
}

template <typename T, int NT>
__global__ void __launch_bounds__(NT * 16)
attention_bwd_mma_kernel(const T* __restrict__ qkv, const T* __restrict__ g_o, const T* __restrict__ rope_cos,
                         const T* __restrict__ rope_sin, int L, int H, int causal, T* __restrict__ g_qkv) {
  extern __shared__ __align__(1024) uint8_t sm_gen[];
  const int Lp = (L + 15) & ~15;
  const int ntl = Lp >> 3;                       // 8-row tiles in use (<= NT)
  const uint32_t tile_bytes = (uint32_t)Lp * 128u;
  const uint32_t sm_base = (uint32_t)__cvta_generic_to_shared(sm_gen);
  const uint32_t sq = sm_base, sk = sq + tile_bytes, sv = sk + tile_bytes, sdo = sv + tile_bytes;
  float* st_m = reinterpret_cast<float*>(sm_gen + 4 * tile_bytes);      // row max * scale * log2e
  float* st_li = st_m + Lp;                                             // 1 / row sum
  float* st_D = st_li + Lp;                                             // sum_j P_ij dP_ij
  uint8_t* stage_all = reinterpret_cast<uint8_t*>(st_D + Lp);

  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3, l7 = lane & 7;
  const int64_t row_stride = (int64_t)3 * H * 64, o_stride = (int64_t)H * 64;
  const T* base = qkv + (int64_t)b * L * row_stride + (int64_t)h * 64;
  const T* gbase = g_o + (int64_t)b * L * o_stride + (int64_t)h * 64;

  // ---- fill the four tiles (rows >= L: zeros)
  for (int i = tid; i < Lp * 32; i += blockDim.x) {
    const int which = i / (Lp * 8), rc = i % (Lp * 8), r = rc >> 3, ch = rc & 7;
    const uint32_t dst = sm_base + (uint32_t)which * tile_bytes + sw_off(r, ch);
    if (r < L) {
      const T* src = which < 3 ? base + (int64_t)r * row_stride + (int64_t)which * H * 64 + ch * 8
                               : gbase + (int64_t)r * o_stride + ch * 8;
      cp_async16(dst, src);
    } else {
      *reinterpret_cast<uint4*>(sm_gen + (dst - sm_base)) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  // per-lane ldmatrix offsets (row & 7 == lane & 7 for every tile addressed)
  const uint32_t row_a = (uint32_t)(l7 + ((lane >> 3) & 1) * 8) * 128u;
  uint32_t off_a[4], off_b[2];
#pragma unroll
  for (int i = 0; i < 4; ++i) off_a[i] = row_a + (uint32_t)(((2 * i + (lane >> 4)) ^ l7) << 4);
#pragma unroll
  for (int i = 0; i < 2; ++i) off_b[i] = (uint32_t)l7 * 128u + (uint32_t)(((4 * i + (lane >> 3)) ^ l7) << 4);

  const int r0 = warp * 16;                      // this warp's rows (queries in A, keys in B)
  uint8_t* stage = stage_all + warp * 2048;
  T* out_base = g_qkv + (int64_t)b * L * row_stride + (int64_t)h * 64;
  if (r0 < Lp) {
    // ================= pass A: query rows r0 .. r0 + 15
    uint32_t qa[4][4], da[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      ldsm_x4(qa[ks], sq + (uint32_t)r0 * 128u + off_a[ks]);
      ldsm_x4(da[ks], sdo + (uint32_t)r0 * 128u + off_a[ks]);
    }
    float s[NT][4], dp[NT][4];
    rows_times_rows_t<T, NT>(s, qa, sk, off_b, ntl);
    rows_times_rows_t<T, NT>(dp, da, sv, off_b, ntl);
    float m[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int key = nt * 8 + 2 * t + (e & 1), qrow = r0 + g + (e >> 1) * 8;
        if (key >= L || (causal && key > qrow)) s[nt][e] = -INFINITY;
        m[e >> 1] = fmaxf(m[e >> 1], s[nt][e]);
      }
    float msc[2], li[2], D[2];
#pragma unroll
    for (int hrow = 0; hrow < 2; ++hrow) {
      float v = m[hrow];
      v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
      v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
      msc[hrow] = v * kScale2;
    }
    float l[2] = {0.f, 0.f};
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        s[nt][e] = ex2_approx(fmaf(s[nt][e], kScale2, -msc[e >> 1]));      // -inf -> 0
        l[e >> 1] += s[nt][e];
      }
#pragma unroll
    for (int hrow = 0; hrow < 2; ++hrow) {
      float v = l[hrow];
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      li[hrow] = 1.f / v;
    }
    float dsum[2] = {0.f, 0.f};
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        s[nt][e] *= li[e >> 1];                                            // P
        dsum[e >> 1] = fmaf(s[nt][e], dp[nt][e], dsum[e >> 1]);
      }
#pragma unroll
    for (int hrow = 0; hrow < 2; ++hrow) {
      float v = dsum[hrow];
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      D[hrow] = v;
      if (t == 0) {
        const int row = r0 + g + 8 * hrow;
        st_m[row] = msc[hrow];
        st_li[row] = li[hrow];
        st_D[row] = v;
      }
    }
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) s[nt][e] = s[nt][e] * (dp[nt][e] - D[e >> 1]) * kScale;   // dS
    float acc[8][4];
    weights_times_rows<T, NT>(acc, s, sk, off_a, ntl);
    unrotate_tile<T>(acc, rope_cos, rope_sin, r0, L, g, t);
    store_tile<T>(acc, stage, out_base, row_stride, r0, L, lane);
  }
  __syncthreads();
  if (r0 < Lp) {
    // ================= pass B: key rows r0 .. r0 + 15
    uint32_t ka[4][4], va[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      ldsm_x4(ka[ks], sk + (uint32_t)r0 * 128u + off_a[ks]);
      ldsm_x4(va[ks], sv + (uint32_t)r0 * 128u + off_a[ks]);
    }
    float p[NT][4], ds[NT][4];
    rows_times_rows_t<T, NT>(p, ka, sq, off_b, ntl);       // S^T: rows = keys, cols = queries
    rows_times_rows_t<T, NT>(ds, va, sdo, off_b, ntl);     // dP^T
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      if (nt < ntl) {
        const int q0 = nt * 8 + 2 * t;
        const float2 mq = *reinterpret_cast<const float2*>(st_m + q0);
        const float2 lq = *reinterpret_cast<const float2*>(st_li + q0);
        const float2 Dq = *reinterpret_cast<const float2*>(st_D + q0);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int query = q0 + (e & 1), key = r0 + g + (e >> 1) * 8;
          const bool dead = query >= L || (causal && key > query);
          const float pe = dead ? 0.f
                                : ex2_approx(fmaf(p[nt][e], kScale2, -((e & 1) ? mq.y : mq.x))) * ((e & 1) ? lq.y : lq.x);
          p[nt][e] = pe;
          ds[nt][e] = pe * (ds[nt][e] - ((e & 1) ? Dq.y : Dq.x)) * kScale;
        }
      }
    }
    float acc[8][4];
    weights_times_rows<T, NT>(acc, ds, sq, off_a, ntl);    // dk = dS^T q
    unrotate_tile<T>(acc, rope_cos, rope_sin, r0, L, g, t);
    store_tile<T>(acc, stage, out_base + (int64_t)H * 64, row_stride, r0, L, lane);
    weights_times_rows<T, NT>(acc, p, sdo, off_a, ntl);    // dv = P^T dO
    store_tile<T>(acc, stage, out_base + (int64_t)2 * H * 64, row_stride, r0, L, lane);
  }
}

template <typename T, int NT>
int launch_nt(const T* qkv, const T* g_o, const T* cosb, const T* sinb, int batch, int L, int H, int causal, T* g_qkv,
              cudaStream_t s) {
  const int Lp = (L + 15) & ~15;
  const size_t smem = (size_t)4 * Lp * 128 + (size_t)3 * Lp * 4 + (size_t)(NT / 2) * 2048;
  if (smem > 48 * 1024)
    DDG_CUDA(cudaFuncSetAttribute(attention_bwd_mma_kernel<T, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
  attention_bwd_mma_kernel<T, NT><<<(unsigned)(batch * H), NT * 16, smem, s>>>(qkv, g_o, cosb, sinb, L, H, causal,
                                                                              g_qkv);
  return check_cuda(cudaGetLastError(), "attention_backward (mma) launch");
}

template <typename T>
int launch_t(const void* qkv, const void* g_o, const void* cosb, const void* sinb, int batch, int L, int H, int causal,
             void* g_qkv, cudaStream_t s) {
  const T* a = reinterpret_cast<const T*>(qkv);
  const T* b = reinterpret_cast<const T*>(g_o);
  const T* c = reinterpret_cast<const T*>(cosb);
  const T* d = reinterpret_cast<const T*>(sinb);
  T* o = reinterpret_cast<T*>(g_qkv);
  if (L <= 32) return launch_nt<T, 4>(a, b, c, d, batch, L, H, causal, o, s);
  if (L <= 64) return launch_nt<T, 8>(a, b, c, d, batch, L, H, causal, o, s);
  return launch_nt<T, 16>(a, b, c, d, batch, L, H, causal, o, s);
}

}  // namespace

// seq_len <= 128, head_dim 64; dtype DDG_F16 / DDG_BF16 (checked by the caller).
int launch_attention_bwd_mma(const void* qkv, const void* g_o, const void* rope_cos, const void* rope_sin, int batch,
                             int seq_len, int n_heads, int causal, int dtype, void* g_qkv, cudaStream_t s) {
  if (dtype == DDG_F16)
    return launch_t<__half>(qkv, g_o, rope_cos, rope_sin, batch, seq_len, n_heads, causal, g_qkv, s);
  return launch_t<__nv_bfloat16>(qkv, g_o, rope_cos, rope_sin, batch, seq_len, n_heads, causal, g_qkv, s);
}

}  // namespace ddg
