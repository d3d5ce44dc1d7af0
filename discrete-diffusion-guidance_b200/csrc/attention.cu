// Whole-sequence softmax attention for the DiT blocks (reference
// models/dit.py:303-307: F.scaled_dot_product_attention on (B, H, L, 64) 16-bit
// tensors, scale 1/sqrt(64), causal only for parameterization=ar).
//
// At L <= 256, d = 64 one (batch, head) problem is 48 KB of q/k/v and 4 MFLOP:
// HBM-bound (65 FLOP/B), so the math stays on warp-level mma.sync m16n8k16 with
// fp32 accumulation and the effort goes into keeping HBM busy: persistent CTAs
// walk the (batch, head) problems; the q, k, v tiles of the NEXT problem are
// fetched by TMA (three 128B-swizzled boxes, one mbarrier) into the second half
// of shared memory while the warps work on the current one.  ldmatrix fragments,
// online softmax over 64-key blocks, output staged through shared memory for
// 16-byte stores.  Warp w owns query rows 16w..16w+15.
#include <cuda.h>

#include "ddg_common.cuh"

namespace ddg {
int make_tma_rows64(void* map, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows, int dtype);  // gemm_tcgen05.cu
namespace {

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
    "{\n\t"
    ".reg .pred P1;\n\t"
    "ATT_WAIT:\n\t"
    "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
    "@P1 bra.uni ATT_DONE;\n\t"
    "bra.uni ATT_WAIT;\n\t"
    "ATT_DONE:\n\t"
    "}" ::"r"(bar), "r"(parity)
    : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* m, uint32_t bar, uint32_t dst, int c0, int c1) {
  asm volatile(
    "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
    ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
    : "memory");
}

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

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// FULL: non-causal and L a multiple of 64 -> no masking code at all.
// Shared memory: two buffers of [q | k | v] tiles (L x 128 B each, 1024-byte aligned for
// the 128B swizzle), then two mbarriers.
template <typename T, bool CAUSAL, bool FULL>
__global__ void __launch_bounds__(512)
attention_kernel(const __grid_constant__ CUtensorMap tma_qkv, T* __restrict__ out, int L, int H, int n_items) {
  extern __shared__ uint8_t sm_raw[];
  const uint32_t sm_base = (static_cast<uint32_t>(__cvta_generic_to_shared(sm_raw)) + 1023u) & ~1023u;
  uint8_t* sm_gen = sm_raw + (sm_base - static_cast<uint32_t>(__cvta_generic_to_shared(sm_raw)));
  const uint32_t tile_bytes = (uint32_t)L * 128u;
  const uint32_t buf_bytes = (3u * tile_bytes + 1023u) & ~1023u;
  const uint32_t bar0 = sm_base + 2u * buf_bytes;

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_qkv)) : "memory");
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // q, k, v tiles of problem `item` -> buffer `b` (one thread; completion on the buffer's mbarrier)
  auto issue_load = [&](int item, int b) {
    const int bb = item / H, hh = item % H;
    const uint32_t bar = bar0 + 8u * (uint32_t)b;
    const uint32_t dst = sm_base + (uint32_t)b * buf_bytes;
    mbar_expect_tx(bar, 3u * tile_bytes);
#pragma unroll
    for (int which = 0; which < 3; ++which)
      tma_load_2d(&tma_qkv, bar, dst + (uint32_t)which * tile_bytes, (which * H + hh) * 64, bb * L);
  };
  if (threadIdx.x == 0 && (int)blockIdx.x < n_items) issue_load(blockIdx.x, 0);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int q0 = warp * 16;
  const int l7 = lane & 7;

  // per-lane ldmatrix offsets: row & 7 == lane & 7 for every tile we address
  // Q: rows q0 + l7 + 8*((lane>>3)&1), chunk 2ks + (lane>>4)
  // K: rows kb + l7,                    chunk 4kp + (lane>>3)
  // V: rows kb + l7 + 8*((lane>>3)&1), chunk 2dp + (lane>>4)
  const uint32_t row_qv = (uint32_t)(l7 + ((lane >> 3) & 1) * 8) * 128u;
  uint32_t off_q[4], off_k[2], off_v[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    off_q[i] = row_qv + (uint32_t)(((2 * i + (lane >> 4)) ^ l7) << 4);
    off_v[i] = off_q[i];
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) off_k[i] = (uint32_t)l7 * 128u + (uint32_t)(((4 * i + (lane >> 3)) ^ l7) << 4);

  int it = 0;
  for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
    const int cur = it & 1;
    // the other buffer was last read in iteration it - 1 (barrier at its end): refill it now
    if (threadIdx.x == 0 && item + (int)gridDim.x < n_items) issue_load(item + gridDim.x, cur ^ 1);
    const int b = item / H, h = item % H;
    const uint32_t sq = sm_base + (uint32_t)cur * buf_bytes, sk = sq + tile_bytes, sv = sq + 2 * tile_bytes;
    uint8_t* smb = sm_gen + (uint32_t)cur * buf_bytes;
    mbar_wait(bar0 + 8u * (uint32_t)cur, (uint32_t)((it >> 1) & 1));

    uint32_t qf[4][4];
  #pragma unroll
    for (int ks = 0; ks < 4; ++ks) ldsm_x4(qf[ks], sq + (uint32_t)q0 * 128u + off_q[ks]);

    float o[8][4];
  #pragma unroll
    for (int i = 0; i < 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
    float m_run[2] = {-INFINITY, -INFINITY};
    float l_run[2] = {0.f, 0.f};
    constexpr float kScale2 = 0.125f * kLog2e;       // softmax scale 1/sqrt(64), in log2 units

    const int kv_end = CAUSAL ? (q0 + 16) : L;
    for (int kb0 = 0; kb0 < kv_end; kb0 += 64) {
      const int nvalid = FULL ? 64 : min(64, kv_end - kb0);   // keys in this block (multiple of 16)
      float s[8][4];
      const uint32_t kbase = sk + (uint32_t)kb0 * 128u;
  #pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
        if (FULL || nt * 8 < nvalid) {
  #pragma unroll
          for (int kp = 0; kp < 2; ++kp) {
            uint32_t kf[4];
            ldsm_x4(kf, kbase + (uint32_t)(nt * 8) * 128u + off_k[kp]);
            Mma<T>::run(s[nt], qf[2 * kp], kf[0], kf[1]);
            Mma<T>::run(s[nt], qf[2 * kp + 1], kf[2], kf[3]);
          }
        }
      }
      // ---- mask (ragged / causal only) and block row-max of the raw scores
      float bm[2] = {-INFINITY, -INFINITY};
  #pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
  #pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (!FULL) {
            const int key = kb0 + nt * 8 + 2 * t + (e & 1);
            const int qrow = q0 + g + (e >> 1) * 8;
            const bool dead = (nt * 8 >= nvalid) || (CAUSAL && key > qrow);
            if (dead) s[nt][e] = -INFINITY;
          }
          bm[e >> 1] = fmaxf(bm[e >> 1], s[nt][e]);
        }
      }
      float msc[2];
  #pragma unroll
      for (int hrow = 0; hrow < 2; ++hrow) {
        float v = bm[hrow];
        v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
        v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
        const float mnew = fmaxf(m_run[hrow], v);
        const float alpha = ex2_approx((m_run[hrow] - mnew) * kScale2);   // m_run = -inf -> 0
        m_run[hrow] = mnew;
        msc[hrow] = mnew * kScale2;
        l_run[hrow] *= alpha;
  #pragma unroll
        for (int i = 0; i < 8; ++i) { o[i][2 * hrow] *= alpha; o[i][2 * hrow + 1] *= alpha; }
      }
      // ---- p = exp(scale (s - m)), O += P V
      const uint32_t vbase = sv + (uint32_t)kb0 * 128u;
  #pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        if (FULL || kk * 16 < nvalid) {
          uint32_t pf[4];
  #pragma unroll
          for (int half = 0; half < 2; ++half) {
            const int nt = 2 * kk + half;
            const float p0 = ex2_approx(fmaf(s[nt][0], kScale2, -msc[0]));
            const float p1 = ex2_approx(fmaf(s[nt][1], kScale2, -msc[0]));
            const float p2 = ex2_approx(fmaf(s[nt][2], kScale2, -msc[1]));
            const float p3 = ex2_approx(fmaf(s[nt][3], kScale2, -msc[1]));
            l_run[0] += p0 + p1;
            l_run[1] += p2 + p3;
            pf[2 * half] = Mma<T>::pack(p0, p1);
            pf[2 * half + 1] = Mma<T>::pack(p2, p3);
          }
  #pragma unroll
          for (int dp = 0; dp < 4; ++dp) {
            uint32_t vf[4];
            ldsm_x4_trans(vf, vbase + (uint32_t)(kk * 16) * 128u + off_v[dp]);
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
      *reinterpret_cast<uint32_t*>(smb + sw_off(q0 + g, nt) + 4 * t) = w0;
      *reinterpret_cast<uint32_t*>(smb + sw_off(q0 + g + 8, nt) + 4 * t) = w1;
    }
    __syncwarp();
    T* obase = out + ((int64_t)b * L + q0) * ((int64_t)H * 64) + (int64_t)h * 64;
  #pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = i * 32 + lane;     // 128 chunks: 16 rows x 8
      const int r = c >> 3, ch = c & 7;
      const uint4 v = *reinterpret_cast<const uint4*>(smb + sw_off(q0 + r, ch));
      *reinterpret_cast<uint4*>(obase + (int64_t)r * H * 64 + ch * 8) = v;
    }
    // the staged output (generic proxy) must be ordered before the TMA refill (async proxy) of this buffer
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
  }
}

template <typename T, bool CAUSAL, bool FULL>
int launch_attention_k(const void* qkv, void* out, int batch, int L, int H, int dtype, cudaStream_t s) {
  const int buf = (3 * L * 128 + 1023) / 1024 * 1024;
  const int smem = 2 * buf + 1024 /*align slack*/ + 64 /*barriers*/;
  static bool attr_done[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_done[dev]) {
    DDG_CUDA(cudaFuncSetAttribute(attention_kernel<T, CAUSAL, FULL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  2 * 3 * 256 * 128 + 1024 + 64));
    attr_done[dev] = true;
  }
  CUtensorMap map;
  const int64_t cols = (int64_t)3 * H * 64;
  int rc = make_tma_rows64(&map, qkv, (int64_t)batch * L, cols, cols, L, dtype);
  if (rc) return rc;
  const int n_items = batch * H;
  const int ctas_per_sm = smem <= 100 * 1024 ? 2 : 1;
  int grid = sm_count() * ctas_per_sm;
  if (grid > n_items) grid = n_items;
  attention_kernel<T, CAUSAL, FULL><<<grid, (L / 16) * 32, smem, s>>>(map, reinterpret_cast<T*>(out), L, H, n_items);
  return check_cuda(cudaGetLastError(), "attention launch");
}

template <typename T, bool CAUSAL>
int launch_attention(const void* qkv, void* out, int batch, int L, int H, int dtype, cudaStream_t s) {
  if (!CAUSAL && L % 64 == 0) return launch_attention_k<T, false, true>(qkv, out, batch, L, H, dtype, s);
  return launch_attention_k<T, CAUSAL, false>(qkv, out, batch, L, H, dtype, s);
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
    return causal ? launch_attention<__half, true>(qkv, out, batch, seq_len, n_heads, dtype, s)
                  : launch_attention<__half, false>(qkv, out, batch, seq_len, n_heads, dtype, s);
  if (dtype == DDG_BF16)
    return causal ? launch_attention<__nv_bfloat16, true>(qkv, out, batch, seq_len, n_heads, dtype, s)
                  : launch_attention<__nv_bfloat16, false>(qkv, out, batch, seq_len, n_heads, dtype, s);
  set_error("attention: dtype must be f16 or bf16");
  return DDG_ERR_INVALID;
}
