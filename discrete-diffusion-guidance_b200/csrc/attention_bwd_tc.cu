// Attention input-gradient on tcgen05 / tensor memory for the classifier backward of
// first-order D-CBG (reference diffusion.py:1532-1548 back-propagates through
// models/dit.py:284-307: rotary -> softmax(q k^T / 8) v), non-causal, head_dim 64,
// sequences of 32 / 64 / 128 tokens.  Same item scheme as the forward kernel
// (attention.cu, namespace tc): one item = 128 consecutive token rows of one head --
// one 128-token sequence, or 128 / L whole short sequences whose cross-sequence score
// blocks are simply not read (P = dS = 0 there).
//
// One persistent CTA per SM, 160 threads: warp 4 drives TMA + the MMAs, warps 0-3 are
// "thread = row" (row r of the item = tensor-memory lane r).  Per item, all products are
// M = 128 tcgen05.mma with operands in 128B-swizzled shared memory, K-major or (for the
// 64-wide right-hand sides) MN-major exactly as in the forward kernel:
//   pass A (thread = query row)   S = q k^T, dP = dO v^T  -> row max / sum / D = sum_j P dP,
//                                 dS = P o (dP - D) / 8 (16-bit, K-major in smem) -> dq = dS k
//   pass B (thread = key row)     S^T = k q^T, dP^T = v dO^T -> P^T and dS^T from the saved row
//                                 statistics (16-bit, K-major in smem) -> dv = P^T dO, dk = dS^T q
// (forming the transposed scores with two more MMAs is cheaper than transposing P / dS
// through shared memory).  P and dS are rounded to the 16-bit type for the second product
// of each pass, as the reference's own (flash-attention, autocast) backward does.  dq / dk
// go back through the rotary (the transpose of the rotation), so g_qkv is the gradient of
// the QKV GEMM output.  q, k, v, dO of the next item arrive by TMA in the second buffer
// set while this one is processed.  Tensor memory: S 128 + dP 128 + dq 64 + dk 64 + dv 64
// columns.
#include <cuda.h>

#include <type_traits>

#include "ddg_common.cuh"

namespace ddg {
int make_tma_rows64(void* map, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows, int dtype);  // gemm_tcgen05.cu
namespace {
namespace bwdtc {

constexpr int ROWS = 128;
constexpr int TILE = ROWS * 128;                  // one 128 x 64 16-bit tile
// 2 x (q, k, v, dO) + dS (2) + P^T (2) + output staging (1) tiles, statistics, barriers, align slack
constexpr int SMEM_BYTES = 13 * TILE + 3 * ROWS * 4 + 128 + 1024;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kScale = 0.125f;                  // 1 / sqrt(64)
constexpr float kScale2 = 0.125f * kLog2e;

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
    "{\n\t"
    ".reg .pred P1;\n\t"
    "ATTB_WAIT:\n\t"
    "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
    "@P1 bra.uni ATTB_DONE;\n\t"
    "bra.uni ATTB_WAIT;\n\t"
    "ATTB_DONE:\n\t"
    "}" ::"r"(bar), "r"(parity)
    : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* m, uint32_t bar, uint32_t dst, int c0, int c1) {
  asm volatile(
    "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
    ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
    : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
    "{\n\t"
    ".reg .pred p;\n\t"
    "setp.ne.b32 p, %4, 0;\n\t"
    "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
    "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
    : "memory");
}
__device__ __forceinline__ void tmem_ld32_sync(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
    "tcgen05.wait::ld.sync.aligned;"
    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
      "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
      "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
      "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
    : "r"(taddr)
    : "memory");
}
// 128B-swizzled shared-memory matrix descriptor; SBO = 1024 B (8 rows of 128 B)
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {
  uint64_t d = (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(1024 >> 4) << 16;   // LBO (unused by the swizzled K-major / single-atom MN-major cases)
  d |= (uint64_t)(1024 >> 4) << 32;   // SBO
  d |= (uint64_t)1 << 46;             // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;             // SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: D = f32, A/B = T, M = 128, N, B optionally MN-major
template <typename T> __device__ __forceinline__ constexpr uint32_t idesc(int n, int b_mn_major) {
  const uint32_t fmt = std::is_same<T, __nv_bfloat16>::value ? 1u : 0u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(128 >> 4) << 24);
}
template <typename T> __device__ __forceinline__ uint32_t pack16(float x, float y);
template <> __device__ __forceinline__ uint32_t pack16<__half>(float x, float y) {
  __half2 h = __floats2half2_rn(x, y);
  return *reinterpret_cast<uint32_t*>(&h);
}
template <> __device__ __forceinline__ uint32_t pack16<__nv_bfloat16>(float x, float y) {
  __nv_bfloat162 h = __floats2bfloat162_rn(x, y);
  return *reinterpret_cast<uint32_t*>(&h);
}
template <typename T> __device__ __forceinline__ float2 unpack16(uint32_t w);
template <> __device__ __forceinline__ float2 unpack16<__half>(uint32_t w) {
  return __half22float2(*reinterpret_cast<__half2*>(&w));
}
template <> __device__ __forceinline__ float2 unpack16<__nv_bfloat16>(uint32_t w) {
  return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float4 ld_shared_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr)
               : "memory");
  return v;
}

// 32 columns (chunk c of the 128) of row `row` of a two-tile K-major operand: tile c / 2,
// 16-byte chunks 4 (c & 1) .. + 3, swizzled
__device__ __forceinline__ void store_chunk16(uint32_t base, int row, int c, const uint32_t (&w)[16]) {
  const uint32_t tile = base + (uint32_t)(c >> 1) * TILE + (uint32_t)row * 128u;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint32_t ch = (uint32_t)((4 * (c & 1) + q) ^ (row & 7));
    st_shared_v4(tile + (ch << 4), w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
  }
}

template <typename T, int LSEQ>
__global__ void __launch_bounds__(160, 1)
attention_bwd_tc_kernel(const __grid_constant__ CUtensorMap tma_qkv, const __grid_constant__ CUtensorMap tma_do,
                        const T* __restrict__ rope_cos, const T* __restrict__ rope_sin, T* __restrict__ g_qkv, int H,
                        int n_items, int64_t total_rows) {
  constexpr uint32_t TMEM_COLS = 512u;
  constexpr int CPS = LSEQ / 32;                    // 32-column chunks per sequence
  extern __shared__ uint8_t sm_raw[];
  const uint32_t sm_base = (static_cast<uint32_t>(__cvta_generic_to_shared(sm_raw)) + 1023u) & ~1023u;
  uint8_t* sm_gen = sm_raw + (sm_base - static_cast<uint32_t>(__cvta_generic_to_shared(sm_raw)));
  // tiles: buffer set b: q, k, v, dO at (4 b + 0..3); dS / dS^T: 8, 9; P^T: 10, 11; staging: 12
  const uint32_t sds = sm_base + 8 * TILE, spt = sm_base + 10 * TILE, so = sm_base + 12 * TILE;
  const uint32_t stats = sm_base + 13 * TILE;       // st_m[128], st_li[128], st_D[128]
  const uint32_t bars = stats + 3 * ROWS * 4;
  // barriers (8 B each): 0,1 load[b]; 2 sA; 3 dsA (128); 4 dq; 5 sB; 6 ptB (128); 7 dkv; 8 done (128)
  const uint32_t b_load = bars, b_sA = bars + 16, b_dsA = bars + 24, b_dq = bars + 32, b_sB = bars + 40,
                 b_ptB = bars + 48, b_dkv = bars + 56, b_done = bars + 64;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm_gen + 13 * TILE + 3 * ROWS * 4 + 96);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_qkv)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_do)) : "memory");
    mbar_init(b_load, 1); mbar_init(b_load + 8, 1);
    mbar_init(b_sA, 1); mbar_init(b_dsA, 128); mbar_init(b_dq, 1); mbar_init(b_sB, 1);
    mbar_init(b_ptB, 128); mbar_init(b_dkv, 1); mbar_init(b_done, 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   static_cast<uint32_t>(__cvta_generic_to_shared(tmem_slot))), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tm_s = tmem_base, tm_dp = tmem_base + 128, tm_dq = tmem_base + 256, tm_dk = tmem_base + 320,
                 tm_dv = tmem_base + 384;

  if (warp == 4) {
    if (lane == 0) {
      constexpr uint32_t id_s = idesc<T>(128, 0), id_o = idesc<T>(64, 1);
      auto load = [&](int item, int b) {
        const int hh = item % H, q0 = (item / H) * ROWS;
        const uint32_t bar = b_load + 8 * b, sq = sm_base + (uint32_t)(4 * b) * TILE;
        mbar_expect_tx(bar, 4 * TILE);
        tma_load_2d(&tma_qkv, bar, sq, hh * 64, q0);
        tma_load_2d(&tma_qkv, bar, sq + TILE, (H + hh) * 64, q0);
        tma_load_2d(&tma_qkv, bar, sq + 2 * TILE, (2 * H + hh) * 64, q0);
        tma_load_2d(&tma_do, bar, sq + 3 * TILE, hh * 64, q0);
      };
      if ((int)blockIdx.x < n_items) load(blockIdx.x, 0);
      if ((int)(blockIdx.x + gridDim.x) < n_items) load(blockIdx.x + gridDim.x, 1);
      int it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const uint32_t ph = (uint32_t)(it & 1);
        const int b = it & 1;
        const uint32_t sq = sm_base + (uint32_t)(4 * b) * TILE, sk = sq + TILE, sv = sk + TILE, sdo = sv + TILE;
        mbar_wait(b_load + 8 * b, (uint32_t)((it >> 1) & 1));
        tc_fence_after();
        // ---- pass A: S = q k^T, dP = dO v^T (tensor memory S / dP was released by ptB of the previous item)
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(tm_s, smem_desc(sq) + (uint64_t)(2 * k), smem_desc(sk) + (uint64_t)(2 * k), id_s, (uint32_t)(k != 0));
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(tm_dp, smem_desc(sdo) + (uint64_t)(2 * k), smem_desc(sv) + (uint64_t)(2 * k), id_s, (uint32_t)(k != 0));
        umma_commit(b_sA);
        mbar_wait(b_dsA, ph);                           // dS in shared memory, S / dP read out
        if (it > 0) mbar_wait(b_done, ph ^ 1);          // dq / dk / dv of the previous item read out
        tc_fence_after();
        // ---- dq = dS k (k as the MN-major right-hand side, 16 keys per step)
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_f16(tm_dq, smem_desc(sds + (k >> 2) * TILE) + (uint64_t)(2 * (k & 3)), smem_desc(sk) + (uint64_t)(k * 128),
                   id_o, (uint32_t)(k != 0));
        umma_commit(b_dq);
        // ---- pass B: S^T = k q^T, dP^T = v dO^T
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(tm_s, smem_desc(sk) + (uint64_t)(2 * k), smem_desc(sq) + (uint64_t)(2 * k), id_s, (uint32_t)(k != 0));
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(tm_dp, smem_desc(sv) + (uint64_t)(2 * k), smem_desc(sdo) + (uint64_t)(2 * k), id_s, (uint32_t)(k != 0));
        umma_commit(b_sB);
        mbar_wait(b_ptB, ph);                           // P^T, dS^T in shared memory, S^T / dP^T read out
        tc_fence_after();
        // ---- dv = P^T dO, dk = dS^T q
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_f16(tm_dv, smem_desc(spt + (k >> 2) * TILE) + (uint64_t)(2 * (k & 3)), smem_desc(sdo) + (uint64_t)(k * 128),
                   id_o, (uint32_t)(k != 0));
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_f16(tm_dk, smem_desc(sds + (k >> 2) * TILE) + (uint64_t)(2 * (k & 3)), smem_desc(sq) + (uint64_t)(k * 128),
                   id_o, (uint32_t)(k != 0));
        umma_commit(b_dkv);
        mbar_wait(b_dkv, ph);                           // this buffer set is free: fetch the item after next
        const int nn = item + 2 * (int)gridDim.x;
        if (nn < n_items) load(nn, b);
      }
    }
    __syncwarp();
  } else {
    // ---------------------------------------------------------------- thread = row
    const int row = warp * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    const int c_lo = (warp / CPS) * CPS;              // this row's sequence: chunks c_lo .. c_lo + CPS - 1
    const uint32_t st_m = stats, st_li = stats + ROWS * 4, st_D = stats + 2 * ROWS * 4;
    const int64_t row_stride = (int64_t)3 * H * 64;
    // stage a 32 x 64 tile (this warp's rows, o0 = columns 0-31, o1 = 32-63) and store it coalesced
    auto store_rows = [&](const uint32_t (&w0)[16], const uint32_t (&w1)[16], int q0, int64_t col0) {
      const uint32_t orow = so + (uint32_t)row * 128u;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const uint32_t* src = q < 4 ? w0 + 4 * q : w1 + 4 * (q - 4);
        st_shared_v4(orow + ((uint32_t)(q ^ (row & 7)) << 4), src[0], src[1], src[2], src[3]);
      }
      __syncwarp();
      T* obase = g_qkv + ((int64_t)q0 + warp * 32) * row_stride + col0;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int cidx = i * 32 + lane;                  // 256 chunks: 32 rows x 8
        const int r = cidx >> 3, ch = cidx & 7;
        uint4 v;
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                     : "r"(so + (uint32_t)(warp * 32 + r) * 128u + ((uint32_t)(ch ^ (r & 7)) << 4)) : "memory");
        if ((int64_t)q0 + warp * 32 + r < total_rows) *reinterpret_cast<uint4*>(obase + (int64_t)r * row_stride + ch * 8) = v;
      }
      __syncwarp();
    };
    // gradient w.r.t. the rotary's input: x1' = x1 cos - x2 sin, x2' = x2 cos + x1 sin (models/dit.py:60-75)
    auto unrotate_pack = [&](const uint32_t (&o0)[32], const uint32_t (&o1)[32], int pos, uint32_t (&w0)[16],
                             uint32_t (&w1)[16]) {
      const uint4* pc = reinterpret_cast<const uint4*>(rope_cos + (int64_t)pos * 32);
      const uint4* ps = reinterpret_cast<const uint4*>(rope_sin + (int64_t)pos * 32);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint4 cw = __ldg(pc + i), sw = __ldg(ps + i);
        const uint32_t cws[4] = {cw.x, cw.y, cw.z, cw.w}, sws[4] = {sw.x, sw.y, sw.z, sw.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 c = unpack16<T>(cws[k]), s = unpack16<T>(sws[k]);
          const int d = 8 * i + 2 * k;
          const float g1a = __uint_as_float(o0[d]), g1b = __uint_as_float(o0[d + 1]);
          const float g2a = __uint_as_float(o1[d]), g2b = __uint_as_float(o1[d + 1]);
          w0[4 * i + k] = pack16<T>(g1a * c.x + g2a * s.x, g1b * c.y + g2b * s.y);
          w1[4 * i + k] = pack16<T>(g2a * c.x - g1a * s.x, g2b * c.y - g1b * s.y);
        }
      }
    };
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const uint32_t ph = (uint32_t)(it & 1);
      const int hh = item % H, q0 = (item / H) * ROWS;
      const int pos = (int)(((int64_t)q0 + row) % LSEQ);
      // ================= pass A: this thread = query row
      mbar_wait(b_sA, ph);
      tc_fence_after();
      // (row maximum, sum and D as four independent chains each: the thread phases are latency-bound)
      float e[LSEQ];
      float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int cc = 0; cc < CPS; ++cc) {
        uint32_t r[32];
        tmem_ld32_sync(tm_s + lane_addr + (uint32_t)((c_lo + cc) * 32), r);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          e[cc * 32 + i] = __uint_as_float(r[i]);
          m4[i & 3] = fmaxf(m4[i & 3], e[cc * 32 + i]);
        }
      }
      const float m = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      const float msc = m * kScale2;
      float l4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < LSEQ; ++i) {
        e[i] = ex2_approx(fmaf(e[i], kScale2, -msc));
        l4[i & 3] += e[i];
      }
      const float inv = 1.f / ((l4[0] + l4[1]) + (l4[2] + l4[3]));
      float d4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int cc = 0; cc < CPS; ++cc) {
        uint32_t r[32];
        tmem_ld32_sync(tm_dp + lane_addr + (uint32_t)((c_lo + cc) * 32), r);
#pragma unroll
        for (int i = 0; i < 32; ++i) d4[i & 3] = fmaf(e[cc * 32 + i], __uint_as_float(r[i]), d4[i & 3]);
      }
      const float D = ((d4[0] + d4[1]) + (d4[2] + d4[3])) * inv;
      const float pscale = inv * kScale;
#pragma unroll
      for (int cc = 0; cc < CPS; ++cc) {
        uint32_t r[32], w[16];
        tmem_ld32_sync(tm_dp + lane_addr + (uint32_t)((c_lo + cc) * 32), r);
#pragma unroll
        for (int i = 0; i < 16; ++i)
          w[i] = pack16<T>(e[cc * 32 + 2 * i] * pscale * (__uint_as_float(r[2 * i]) - D),
                           e[cc * 32 + 2 * i + 1] * pscale * (__uint_as_float(r[2 * i + 1]) - D));
        store_chunk16(sds, row, c_lo + cc, w);
      }
      if (CPS < 4) {                                    // keys of the other sequences of the item
        uint32_t z[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) z[i] = 0u;
#pragma unroll 1
        for (int c = 0; c < 4; ++c)
          if (c < c_lo || c >= c_lo + CPS) store_chunk16(sds, row, c, z);
      }
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(st_m + 4 * row), "f"(msc) : "memory");
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(st_li + 4 * row), "f"(inv) : "memory");
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(st_D + 4 * row), "f"(D) : "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // dS visible to the tensor core
      tc_fence_before();
      mbar_arrive(b_dsA);
      asm volatile("bar.sync 1, 128;" ::: "memory");                 // row statistics visible to pass B
      // ---- dq out
      mbar_wait(b_dq, ph);
      tc_fence_after();
      {
        uint32_t o0[32], o1[32], w0[16], w1[16];
        tmem_ld32_sync(tm_dq + lane_addr, o0);
        tmem_ld32_sync(tm_dq + lane_addr + 32u, o1);
        unrotate_pack(o0, o1, pos, w0, w1);
        store_rows(w0, w1, q0, (int64_t)hh * 64);
      }
      // ================= pass B: this thread = key row (dS in shared memory has been consumed by dq)
      mbar_wait(b_sB, ph);
      tc_fence_after();
#pragma unroll
      for (int cc = 0; cc < CPS; ++cc) {
        const int c = c_lo + cc;
        uint32_t r[32], r2[32], wp[16], wd[16];
        tmem_ld32_sync(tm_s + lane_addr + (uint32_t)(c * 32), r);
        tmem_ld32_sync(tm_dp + lane_addr + (uint32_t)(c * 32), r2);
#pragma unroll
        for (int i = 0; i < 8; ++i) {                   // 4 query columns at a time (broadcast reads)
          const float4 mq = ld_shared_f4(st_m + (uint32_t)(c * 32 + 4 * i) * 4);
          const float4 lq = ld_shared_f4(st_li + (uint32_t)(c * 32 + 4 * i) * 4);
          const float4 dq = ld_shared_f4(st_D + (uint32_t)(c * 32 + 4 * i) * 4);
          const float p0 = ex2_approx(fmaf(__uint_as_float(r[4 * i]), kScale2, -mq.x)) * lq.x;
          const float p1 = ex2_approx(fmaf(__uint_as_float(r[4 * i + 1]), kScale2, -mq.y)) * lq.y;
          const float p2 = ex2_approx(fmaf(__uint_as_float(r[4 * i + 2]), kScale2, -mq.z)) * lq.z;
          const float p3 = ex2_approx(fmaf(__uint_as_float(r[4 * i + 3]), kScale2, -mq.w)) * lq.w;
          wp[2 * i] = pack16<T>(p0, p1);
          wp[2 * i + 1] = pack16<T>(p2, p3);
          wd[2 * i] = pack16<T>(p0 * (__uint_as_float(r2[4 * i]) - dq.x) * kScale,
                                p1 * (__uint_as_float(r2[4 * i + 1]) - dq.y) * kScale);
          wd[2 * i + 1] = pack16<T>(p2 * (__uint_as_float(r2[4 * i + 2]) - dq.z) * kScale,
                                    p3 * (__uint_as_float(r2[4 * i + 3]) - dq.w) * kScale);
        }
        store_chunk16(spt, row, c, wp);
        store_chunk16(sds, row, c, wd);
      }
      if (CPS < 4) {
        uint32_t z[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) z[i] = 0u;
#pragma unroll 1
        for (int c = 0; c < 4; ++c)
          if (c < c_lo || c >= c_lo + CPS) {
            store_chunk16(spt, row, c, z);
            store_chunk16(sds, row, c, z);
          }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      tc_fence_before();
      mbar_arrive(b_ptB);
      // ---- dk, dv out
      mbar_wait(b_dkv, ph);
      tc_fence_after();
      {
        uint32_t o0[32], o1[32], w0[16], w1[16];
        tmem_ld32_sync(tm_dk + lane_addr, o0);
        tmem_ld32_sync(tm_dk + lane_addr + 32u, o1);
        unrotate_pack(o0, o1, pos, w0, w1);
        store_rows(w0, w1, q0, (int64_t)(H + hh) * 64);
        tmem_ld32_sync(tm_dv + lane_addr, o0);
        tmem_ld32_sync(tm_dv + lane_addr + 32u, o1);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          w0[i] = pack16<T>(__uint_as_float(o0[2 * i]), __uint_as_float(o0[2 * i + 1]));
          w1[i] = pack16<T>(__uint_as_float(o1[2 * i]), __uint_as_float(o1[2 * i + 1]));
        }
        tc_fence_before();
        mbar_arrive(b_done);
        store_rows(w0, w1, q0, (int64_t)(2 * H + hh) * 64);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

template <typename T, int LSEQ>
int launch(const void* qkv, const void* g_o, const void* cosb, const void* sinb, int batch, int H, int dtype,
           void* g_qkv, cudaStream_t s) {
  const int64_t rows = (int64_t)batch * LSEQ;
  const int64_t cols = (int64_t)3 * H * 64;
  CUtensorMap map_qkv, map_do;
  int rc = make_tma_rows64(&map_qkv, qkv, rows, cols, cols, ROWS, dtype);
  if (rc) return rc;
  rc = make_tma_rows64(&map_do, g_o, rows, (int64_t)H * 64, (int64_t)H * 64, ROWS, dtype);
  if (rc) return rc;
  static bool attr_done[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_done[dev]) {
    DDG_CUDA(cudaFuncSetAttribute(attention_bwd_tc_kernel<T, LSEQ>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  SMEM_BYTES));
    attr_done[dev] = true;
  }
  const int n_items = (int)((rows + ROWS - 1) / ROWS) * H;
  int grid = sm_count();
  if (grid > n_items) grid = n_items;
  attention_bwd_tc_kernel<T, LSEQ><<<grid, 160, SMEM_BYTES, s>>>(
    map_qkv, map_do, reinterpret_cast<const T*>(cosb), reinterpret_cast<const T*>(sinb), reinterpret_cast<T*>(g_qkv), H,
    n_items, rows);
  return check_cuda(cudaGetLastError(), "attention_backward (tcgen05) launch");
}

}  // namespace bwdtc
}  // namespace

// Non-causal, seq_len in {32, 64, 128}, at least 128 token rows, head_dim 64; 16-byte aligned buffers.
bool attention_bwd_tc_supported(int batch, int seq_len, int causal) {
  return !causal && (seq_len == 32 || seq_len == 64 || seq_len == 128) && (int64_t)batch * seq_len >= 128;
}

int launch_attention_bwd_tc(const void* qkv, const void* g_o, const void* rope_cos, const void* rope_sin, int batch,
                            int seq_len, int n_heads, int dtype, void* g_qkv, cudaStream_t s) {
  using namespace bwdtc;
#define DDG_BWD_TC(T)                                                                                   \
  (seq_len == 128 ? launch<T, 128>(qkv, g_o, rope_cos, rope_sin, batch, n_heads, dtype, g_qkv, s)         \
   : seq_len == 64 ? launch<T, 64>(qkv, g_o, rope_cos, rope_sin, batch, n_heads, dtype, g_qkv, s)         \
                   : launch<T, 32>(qkv, g_o, rope_cos, rope_sin, batch, n_heads, dtype, g_qkv, s))
  if (dtype == DDG_F16) return DDG_BWD_TC(__half);
  return DDG_BWD_TC(__nv_bfloat16);
#undef DDG_BWD_TC
}

}  // namespace ddg
