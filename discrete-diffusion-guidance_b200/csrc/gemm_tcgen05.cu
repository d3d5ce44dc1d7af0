// tcgen05 / TMEM / TMA GEMM for the DiT linears (reference models/dit.py:284,
// 312, 242-245, 359-366 -- there they are cuBLAS calls behind nn.Linear under
// autocast).  out = epilogue(A[M,K] @ W[N,K]^T), 16-bit operands, fp32
// accumulation in tensor memory.
//
// Structure: persistent CTA pairs (cluster of 2 = the two SMs of a TPC), one
// CTA per SM, 384 threads each.  A pair owns a 256 x 256 output tile: each CTA
// stages 128 rows of A and 128 rows (half of N) of W per k-block, and the
// leader CTA issues tcgen05.mma.cta_group::2 (M=256, N=256, K=16) over both
// CTAs' shared memory, so every byte of W is fetched once per pair instead of
// once per SM (L2 -> SM traffic per FLOP drops 1.5x, and 6 pipeline stages fit).
//   warp 0     TMA producer: A/W k-blocks -> 128B-swizzled smem ring (6 stages);
//              both CTAs signal the leader's "full" mbarrier
//   warp 1     MMA issuer (leader CTA only): one thread issues the UMMAs and
//              multicast-commits stage-release / accumulator-ready mbarriers
//   warp 2     TMEM allocator (512 columns = 2 accumulators of 128x256 fp32)
//   warps 4-11 epilogue: tcgen05.ld accumulator -> registers -> fused epilogue
//              (bias / GELU-tanh / rotary / gate+residual) -> padded smem slab
//              -> coalesced global stores
// The two TMEM accumulators let the epilogue of tile i overlap the MMAs of
// tile i+1.
#include <cuda.h>

#include <type_traits>

#include "ddg_common.cuh"

namespace ddg {
namespace {

constexpr int BM = 128;
constexpr int BN = 256;
constexpr int BK = 64;           // 64 x 16-bit = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int MAX_STAGES = 6;
// internal epilogue id: GATE_RESIDUAL followed by the next LayerNorm + modulate (fused on chip)
constexpr int EPI_GATE_RESIDUAL_LN = 16;
// internal epilogue id: BIAS_GELU that also writes the pre-activation (ddg_gemm_args.aux); a separate
// instantiation so that the plain BIAS_GELU kernel of the sampling step carries no extra branch
constexpr int EPI_BIAS_GELU_AUX = 17;
// the fused-LayerNorm variant trades one pipeline stage for its per-column LayerNorm vectors
// 16-bit outputs with no work in the coalesced phase leave through TMA stores (cp.async.bulk.tensor
// shared -> global): thread = row writes into a 64B-swizzled 32 x 32 slab, one lane issues the copy
template <int EPI> __host__ __device__ constexpr bool tma_out() {
  return EPI == DDG_EPI_STORE || EPI == DDG_EPI_QKV_ROPE || EPI == DDG_EPI_BIAS_GELU;
}
#ifndef DDG_TMA_OUT_STAGES
#define DDG_TMA_OUT_STAGES 5
#endif
template <int EPI> __host__ __device__ constexpr int stages_for() {
  return EPI == EPI_GATE_RESIDUAL_LN ? 5 : tma_out<EPI>() ? DDG_TMA_OUT_STAGES : MAX_STAGES;
}
constexpr int A_STAGE_BYTES = BM * BK * 2;
constexpr int BNH = BN / 2;          // W rows staged by each CTA of the pair
constexpr int B_STAGE_BYTES = BNH * BK * 2;
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr int NUM_THREADS = 384;
constexpr int EPI_WARP0 = 4;
constexpr int NUM_EPI_WARPS = 8;
constexpr int TMEM_COLS = 512;
constexpr int EPI_ROW_BYTES = 80;                  // 64 B payload + 16 B pad: conflict-free 128-bit stores
constexpr int EPI_SLAB_BYTES = 32 * EPI_ROW_BYTES;  // one warp's 32 accumulator rows
constexpr int TMA_SLAB_BYTES = 32 * 64;            // TMA-store slab: 32 rows x 64 B, SWIZZLE_64B
#ifndef DDG_TMA_OUT_BUFS
#define DDG_TMA_OUT_BUFS 2
#endif
constexpr int TMA_OUT_BUFS = DDG_TMA_OUT_BUFS;     // slabs per epilogue warp (2: the copy of chunk c reads while c + 1 is written)
constexpr int LN_ROW_BYTES = 144;                  // fused LayerNorm: 32 fp32 columns (128 B) + 16 B pad per row
constexpr int LN_SLAB_BYTES = 32 * LN_ROW_BYTES;
constexpr int EPI_BIAS_BYTES = NUM_EPI_WARPS * (BN / 2) * 4;   // fp32 bias (and gate) of each epilogue warp's 128 columns
constexpr int EPI_LN_BYTES = 3 * EPI_BIAS_BYTES;   // fused LayerNorm: weight, 1 + scale, shift (fp32) per epilogue warp
template <int EPI> __host__ __device__ constexpr int smem_bytes() {
  return stages_for<EPI>() * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ +
         8 * (EPI == EPI_GATE_RESIDUAL_LN ? LN_SLAB_BYTES : tma_out<EPI>() ? TMA_OUT_BUFS * TMA_SLAB_BYTES : EPI_SLAB_BYTES) +
         2 * EPI_BIAS_BYTES +
         (EPI == EPI_GATE_RESIDUAL_LN ? EPI_LN_BYTES : 0);
}

struct GemmParams {
  void* out;
  const void* bias;
  const void* gate;
  const void* rope_cos;
  const void* rope_sin;
  int64_t ld_out, gate_stride, rows_per_mod;
  int M, N, K;
  int seq_len, rope_cols;
  float2* stats;        // optional per-(128-column chunk, row) softmax partials
  int64_t stats_ld;
  int mask_col;         // column excluded from the partials (-1: none)
  int accumulate;       // ACCUM_F32: add to the existing output
  int rope_pos0;        // QKV_ROPE: position offset
  const int32_t* live_seqs;    // optional device scalar: only the first *live_seqs * seq_len rows of A exist
  const int32_t* out_seq_map;  // optional (STORE): compact sequence s is written to output sequence out_seq_map[s]
  int stream_out;       // output larger than L2 (the vocab head's logits): evict-first stores,
                        // evict-last operand loads, so the write stream does not push W out of L2
  // fused LayerNorm + modulate of the updated residual rows (EPI_GATE_RESIDUAL_LN)
  const float* ln_w;    // (N,) fp32
  const void* ln_shift; // rows of `ln_dtype` (stride ln_mod_stride, row = m / rows_per_mod) or null
  const void* ln_scale; // 1 + scale, same layout
  void* ln_out;         // (M, N) `ln_dtype`
  const float* residual; // GATE_RESIDUAL: rows the delta is added to (null: `out` itself, in place)
  void* aux;            // BIAS_GELU: optional pre-activation output; GELU_BWD: pre-activation input (16-bit, ld_aux)
  int64_t ld_aux;
  int64_t out_rows;     // rows of `out` a TMA-store epilogue may address (M; the whole buffer with out_seq_map)
  uint8_t* ln_work;     // per 128-row block: [counter word, padded to 64 B][128 x N / 128 partial words of 8 B]
                        // -- a layout that does not depend on M, so launches over different row counts
                        // can share one buffer
  int64_t ln_mod_stride;
  int ln_dtype;
};

// ------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  asm volatile(
    "{\n\t"
    ".reg .pred P1;\n\t"
    "WAIT_LOOP:\n\t"
    "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
    "@P1 bra.uni WAIT_DONE;\n\t"
    "bra.uni WAIT_LOOP;\n\t"
    "WAIT_DONE:\n\t"
    "}" ::"r"(addr), "r"(parity)
    : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
    "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
    ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
    : "memory");
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr)
               : "memory");
  return v;
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of the same shared-memory location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// 2-CTA TMA load: data lands in this CTA's smem, bytes are credited to the
// mbarrier at `bar_cluster_addr` (the leader's)
__device__ __forceinline__ void tma_load_2d_2cta(const CUtensorMap* m, uint32_t bar_cluster_addr, void* dst, int c0,
                                                 int c1) {
  asm volatile(
    "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
    ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
    : "memory");
}
// L2 eviction-priority policies (PTX createpolicy) for cache-hinted loads / stores
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void tma_load_2d_2cta_hint(const CUtensorMap* m, uint32_t bar_cluster_addr, void* dst,
                                                      int c0, int c1, uint64_t policy) {
  asm volatile(
    "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
    "[%0], [%1, {%3, %4}], [%2], %5;"
    ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "l"(policy)
    : "memory");
}
__device__ __forceinline__ void st_global_v4_hint(void* ptr, uint4 v, uint64_t policy) {
  asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(ptr), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w), "l"(policy)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// arrives (once all previously issued UMMAs retire) on the mbarrier at this
// smem offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile(
    "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
      smem_u32(bar)), "h"((uint16_t)3)
    : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, 256 x 256 x 16 over the CTA pair (each CTA:
// 128 rows of A, 128 rows of B, 128 x 256 fp32 of D), 16-bit inputs
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
    "{\n\t"
    ".reg .pred p;\n\t"
    "setp.ne.b32 p, %4, 0;\n\t"
    "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
    "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
    : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
      "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
      "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
      "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
    : "r"(taddr));
}
// load + wait in one statement: the outputs are only defined after the wait, so
// nothing can be scheduled in between and no register copies are needed
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
// The registers are listed as in/out operands so that no use of them can be
// scheduled ahead of the wait.
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]),
                 "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]),
                 "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]),
                 "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}

// registers -> tensor memory, same 32x32b layout as tmem_ld32 (lane = row, 32 consecutive columns)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
  asm volatile(
    "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
    "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
    "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
    ::"r"(taddr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]),
      "f"(v[8]), "f"(v[9]), "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15]),
      "f"(v[16]), "f"(v[17]), "f"(v[18]), "f"(v[19]), "f"(v[20]), "f"(v[21]), "f"(v[22]), "f"(v[23]),
      "f"(v[24]), "f"(v[25]), "f"(v[26]), "f"(v[27]), "f"(v[28]), "f"(v[29]), "f"(v[30]), "f"(v[31])
    : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ int ld_relaxed_gpu(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float2 ld_relaxed_gpu_f2(const float2* p) {
  float2 v;
  asm volatile("ld.relaxed.gpu.global.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint64_t ld_relaxed_gpu_b64(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.relaxed.gpu.global.b64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_gpu_b64(void* p, uint64_t v) {
  asm volatile("st.relaxed.gpu.global.b64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void red_release_gpu_add(int* p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// K-major, 128B-swizzled shared-memory matrix descriptor (PTX ISA "matrix
// descriptor"): start>>4 | LBO(ignored for swizzled K-major)=1 | SBO = 8 rows
// * 128 B = 1024 B | version 1 | layout SWIZZLE_128B (2).
__device__ __forceinline__ uint64_t make_smem_desc(const void* p) {
  const uint32_t addr = smem_u32(p);
  uint64_t d = (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Instruction descriptor, kind::f16: D=f32 (bits 4-5 = 1), A/B format (0 f16,
// 1 bf16) at bits 7-9 / 10-12, both K-major, N>>3 at 17-22, M>>4 at 24-28.
template <typename T> __host__ __device__ constexpr uint32_t make_idesc();
template <> __host__ __device__ constexpr uint32_t make_idesc<__half>() {
  return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);
}
template <> __host__ __device__ constexpr uint32_t make_idesc<__nv_bfloat16>() {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);
}

// ------------------------------------------------------------------ 16-bit helpers
template <typename T> struct H2;
template <> struct H2<__half> {
  // packed product in T: the exact product of two 11-bit significands rounded once (= fp32 multiply + cvt.rn)
  static __device__ __forceinline__ uint32_t mul(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
  }
  static __device__ __forceinline__ uint32_t pack(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  static __device__ __forceinline__ float2 unpack(uint32_t w) {
    return __half22float2(*reinterpret_cast<__half2*>(&w));
  }
};
template <> struct H2<__nv_bfloat16> {
  static __device__ __forceinline__ uint32_t mul(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
  }
  static __device__ __forceinline__ uint32_t pack(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  static __device__ __forceinline__ float2 unpack(uint32_t w) {
    return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
  }
};

// loads 32 consecutive 16-bit values (64 B, 16B-aligned) as fp32
template <typename T>
__device__ __forceinline__ void load32(const T* p, float (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint4 r = __ldg(reinterpret_cast<const uint4*>(p) + i);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = H2<T>::unpack(w[k]);
      v[i * 8 + 2 * k] = f.x;
      v[i * 8 + 2 * k + 1] = f.y;
    }
  }
}

// stores 32 fp32 values as 16-bit at out[0..32), columns >= n_valid dropped
template <typename T>
__device__ __forceinline__ void store32(T* out, const float (&v)[32], int n_valid) {
  if (n_valid >= 32) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 r;
      r.x = H2<T>::pack(v[i * 8 + 0], v[i * 8 + 1]);
      r.y = H2<T>::pack(v[i * 8 + 2], v[i * 8 + 3]);
      r.z = H2<T>::pack(v[i * 8 + 4], v[i * 8 + 5]);
      r.w = H2<T>::pack(v[i * 8 + 6], v[i * 8 + 7]);
      reinterpret_cast<uint4*>(out)[i] = r;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i < n_valid) out[i] = Cvt<T>::from_f32(v[i]);
  }
}

// rounds 32 fp32 values to T and back with packed conversions (F2FP + unpack)
template <typename T>
__device__ __forceinline__ void round32(float (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float2 f = H2<T>::unpack(H2<T>::pack(v[2 * i], v[2 * i + 1]));
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}

constexpr float kLog2e = 1.4426950408889634f;
__device__ __forceinline__ float ex2f_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// packed f32x2 arithmetic: one issue slot for two results (sm_100 FADD2 / FMUL2 / FFMA2)
__device__ __forceinline__ float2 add2(float2 x, float2 y) {
  float2 r;
  asm("{\n\t"
      ".reg .b64 a, b, d;\n\t"
      "mov.b64 a, {%2, %3};\n\t"
      "mov.b64 b, {%4, %5};\n\t"
      "add.f32x2 d, a, b;\n\t"
      "mov.b64 {%0, %1}, d;\n\t"
      "}" : "=f"(r.x), "=f"(r.y) : "f"(x.x), "f"(x.y), "f"(y.x), "f"(y.y));
  return r;
}
__device__ __forceinline__ float2 sub2(float2 x, float2 y) {
  float2 r;
  asm("{\n\t"
      ".reg .b64 a, b, d;\n\t"
      "mov.b64 a, {%2, %3};\n\t"
      "mov.b64 b, {%4, %5};\n\t"
      "sub.f32x2 d, a, b;\n\t"
      "mov.b64 {%0, %1}, d;\n\t"
      "}" : "=f"(r.x), "=f"(r.y) : "f"(x.x), "f"(x.y), "f"(y.x), "f"(y.y));
  return r;
}
__device__ __forceinline__ float2 mul2(float2 x, float2 y) {
  float2 r;
  asm("{\n\t"
      ".reg .b64 a, b, d;\n\t"
      "mov.b64 a, {%2, %3};\n\t"
      "mov.b64 b, {%4, %5};\n\t"
      "mul.f32x2 d, a, b;\n\t"
      "mov.b64 {%0, %1}, d;\n\t"
      "}" : "=f"(r.x), "=f"(r.y) : "f"(x.x), "f"(x.y), "f"(y.x), "f"(y.y));
  return r;
}
__device__ __forceinline__ float2 fma2(float2 x, float2 y, float2 z) {
  float2 r;
  asm("{\n\t"
      ".reg .b64 a, b, c, d;\n\t"
      "mov.b64 a, {%2, %3};\n\t"
      "mov.b64 b, {%4, %5};\n\t"
      "mov.b64 c, {%6, %7};\n\t"
      "fma.rn.f32x2 d, a, b, c;\n\t"
      "mov.b64 {%0, %1}, d;\n\t"
      "}" : "=f"(r.x), "=f"(r.y) : "f"(x.x), "f"(x.y), "f"(y.x), "f"(y.y), "f"(z.x), "f"(z.y));
  return r;
}

// tanh-GELU of a column pair (same formula as gelu_tanh below)
__device__ __forceinline__ float2 gelu_tanh2(float2 v) {
  const float2 t = fma2(mul2(v, v), make_float2(-0.10294324f, -0.10294324f), make_float2(-2.3022082f, -2.3022082f));
  const float2 z = mul2(v, t);
  float ex, ey, rx, ry;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(z.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ey) : "f"(z.y));
  const float2 d = add2(make_float2(ex, ey), make_float2(1.f, 1.f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rx) : "f"(d.x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(ry) : "f"(d.y));
  return mul2(v, make_float2(rx, ry));
}

// gelu_tanh_grad() of ddg_common.cuh for a column pair, operation for operation in packed f32x2
__device__ __forceinline__ float2 gelu_tanh_grad2(float2 v) {
  const float2 one = make_float2(1.f, 1.f);
  const float2 v2 = mul2(v, v);
  const float2 arg = mul2(v, fma2(make_float2(kGeluK1, kGeluK1), v2, one));
  const float2 z = mul2(arg, make_float2(kGeluExp, kGeluExp));
  float ex, ey, sx, sy;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(z.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ey) : "f"(z.y));
  const float2 d = add2(make_float2(ex, ey), one);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(sx) : "f"(d.x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(sy) : "f"(d.y));
  const float2 s = make_float2(sx, sy);
  const float2 w = mul2(v, fma2(make_float2(kGeluD1, kGeluD1), v2, make_float2(kGeluD0, kGeluD0)));
  const float2 oms = fma2(s, make_float2(-1.f, -1.f), one);
  return mul2(s, fma2(w, oms, one));
}

__device__ __forceinline__ float gelu_tanh(float v) {
  // 0.5 v (1 + tanh(u)) == v * sigmoid(2u),  u = sqrt(2/pi) (v + 0.044715 v^3)
  // exp(-2u) = ex2(v * (c1 + c2 v^2)), c1 = -2 sqrt(2/pi) log2(e), c2 = 0.044715 c1
  const float z = v * fmaf(v * v, -0.10294324f, -2.3022082f);
  float ez, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ez) : "f"(z));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + ez));
  return v * r;
}

// Tile order.  Default: N inner (consecutive tiles = the N-blocks of one M-block, so the
// clusters running at the same time share one A block and W -- a few MB -- stays in L2).
// The vocab head (120 N-blocks, W = 47 MB, 8 GB of output streaming through L2) does not
// keep W resident that way: ncu showed 3-4.4 GB of DRAM reads per launch against 0.25 GB
// algorithmic.  `n_panel` > 0 splits N into panels of `n_panel` blocks and finishes all
// M-blocks of one panel before the next: the W panel (24 MB for two panels) is what has to
// stay in L2, at the price of reading A (0.2 GB) once per panel.
__device__ __forceinline__ void tile_coords(int tile, int num_m, int num_n, int n_panel, int& mb, int& nb) {
  if (n_panel <= 0 || n_panel >= num_n) {
    mb = tile / num_n;
    nb = tile % num_n;
    return;
  }
  const int per_panel = num_m * n_panel;           // tiles of a full panel
  const int pnl = tile / per_panel;
  const int n_first = pnl * n_panel;
  const int width = min(n_panel, num_n - n_first); // the last panel may be narrower
  const int r = tile - pnl * per_panel;            // per_panel is exact for every earlier panel
  mb = r / width;
  nb = n_first + (r - mb * width);
}

// Developer build only (-DDDG_LN_PROF, scripts/ln_epilogue_prof.py): clock64() per phase of the
// fused-LayerNorm epilogue, summed by one warp of one CTA.
#ifdef DDG_LN_PROF
__device__ unsigned long long g_ln_prof[8];
#define DDG_LNP(i)                                                                   \
  if (e == 0 && lane == 0 && blockIdx.x == 2) {                                      \
    const long long lnp_now = clock64();                                             \
    atomicAdd(&g_ln_prof[i], (unsigned long long)(lnp_now - lnp_t));                 \
    lnp_t = lnp_now;                                                                 \
  }
#else
#define DDG_LNP(i)
#endif

// ------------------------------------------------------------------ kernel
template <typename T, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
            const __grid_constant__ CUtensorMap tma_o, const __grid_constant__ CUtensorMap tma_o2,
            const GemmParams p) {
  constexpr int STAGES = stages_for<EPI>();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;
  constexpr int SLAB_BYTES = EPI == EPI_GATE_RESIDUAL_LN ? LN_SLAB_BYTES
                             : tma_out<EPI>() ? TMA_OUT_BUFS * TMA_SLAB_BYTES : EPI_SLAB_BYTES;
  // the epilogue warps' slabs follow the stages (1024-byte aligned: the TMA-store slabs are swizzled),
  // then the barriers, then the per-column vectors
  uint8_t* smem_epi = smem + STAGES * STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_epi + NUM_EPI_WARPS * SLAB_BYTES);
  uint64_t* full_bar = bars;                 // [STAGES]
  uint64_t* empty_bar = bars + STAGES;       // [STAGES]
  uint64_t* tmem_full = bars + 2 * STAGES;   // [2]
  uint64_t* tmem_empty = bars + 2 * STAGES + 2;  // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  uint8_t* smem_bias = reinterpret_cast<uint8_t*>(bars) + 256;
  uint8_t* smem_gate = smem_bias + EPI_BIAS_BYTES;
  uint8_t* smem_ln = smem_gate + EPI_BIAS_BYTES;   // EPI_GATE_RESIDUAL_LN only

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  // dynamic batch (per-sequence logits cache of the MDLM loop): the live row count is a device scalar
  const int M = p.live_seqs ? min(p.M, __ldg(p.live_seqs) * p.seq_len) : p.M;
  const int num_m = (M + 2 * BM - 1) / (2 * BM);
  const int num_n = (p.N + BN - 1) / BN;
  const int num_tiles = num_m * num_n;
  const int num_kb = (p.K + BK - 1) / BK;
  const int n_panel = p.stream_out ? (num_n + 1) / 2 : 0;   // two N panels for the vocab head

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    if (tma_out<EPI>() || EPI == EPI_GATE_RESIDUAL_LN) tma_prefetch_desc(&tma_o);
    if (EPI == EPI_GATE_RESIDUAL_LN) tma_prefetch_desc(&tma_o2);
  }
  if (warp == 1 && lane == 0) {
    // full: the leader producer's arrival + both CTAs' TMA bytes, leader's copy is the live one;
    // empty / tmem_full: one multicast commit from the leader's MMA thread;
    // tmem_empty (leader's copy): one arrival per epilogue warp of both CTAs
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 2 * NUM_EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_ptr, TMEM_COLS);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint64_t keep = policy_evict_last();
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        int mb, nb;
        tile_coords(tile, num_m, num_n, n_panel, mb, nb);
        const int m0 = mb * (2 * BM) + (int)cta_rank * BM;     // this CTA's A rows
        const int nb0 = nb * BN + (int)cta_rank * BNH;         // this CTA's half of W
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          const uint32_t full_leader = map_to_cta(smem_u32(&full_bar[stage]), 0);
          // only the leader arrives; its expect-tx covers both CTAs' bytes
          if (leader) mbar_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
          if (p.stream_out) {
            // W (and the A block the concurrent tiles share) are re-read by every M-block:
            // keep them in L2 ahead of the output stream
            tma_load_2d_2cta_hint(&tma_a, full_leader, smem_a + stage * A_STAGE_BYTES, kb * BK, m0, keep);
            tma_load_2d_2cta_hint(&tma_b, full_leader, smem_b + stage * B_STAGE_BYTES, kb * BK, nb0, keep);
          } else {
            tma_load_2d_2cta(&tma_a, full_leader, smem_a + stage * A_STAGE_BYTES, kb * BK, m0);
            tma_load_2d_2cta(&tma_b, full_leader, smem_b + stage * B_STAGE_BYTES, kb * BK, nb0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer (leader CTA)
    if (lane == 0 && leader) {
      constexpr uint32_t idesc = make_idesc<T>();
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
#ifdef DDG_LN_PROF
        const long long mma_t0 = clock64();
#endif
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
#ifdef DDG_LN_PROF
        if (blockIdx.x == 2) atomicAdd(&g_ln_prof[6], (unsigned long long)(clock64() - mma_t0));
#endif
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t adesc = make_smem_desc(smem_a + stage * A_STAGE_BYTES);
          const uint64_t bdesc = make_smem_desc(smem_b + stage * B_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // advance 16 elements (32 B) along K inside the 128B swizzle row
            umma_f16(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                     (uint32_t)((kb | k) != 0));
          }
          umma_commit(&empty_bar[stage]);   // frees the smem stage when the MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[acc]);       // accumulator ready for the epilogue
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
    __syncwarp();
  } else if (warp >= EPI_WARP0) {
    // ------------------------------------------------ epilogue (8 warps)
    // Warp e owns accumulator rows 32*(e%4).. (its TMEM lane quarter) and
    // columns 128*(e/4)..: two warps per scheduler keep the FMA/MUFU work of
    // the fused epilogues latency-hidden.  Register phase: thread = row (what
    // tcgen05.ld 32x32b gives).  Finished values pass through a warp-private
    // padded shared-memory slab (32 rows x 80 B) and come back with lanes
    // running along the row, so each global access covers 8 rows x 64
    // contiguous bytes instead of 32 rows x 16 bytes.
    const int e = warp - EPI_WARP0;
    const int quarter = e & 3;                  // == warp % 4: TMEM lanes 32*quarter..
    const int chalf = e >> 2;                   // column half of the tile
    const int row_in_tile = quarter * 32 + lane;
    const uint32_t stg = smem_u32(smem_epi + e * SLAB_BYTES);
    const uint32_t stg_wr = stg + (uint32_t)lane * EPI_ROW_BYTES;
    const int crow = lane >> 2;                 // coalesced phase: row in an 8-row slab
    const int cchunk = lane & 3;                // 16-byte chunk of the 64-byte row
    const uint32_t stg_rd = stg + (uint32_t)crow * EPI_ROW_BYTES + (uint32_t)cchunk * 16;
    int acc = 0;
    uint32_t acc_phase = 0;
    const uint32_t tmem_empty_leader0 = map_to_cta(smem_u32(&tmem_empty[0]), 0);
    uint32_t tma_chunk = 0;                     // TMA-store epilogues: chunks this warp has sent (slab = parity)
#ifdef DDG_LN_PROF
    long long lnp_t = clock64();
#endif
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      int mb, nb;
      tile_coords(tile, num_m, num_n, n_panel, mb, nb);
      const int m0 = mb * (2 * BM) + (int)cta_rank * BM;
      const int n0 = nb * BN + chalf * (BN / 2);
      const int m = m0 + row_in_tile;
      const bool row_ok = m < M;
      const int mw = m0 + quarter * 32;         // first row of this warp's slab
      // fused LayerNorm: fp32 tiles pass through the slab as 32 rows x (128 + 16) B; with lanes
      // along the rows one access covers 4 rows x 128 contiguous bytes (r4 = row, c8 = 16-byte chunk)
      const int r4 = lane >> 3, c8 = lane & 7;
      const uint32_t xs_row = stg + (uint32_t)lane * LN_ROW_BYTES;
      const uint32_t xs_co = stg + (uint32_t)r4 * LN_ROW_BYTES + (uint32_t)c8 * 16;
      const bool cols_ok = n0 < p.N;
      float4 xn[8];                             // x of the next 32 columns (lanes along the rows)
      auto ln_prefetch_x = [&](int nc) {
        const float* xb = (p.residual ? p.residual : reinterpret_cast<const float*>(p.out)) +
                          (int64_t)(mw + r4) * p.ld_out + nc + c8 * 4;
#pragma unroll
        for (int i = 0; i < 8; ++i)
          xn[i] = (mw + 4 * i + r4 < M) ? *reinterpret_cast<const float4*>(xb + (int64_t)(4 * i) * p.ld_out)
                                        : make_float4(0.f, 0.f, 0.f, 0.f);
      };
      if constexpr (EPI == EPI_GATE_RESIDUAL_LN) {
        // the residual rows do not depend on the accumulator: their first chunk is on its way
        // (and the rest of the warp's 32 x 128 block is pulled into L2) while the MMAs finish
        if (cols_ok) {
          ln_prefetch_x(n0);
          if (mw + lane < M) {
            const float* xb = (p.residual ? p.residual : reinterpret_cast<const float*>(p.out)) +
                              (int64_t)(mw + lane) * p.ld_out + n0;
#pragma unroll
            for (int k = 1; k < 4; ++k) asm volatile("prefetch.global.L2 [%0];" ::"l"(xb + k * 32));
          }
        }
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      DDG_LNP(0)
      const uint32_t taddr =
        tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN + chalf * (BN / 2));

      if constexpr (EPI == EPI_GATE_RESIDUAL_LN) {
        // -------- x += round(gate * round(acc + bias)) AND the LayerNorm + modulate of the updated
        // row (models/dit.py:312-321: the norm that follows every residual add), without a second
        // trip of x through HBM.  A row's 768 columns are spread over num_n tiles that different
        // CTA pairs process at about the same time (N-inner tile order), so:
        //   pass 1   x_new = x + delta, formed with thread = row (x comes in through the transposing
        //            slab), PARKED in this tile's accumulator columns of tensor memory (tcgen05.st)
        //            and reduced to a per-row (mean, M2) over the warp's 128 columns (shifted sums);
        //   exchange each partial travels as ONE 64-bit word -- mean | M2 with the block's epoch
        //            parity in M2's sign bit (M2 >= 0) -- so data and validity arrive together and
        //            no fence is needed; a reader polls the words of its row until the parity
        //            matches.  Every owner writes all 32 rows of its slab (also rows past M), so a
        //            block's words always carry the parity of the block's last visit; the block's
        //            counter word is (parity << 8) | departures and the last owner to leave flips
        //            the parity for the next launch;
        //   store    while the other owners' words travel, x_new goes from tensor memory to the
        //            fp32 residual stream (coalesced through the slab);
        //   pass 2   x_new comes back from tensor memory once more, ((x - mean) * rstd) * w
        //            [* (1 + scale) + shift], one rounding to the LayerNorm's output type, coalesced
        //            16-bit stores.
        // Combining the partials in slot order gives every owner of a row the same mean / rstd.
        // No deadlock: the owners of a block's tiles are distinct resident CTA pairs, and a pair's
        // epilogue only ever waits for tiles numbered within num_n of its own current tile, whose
        // mainloops depend on strictly earlier blocks only.
        const int nvalid = p.N >> 7;                       // 128-column runs per row (host: N % 128 == 0, <= 8)
        if (cols_ok) {
          const bool have_bias = p.bias != nullptr;
          const bool have_mod = p.ln_shift != nullptr;
          const uint32_t bias_s = smem_u32(smem_bias) + (uint32_t)e * (BN / 2) * 4;
          const uint32_t gate_s = smem_u32(smem_gate) + (uint32_t)e * (BN / 2) * 4;
          const uint32_t lnw_s = smem_u32(smem_ln) + (uint32_t)e * (BN / 2) * 4;
          const uint32_t lng_s = lnw_s + EPI_BIAS_BYTES, lns_s = lng_s + EPI_BIAS_BYTES;
          // host guarantees rows_per_mod % 32 == 0: the warp's 32 rows share one conditioning row
          const int rpm = p.rows_per_mod > 0x7fffffff ? 0x7fffffff : (int)p.rows_per_mod;
          const int mrow = (mw < M ? mw : 0) / rpm;
          const T* gate_row = p.gate ? reinterpret_cast<const T*>(p.gate) + (int64_t)mrow * p.gate_stride : nullptr;
          if (have_bias) {
            const uint2 raw = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const T*>(p.bias) + n0) + lane);
            const float2 b01 = H2<T>::unpack(raw.x), b23 = H2<T>::unpack(raw.y);
            st_shared_v4(bias_s + 16 * lane, __float_as_uint(b01.x), __float_as_uint(b01.y),
                         __float_as_uint(b23.x), __float_as_uint(b23.y));
          }
          if (gate_row) {   // kept 16-bit: gate * round(acc + bias) is ONE packed multiply in T (exact product,
            // one rounding = the fp32 multiply + conversion it replaces, bit for bit)
            const uint2 raw = __ldg(reinterpret_cast<const uint2*>(gate_row + n0) + lane);
            asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(gate_s + 8 * lane), "r"(raw.x), "r"(raw.y) : "memory");
          }
          {
            const float4 w4 = __ldg(reinterpret_cast<const float4*>(p.ln_w + n0) + lane);
            st_shared_v4(lnw_s + 16 * lane, __float_as_uint(w4.x), __float_as_uint(w4.y), __float_as_uint(w4.z),
                         __float_as_uint(w4.w));
          }
          __syncwarp();
          // the block's counter word, read before this warp publishes anything: no owner can have
          // departed yet, so it is (parity << 8) exactly (used after pass 1: latency hidden)
          uint8_t* blk = p.ln_work + (int64_t)(mb * 2 + (int)cta_rank) * (64 + BM * nvalid * 8);
          int* cnt = reinterpret_cast<int*>(blk);
          uint64_t* part = reinterpret_cast<uint64_t*>(blk + 64) + (int64_t)row_in_tile * nvalid;   // this row's words
          const int epoch_word = ld_relaxed_gpu(cnt);
          // the slab doubles as the source of the previous tile's last TMA stores: they have been read
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          __syncwarp();
          // ---- pass 1
          float s1 = 0.f, s2 = 0.f, c0 = 0.f;
#pragma unroll 1
          for (int g = 0; g < BN / 64; ++g) {
            const int nc = n0 + g * 32;
            uint32_t r0[32];
            tmem_ld32(taddr + (uint32_t)(g * 32), r0);      // in flight while x goes through the slab
            // x of this chunk (prefetched, lanes along the rows) -> slab
#pragma unroll
            for (int i = 0; i < 8; ++i)
              st_shared_v4(xs_co + (uint32_t)(4 * i) * LN_ROW_BYTES, __float_as_uint(xn[i].x), __float_as_uint(xn[i].y),
                           __float_as_uint(xn[i].z), __float_as_uint(xn[i].w));
            __syncwarp();
            if (g + 1 < BN / 64) ln_prefetch_x(nc + 32);
            tmem_ld_wait(r0);
            float2 v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = make_float2(__uint_as_float(r0[2 * i]), __uint_as_float(r0[2 * i + 1]));
            if (have_bias) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const uint4 b = ld_shared_v4(bias_s + (uint32_t)(g * 128 + 16 * i));
                v[2 * i] = add2(v[2 * i], make_float2(__uint_as_float(b.x), __uint_as_float(b.y)));
                v[2 * i + 1] = add2(v[2 * i + 1], make_float2(__uint_as_float(b.z), __uint_as_float(b.w)));
              }
            }
            uint32_t dw[16];                   // round(acc + bias), packed
#pragma unroll
            for (int i = 0; i < 16; ++i) dw[i] = H2<T>::pack(v[i].x, v[i].y);
            if (gate_row) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const uint4 b = ld_shared_v4(gate_s + (uint32_t)(g * 64 + 16 * i));
                dw[4 * i] = H2<T>::mul(dw[4 * i], b.x);
                dw[4 * i + 1] = H2<T>::mul(dw[4 * i + 1], b.y);
                dw[4 * i + 2] = H2<T>::mul(dw[4 * i + 2], b.z);
                dw[4 * i + 3] = H2<T>::mul(dw[4 * i + 3], b.w);
              }
            }
            // thread = row: x_new = x + delta (packed f32x2), then the shifted sums of this chunk
            float xr[32];
            float2 s1p = make_float2(0.f, 0.f), s2p = make_float2(0.f, 0.f);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const uint4 u = ld_shared_v4(xs_row + q * 16);
              const float2 a0 = add2(make_float2(__uint_as_float(u.x), __uint_as_float(u.y)), H2<T>::unpack(dw[2 * q]));
              const float2 a1 = add2(make_float2(__uint_as_float(u.z), __uint_as_float(u.w)), H2<T>::unpack(dw[2 * q + 1]));
              xr[4 * q] = a0.x; xr[4 * q + 1] = a0.y; xr[4 * q + 2] = a1.x; xr[4 * q + 3] = a1.y;
            }
            __syncwarp();                      // slab free for the next chunk
            if (g == 0) c0 = xr[0];
            const float2 c02 = make_float2(c0, c0);
#pragma unroll
            for (int k = 0; k < 16; ++k) {
              const float2 d = sub2(make_float2(xr[2 * k], xr[2 * k + 1]), c02);
              s1p = add2(s1p, d);
              s2p = fma2(d, d, s2p);
            }
            s1 += s1p.x + s1p.y;
            s2 += s2p.x + s2p.y;
            tmem_st32(taddr + (uint32_t)(g * 32), xr);
          }
          tmem_st_wait();
          DDG_LNP(1)
          // ---- publish this warp's partial of each of its 32 rows (also rows past M)
          const int target = 4 * nvalid;
          const uint32_t tagbit = (((uint32_t)epoch_word >> 8) + 1u) & 1u;
          {
            const float mean_i = c0 + s1 * (1.f / 128.f);
            const float m2_i = fmaxf(s2 - s1 * s1 * (1.f / 128.f), 0.f);
            const uint64_t word = (uint64_t)__float_as_uint(mean_i) |
                                  ((uint64_t)((__float_as_uint(m2_i) & 0x7fffffffu) | (tagbit << 31)) << 32);
            st_relaxed_gpu_b64(part + (nb * 2 + chalf), word);
          }
          // ---- x_new -> fp32 residual stream while the other owners' words travel: thread = row writes
          // 16 columns (64 B) at a time into one of two 64B-swizzled 32 x 64 B slabs, one lane sends the
          // 32 x 16 box with a TMA store (rows >= M clipped by the tensor map), the slabs alternate
          const uint32_t sw64 = (uint32_t)(lane >> 1) & 3u;
#pragma unroll 1
          for (int g = 0; g < BN / 64; ++g) {
            const int nc = n0 + g * 32;
            uint32_t r[32];
            tmem_ld32_sync(taddr + (uint32_t)(g * 32), r);
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              const uint32_t slab = stg + (tma_chunk & 1u) * TMA_SLAB_BYTES;
              if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
              __syncwarp();
              const uint32_t rowp = slab + (uint32_t)lane * 64u;
#pragma unroll
              for (int q = 0; q < 4; ++q)
                st_shared_v4(rowp + (((uint32_t)q ^ sw64) << 4), r[16 * hh + 4 * q], r[16 * hh + 4 * q + 1],
                             r[16 * hh + 4 * q + 2], r[16 * hh + 4 * q + 3]);
              fence_proxy_async();
              __syncwarp();
              if (lane == 0) {
                if (mw < M)
                  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                               ::"l"(reinterpret_cast<uint64_t>(&tma_o)), "r"(slab), "r"(nc + 16 * hh), "r"(mw) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
              }
              ++tma_chunk;
            }
          }
          DDG_LNP(2)
          // ---- the row's partials from all owners, combined in slot order (same result everywhere)
          float mean, rstd;
          {
            float2 pt[8];
            const uint64_t* slots = part;
            const long long t0 = clock64();
            bool all;
            do {
              all = true;
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                if (k < nvalid) {
                  const uint64_t w = ld_relaxed_gpu_b64(slots + k);
                  all = all && ((uint32_t)(w >> 63) == tagbit);
                  pt[k] = make_float2(__uint_as_float((uint32_t)w), __uint_as_float((uint32_t)(w >> 32) & 0x7fffffffu));
                } else {
                  pt[k] = make_float2(0.f, 0.f);
                }
              }
              // ~0.15 s: never hang the device on a stale scratch buffer
            } while (!all && clock64() - t0 < (1ll << 28));
            float sm = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) sm += pt[k].x;
            mean = sm / (float)nvalid;
            float m2 = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k)
              if (k < nvalid) {
                const float d = pt[k].x - mean;
                m2 += pt[k].y + 128.f * d * d;
              }
            rstd = rsqrtf(m2 / (float)p.N + 1e-5f);
          }
          __syncwarp();
          DDG_LNP(3)
          // departure (relaxed): the last owner to leave flips the block's parity for the next launch
          int departed = 0;
          if (lane == 0) departed = atomicAdd(cnt, 1);
          // ---- pass 2
          auto pass2 = [&](auto tag) {
            using TO = decltype(tag);
            if (have_mod) {
              const TO* sh = reinterpret_cast<const TO*>(p.ln_shift) + (int64_t)mrow * p.ln_mod_stride + n0;
              const TO* sc = reinterpret_cast<const TO*>(p.ln_scale) + (int64_t)mrow * p.ln_mod_stride + n0;
              const uint2 rs = __ldg(reinterpret_cast<const uint2*>(sh) + lane);
              const uint2 rg = __ldg(reinterpret_cast<const uint2*>(sc) + lane);
              const float2 s01 = H2<TO>::unpack(rs.x), s23 = H2<TO>::unpack(rs.y);
              const float2 g01 = H2<TO>::unpack(rg.x), g23 = H2<TO>::unpack(rg.y);
              st_shared_v4(lns_s + 16 * lane, __float_as_uint(s01.x), __float_as_uint(s01.y), __float_as_uint(s23.x),
                           __float_as_uint(s23.y));
              st_shared_v4(lng_s + 16 * lane, __float_as_uint(g01.x), __float_as_uint(g01.y), __float_as_uint(g23.x),
                           __float_as_uint(g23.y));
            }
            __syncwarp();
            const float2 mean2 = make_float2(mean, mean), rs2 = make_float2(rstd, rstd);
#pragma unroll 1
            for (int g = 0; g < BN / 64; ++g) {
              const int nc = n0 + g * 32;
              uint32_t r[32];
              tmem_ld32_sync(taddr + (uint32_t)(g * 32), r);
              uint32_t w[16];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const uint4 lw = ld_shared_v4(lnw_s + (uint32_t)(g * 128 + 16 * i));
                // ((x - mean) * rstd) * w, then * (1 + scale) + shift: ln_modulate_kernel's order
                float2 y0 = mul2(mul2(sub2(make_float2(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1])), mean2), rs2),
                                 make_float2(__uint_as_float(lw.x), __uint_as_float(lw.y)));
                float2 y1 = mul2(mul2(sub2(make_float2(__uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3])), mean2), rs2),
                                 make_float2(__uint_as_float(lw.z), __uint_as_float(lw.w)));
                if (have_mod) {
                  const uint4 lg = ld_shared_v4(lng_s + (uint32_t)(g * 128 + 16 * i));
                  const uint4 ls = ld_shared_v4(lns_s + (uint32_t)(g * 128 + 16 * i));
                  y0 = fma2(y0, make_float2(__uint_as_float(lg.x), __uint_as_float(lg.y)),
                            make_float2(__uint_as_float(ls.x), __uint_as_float(ls.y)));
                  y1 = fma2(y1, make_float2(__uint_as_float(lg.z), __uint_as_float(lg.w)),
                            make_float2(__uint_as_float(ls.z), __uint_as_float(ls.w)));
                }
                w[2 * i] = H2<TO>::pack(y0.x, y0.y);
                w[2 * i + 1] = H2<TO>::pack(y1.x, y1.y);
              }
              {
                const uint32_t slab = stg + (tma_chunk & 1u) * TMA_SLAB_BYTES;
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                __syncwarp();
                const uint32_t rowp = slab + (uint32_t)lane * 64u;
#pragma unroll
                for (int q = 0; q < 4; ++q)
                  st_shared_v4(rowp + (((uint32_t)q ^ sw64) << 4), w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                  if (mw < M)
                    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                                 ::"l"(reinterpret_cast<uint64_t>(&tma_o2)), "r"(slab), "r"(nc), "r"(mw) : "memory");
                  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                ++tma_chunk;
              }
            }
          };
          if (p.ln_dtype == DDG_BF16) pass2(__nv_bfloat16());
          else pass2(__half());
          if (lane == 0 && (departed & 255) == target - 1) atomicExch(cnt, (int)(tagbit << 8));
          DDG_LNP(4)
#ifdef DDG_LN_PROF
          if (e == 0 && lane == 0 && blockIdx.x == 2) atomicAdd(&g_ln_prof[5], 1ull);
#endif
        }
      } else if (EPI == DDG_EPI_GATE_RESIDUAL || EPI == DDG_EPI_ACCUM_F32) {
        // ACCUM_F32: out = (accumulate ? out : 0) + acc, no rounding (fp32 dgrad outputs).
        constexpr bool PLAIN = EPI == DDG_EPI_ACCUM_F32;
        // -------- x += round(gate * round(acc + bias)), fp32 residual stream.
        // Bias and (when this warp's 32 rows share one conditioning row, the usual case)
        // gate of the warp's 128 columns are staged once per tile as fp32 in warp-private
        // shared memory; the register phase is packed f32x2 arithmetic.
        const bool have_bias = !PLAIN && p.bias != nullptr;
        const uint32_t bias_s = smem_u32(smem_bias) + (uint32_t)e * (BN / 2) * 4;
        const uint32_t gate_s = smem_u32(smem_gate) + (uint32_t)e * (BN / 2) * 4;
        const T* gate_row = nullptr;
        bool gate_uniform = false;
        if (!PLAIN && p.gate) {
          const int rpm = p.rows_per_mod > 0x7fffffff ? 0x7fffffff : (int)p.rows_per_mod;
          // rows of this warp that exist: [mw, last]; a warp past the end of A uses row 0's gate
          const int first = mw < M ? mw : 0;
          const int last = mw < M ? (mw + 31 < M ? mw + 31 : M - 1) : 0;
          const int mod_first = first / rpm, mod_last = last / rpm;
          gate_uniform = mod_first == mod_last;
          const int mrow = gate_uniform ? mod_first : (row_ok ? m / rpm : 0);
          gate_row = reinterpret_cast<const T*>(p.gate) + (int64_t)mrow * p.gate_stride;
        }
        if (n0 < p.N) {
          if (have_bias) {
            const uint2 raw = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const T*>(p.bias) + n0) + lane);
            const float2 b01 = H2<T>::unpack(raw.x), b23 = H2<T>::unpack(raw.y);
            st_shared_v4(bias_s + 16 * lane, __float_as_uint(b01.x), __float_as_uint(b01.y),
                         __float_as_uint(b23.x), __float_as_uint(b23.y));
          }
          if (gate_uniform) {
            const uint2 raw = __ldg(reinterpret_cast<const uint2*>(gate_row + n0) + lane);
            const float2 g01 = H2<T>::unpack(raw.x), g23 = H2<T>::unpack(raw.y);
            st_shared_v4(gate_s + 16 * lane, __float_as_uint(g01.x), __float_as_uint(g01.y),
                         __float_as_uint(g23.x), __float_as_uint(g23.y));
          }
        }
        __syncwarp();
        float* xslab = reinterpret_cast<float*>(p.out) + (int64_t)mw * p.ld_out + cchunk * 4;
        // the rows the delta is added to: `out` itself (in place) or a separate residual input
        const float* xin = (!PLAIN && p.residual) ? p.residual + (int64_t)mw * p.ld_out + cchunk * 4 : xslab;
        float4 xn[2][4];                         // x of the next 32 columns
        auto prefetch_x = [&](int nc) {
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int row = 8 * i + crow;
              const bool ok = (mw + row < M) && (nc + h * 16 + cchunk * 4 < p.N) && (!PLAIN || p.accumulate);
              xn[h][i] = ok ? *reinterpret_cast<const float4*>(xin + (int64_t)row * p.ld_out + nc + h * 16)
                            : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        prefetch_x(n0);
#pragma unroll 1
        for (int g = 0; g < BN / 64; ++g) {
          const int nc = n0 + g * 32;
          if (nc >= p.N) break;                  // warp-uniform
          uint32_t r0[32];
          tmem_ld32_sync(taddr + (uint32_t)(g * 32), r0);
          float4 xc[2][4];
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int i = 0; i < 4; ++i) xc[h][i] = xn[h][i];
          if (g + 1 < BN / 64 && nc + 32 < p.N) prefetch_x(nc + 32);
          float2 v[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = make_float2(__uint_as_float(r0[2 * i]), __uint_as_float(r0[2 * i + 1]));
          if (have_bias) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const uint4 b = ld_shared_v4(bias_s + (uint32_t)(g * 128 + 16 * i));
              v[2 * i] = add2(v[2 * i], make_float2(__uint_as_float(b.x), __uint_as_float(b.y)));
              v[2 * i + 1] = add2(v[2 * i + 1], make_float2(__uint_as_float(b.z), __uint_as_float(b.w)));
            }
          }
          if (!PLAIN) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = H2<T>::unpack(H2<T>::pack(v[i].x, v[i].y));
          }
          if (gate_uniform) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const uint4 b = ld_shared_v4(gate_s + (uint32_t)(g * 128 + 16 * i));
              v[2 * i] = mul2(v[2 * i], make_float2(__uint_as_float(b.x), __uint_as_float(b.y)));
              v[2 * i + 1] = mul2(v[2 * i + 1], make_float2(__uint_as_float(b.z), __uint_as_float(b.w)));
            }
          } else if (gate_row) {                 // rows of this warp straddle two conditioning rows
            float g0[32];
            load32<T>(gate_row + nc, g0);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = mul2(v[i], make_float2(g0[2 * i], g0[2 * i + 1]));
          }
          if (gate_row) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = H2<T>::unpack(H2<T>::pack(v[i].x, v[i].y));
          }
          float v0[32];
#pragma unroll
          for (int i = 0; i < 16; ++i) { v0[2 * i] = v[i].x; v0[2 * i + 1] = v[i].y; }
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            __syncwarp();
#pragma unroll
            for (int q = 0; q < 4; ++q)
              st_shared_v4(stg_wr + q * 16, __float_as_uint(v0[16 * h + 4 * q]),
                           __float_as_uint(v0[16 * h + 4 * q + 1]), __float_as_uint(v0[16 * h + 4 * q + 2]),
                           __float_as_uint(v0[16 * h + 4 * q + 3]));
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int row = 8 * i + crow;
              const uint4 y = ld_shared_v4(stg_rd + (uint32_t)(8 * i) * EPI_ROW_BYTES);
              if ((mw + row < M) && (nc + h * 16 + cchunk * 4 < p.N)) {
                float4 o = xc[h][i];
                o.x += __uint_as_float(y.x); o.y += __uint_as_float(y.y);
                o.z += __uint_as_float(y.z); o.w += __uint_as_float(y.w);
                *reinterpret_cast<float4*>(xslab + (int64_t)row * p.ld_out + nc + h * 16) = o;
              }
            }
          }
        }
      } else {
        // -------- 16-bit outputs: store / rotary / bias+GELU.  Register phase in packed
        // f32x2 arithmetic on column pairs; the bias of this warp's 128 columns sits as
        // fp32 in warp-private shared memory (one broadcast LDS.128 per 4 columns).
        // a warp's 32 rows lie in one sequence (seq_len % 32 == 0 whenever a map is given)
        const int mw_out = (p.out_seq_map && mw < M)
                             ? __ldg(p.out_seq_map + mw / p.seq_len) * p.seq_len + mw % p.seq_len : mw;
        T* oslab = reinterpret_cast<T*>(p.out) + (int64_t)mw_out * p.ld_out + cchunk * 8;
        const bool have_bias = p.bias != nullptr;
        const uint32_t bias_s = smem_u32(smem_bias) + (uint32_t)e * (BN / 2) * 4;   // shared-window address
        if (have_bias && n0 < p.N) {
          const uint2 raw = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const T*>(p.bias) + n0) + lane);
          const float2 b01 = H2<T>::unpack(raw.x), b23 = H2<T>::unpack(raw.y);
          st_shared_v4(bias_s + 16 * lane, __float_as_uint(b01.x), __float_as_uint(b01.y), __float_as_uint(b23.x),
                       __float_as_uint(b23.y));
        }
        __syncwarp();
        const uint64_t pol_out = policy_evict_first();
        // the 16-bit side tensor (ABI v6): rows mw.., this lane's 8-column chunk
        T* aslab = (EPI == EPI_BIAS_GELU_AUX || EPI == DDG_EPI_GELU_BWD)
                     ? reinterpret_cast<T*>(p.aux) + (int64_t)mw * p.ld_aux + cchunk * 8 : nullptr;
        // GELU_BWD: the pre-activations of this lane's four 16-byte pieces of a 32-column chunk; issued
        // before the chunk's accumulator load so that they travel during it and the transpose
        uint4 pre[4];
        auto load_pre = [&](int nc) {
          const int col = nc + cchunk * 8;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int row = 8 * i + crow;
            pre[i] = (mw + row < M && col + 8 <= p.N)
                       ? __ldcs(reinterpret_cast<const uint4*>(aslab + (int64_t)row * p.ld_aux + nc))
                       : make_uint4(0, 0, 0, 0);
          }
        };
        auto store_chunk = [&](const uint32_t (&w)[16], int nc) {
          if constexpr (tma_out<EPI>()) {
            // thread = row: 64 B of this chunk into the swizzled slab (16-byte piece q of row r at
            // q ^ ((r >> 1) & 3): conflict-free), then ONE bulk tensor store of the 32 x 32 box; columns
            // past N and rows past M are clipped by the tensor map.  The slab written two chunks ago
            // must have been read by its copy before it is overwritten.
            const uint32_t slab = stg + (uint32_t)(tma_chunk % TMA_OUT_BUFS) * TMA_SLAB_BYTES;
            if (lane == 0)
              asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(TMA_OUT_BUFS - 1) : "memory");
            __syncwarp();
            const uint32_t rowp = slab + (uint32_t)lane * 64u;
            const uint32_t sw = (uint32_t)(lane >> 1) & 3u;
#pragma unroll
            for (int q = 0; q < 4; ++q)
              st_shared_v4(rowp + (((uint32_t)q ^ sw) << 4), w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0 && mw < M) {            // (a warp wholly past the live rows sends nothing)
              if (p.stream_out)
                asm volatile(
                  "cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;"
                  ::"l"(reinterpret_cast<uint64_t>(&tma_o)), "r"(slab), "r"(nc), "r"(mw_out), "l"(pol_out) : "memory");
              else
                asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                             ::"l"(reinterpret_cast<uint64_t>(&tma_o)), "r"(slab), "r"(nc), "r"(mw_out) : "memory");
            }
            if (lane == 0) asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            ++tma_chunk;
            return;
          }
          const int col = nc + cchunk * 8;
          __syncwarp();
#pragma unroll
          for (int q = 0; q < 4; ++q) st_shared_v4(stg_wr + q * 16, w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int row = 8 * i + crow;
            uint4 y = ld_shared_v4(stg_rd + (uint32_t)(8 * i) * EPI_ROW_BYTES);
            if (EPI == EPI_BIAS_GELU_AUX) {
              // side output: y is the pre-activation round(acc + bias); the GELU of these 8 values is
              // formed here, lanes along the row, so both tensors leave through ONE trip through the slab
              if (mw + row < M && col + 8 <= p.N) *reinterpret_cast<uint4*>(aslab + (int64_t)row * p.ld_aux + nc) = y;
              const uint32_t pw[4] = {y.x, y.y, y.z, y.w};
              uint32_t o[4];
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const float2 r = gelu_tanh2(H2<T>::unpack(pw[k]));
                o[k] = H2<T>::pack(r.x, r.y);
              }
              y = make_uint4(o[0], o[1], o[2], o[3]);
            }
            if (EPI == DDG_EPI_GELU_BWD) {     // round(g * gelu'(pre)): ddg_gelu_backward's arithmetic
              const uint32_t gw[4] = {y.x, y.y, y.z, y.w}, pw[4] = {pre[i].x, pre[i].y, pre[i].z, pre[i].w};
              uint32_t o[4];
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const float2 r = mul2(H2<T>::unpack(gw[k]), gelu_tanh_grad2(H2<T>::unpack(pw[k])));
                o[k] = H2<T>::pack(r.x, r.y);
              }
              y = make_uint4(o[0], o[1], o[2], o[3]);
            }
            if (mw + row < M) {
              T* dst = oslab + (int64_t)row * p.ld_out + nc;
              if (col + 8 <= p.N) {
                if (p.stream_out) st_global_v4_hint(dst, y, pol_out);
                else *reinterpret_cast<uint4*>(dst) = y;
              } else {
                const T* ye = reinterpret_cast<const T*>(&y);
                for (int k = 0; k < 8; ++k)
                  if (col + k < p.N) dst[k] = ye[k];
              }
            }
          }
        };
        auto add_bias = [&](float2 (&v)[16], int c) {
          if (!have_bias) return;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const uint4 b = ld_shared_v4(bias_s + (uint32_t)(c * 128 + 16 * i));
            v[2 * i] = add2(v[2 * i], make_float2(__uint_as_float(b.x), __uint_as_float(b.y)));
            v[2 * i + 1] = add2(v[2 * i + 1], make_float2(__uint_as_float(b.z), __uint_as_float(b.w)));
          }
        };
        if (EPI == DDG_EPI_QKV_ROPE) {
          // rotary pairs column d with d + 32 of a 64-wide head: two 32-column chunks at a time
          float2 cs[16], sn[16];
          {
            const int pos = row_ok ? (p.rope_pos0 + m % p.seq_len) : 0;
            const uint4* pc = reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(p.rope_cos) + (int64_t)pos * 32);
            const uint4* ps = reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(p.rope_sin) + (int64_t)pos * 32);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const uint4 a = __ldg(pc + i), b = __ldg(ps + i);
              cs[4 * i] = H2<T>::unpack(a.x); cs[4 * i + 1] = H2<T>::unpack(a.y);
              cs[4 * i + 2] = H2<T>::unpack(a.z); cs[4 * i + 3] = H2<T>::unpack(a.w);
              sn[4 * i] = H2<T>::unpack(b.x); sn[4 * i + 1] = H2<T>::unpack(b.y);
              sn[4 * i + 2] = H2<T>::unpack(b.z); sn[4 * i + 3] = H2<T>::unpack(b.w);
            }
          }
#pragma unroll 1
          for (int g = 0; g < BN / 128; ++g) {
            const int nc = n0 + g * 64;
            uint32_t r0[32], r1[32];
            tmem_ld32(taddr + (uint32_t)(g * 64), r0);
            tmem_ld32(taddr + (uint32_t)(g * 64 + 32), r1);
            tmem_ld_wait(r0);
            tmem_ld_wait(r1);
            if (nc >= p.N) continue;                 // warp-uniform
            float2 v0[16], v1[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              v0[i] = make_float2(__uint_as_float(r0[2 * i]), __uint_as_float(r0[2 * i + 1]));
              v1[i] = make_float2(__uint_as_float(r1[2 * i]), __uint_as_float(r1[2 * i + 1]));
            }
            add_bias(v0, 2 * g);
            add_bias(v1, 2 * g + 1);
            uint32_t w0[16], w1[16];
            if (nc < p.rope_cols) {
              // the GEMM output is a 16-bit tensor in the reference; the rotary then runs
              // in fp32 on those values (NeoX: x1 = d < 32, x2 = d >= 32)
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const float2 x1 = H2<T>::unpack(H2<T>::pack(v0[i].x, v0[i].y));
                const float2 x2 = H2<T>::unpack(H2<T>::pack(v1[i].x, v1[i].y));
                const float2 o0 = sub2(mul2(x1, cs[i]), mul2(x2, sn[i]));
                const float2 o1 = fma2(x1, sn[i], mul2(x2, cs[i]));
                w0[i] = H2<T>::pack(o0.x, o0.y);
                w1[i] = H2<T>::pack(o1.x, o1.y);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                w0[i] = H2<T>::pack(v0[i].x, v0[i].y);
                w1[i] = H2<T>::pack(v1[i].x, v1[i].y);
              }
            }
            store_chunk(w0, nc);
            store_chunk(w1, nc + 32);
          }
        } else {
          // 32-column chunks; the TMEM load of the next chunk is in flight while this one is processed
          const bool want_stats = EPI == DDG_EPI_STORE && p.stats != nullptr;
          float st_m = -INFINITY, st_s = 0.f;      // running (max, sum 2^((v-max) log2e)) of this row chunk
#pragma unroll 1
          for (int c = 0; c < BN / 64; ++c) {
            const int nc = n0 + c * 32;
            if (nc >= p.N) break;                    // warp-uniform
            if (EPI == DDG_EPI_GELU_BWD) load_pre(nc);
            uint32_t r[32];
            tmem_ld32_sync(taddr + (uint32_t)(c * 32), r);
            float2 v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = make_float2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
            add_bias(v, c);
            uint32_t w[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) w[i] = H2<T>::pack(v[i].x, v[i].y);
            if (EPI == DDG_EPI_BIAS_GELU) {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const float2 y = gelu_tanh2(H2<T>::unpack(w[i]));
                w[i] = H2<T>::pack(y.x, y.y);
              }
            } else if (want_stats) {
              // log-softmax partials of the stored (rounded) logits, so that the
              // fused sampler never needs its own normaliser pass over the row
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = H2<T>::unpack(w[i]);
              const bool plain = (nc + 32 <= p.N) && !(p.mask_col >= nc && p.mask_col < nc + 32);
              if (!plain) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  if (nc + 2 * i >= p.N || nc + 2 * i == p.mask_col) v[i].x = -INFINITY;
                  if (nc + 2 * i + 1 >= p.N || nc + 2 * i + 1 == p.mask_col) v[i].y = -INFINITY;
                }
              }
              float gm = -INFINITY;
#pragma unroll
              for (int i = 0; i < 16; ++i) gm = fmaxf(gm, fmaxf(v[i].x, v[i].y));
              const float nm = fmaxf(st_m, gm);
              if (nm > -INFINITY) {
                const float nml = nm * kLog2e;
                st_s *= ex2f_approx(st_m * kLog2e - nml);
                float2 acc = make_float2(0.f, 0.f);
                const float2 l2 = make_float2(kLog2e, kLog2e), off = make_float2(-nml, -nml);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  const float2 t = fma2(v[i], l2, off);
                  acc = add2(acc, make_float2(ex2f_approx(t.x), ex2f_approx(t.y)));
                }
                st_s += acc.x + acc.y;
                st_m = nm;
              }
            }
            store_chunk(w, nc);
          }
          if (want_stats && row_ok && n0 < p.N)
            p.stats[(int64_t)(n0 / (BN / 2)) * p.stats_ld + (mw_out + (m - mw))] = make_float2(st_m, st_s);
        }
      }
      // this warp has drained its part of the accumulator: one arrival per warp
      // on the leader's barrier (the leader's MMA thread reuses both CTAs' TMEM)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tmem_empty_leader0 + (uint32_t)acc * 8);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    // TMA-store epilogues: this warp's copies have left shared memory and are complete
    if ((tma_out<EPI>() || EPI == EPI_GATE_RESIDUAL_LN) && lane == 0)
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tc_fence_before();
  cluster_sync_all();     // neither CTA may exit while its pair can still signal its barriers
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// 2-D K-major tensor map: dims {K, rows}, box {64, box_rows}, 128B swizzle, zero OOB fill.
int make_tma_2d(CUtensorMap* map, const void* base, int64_t rows, int64_t k, int64_t ld, int box_rows, int dtype) {
  EncodeTiledFn fn = get_encode_fn();
  DDG_REQUIRE(fn != nullptr, DDG_ERR_CUDA, "cuTensorMapEncodeTiled unavailable");
  cuuint64_t gdim[2] = {(cuuint64_t)k, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, dtype == DDG_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                  const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DDG_REQUIRE(r == CUDA_SUCCESS, DDG_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d): rows=%lld k=%lld ld=%lld", (int)r,
              (long long)rows, (long long)k, (long long)ld);
  return DDG_OK;
}

}  // namespace

// 16-bit row-major matrix as a 2-D tensor map {cols, rows} with a {64 columns, box_rows} box and
// the 128-byte swizzle (16-byte chunk index XOR row % 8); shared with attention.cu
int make_tma_rows64(void* map, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows, int dtype) {
  return make_tma_2d(reinterpret_cast<CUtensorMap*>(map), base, rows, cols, ld, box_rows, dtype);
}

namespace {

// 16-bit row-major output as a 2-D tensor map {cols, rows} with a {32 columns, 32 rows} box (64 B x 32) and
// the 64-byte swizzle: the destination of the TMA-store epilogues (clips columns >= N and rows >= M)
// dtype DDG_F32: fp32 rows, {16 columns, 32 rows} box (also 64 B x 32): the fused-LayerNorm epilogue's x_new
int make_tma_out(CUtensorMap* map, void* base, int64_t rows, int64_t cols, int64_t ld, int dtype) {
  EncodeTiledFn fn = get_encode_fn();
  DDG_REQUIRE(fn != nullptr, DDG_ERR_CUDA, "cuTensorMapEncodeTiled unavailable");
  const bool f32 = dtype == DDG_F32;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * (f32 ? 4 : 2)};
  cuuint32_t box[2] = {f32 ? 16u : 32u, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                        : dtype == DDG_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base,
                  gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                  CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DDG_REQUIRE(r == CUDA_SUCCESS, DDG_ERR_CUDA, "cuTensorMapEncodeTiled (output) failed (%d): rows=%lld cols=%lld ld=%lld",
              (int)r, (long long)rows, (long long)cols, (long long)ld);
  return DDG_OK;
}

template <typename T, int EPI>
int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t s) {
  static bool attr_done[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_done[dev]) {
    DDG_CUDA(cudaFuncSetAttribute(gemm_kernel<T, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  smem_bytes<EPI>()));
    attr_done[dev] = true;
  }
  const int num_tiles = ((p.M + 2 * BM - 1) / (2 * BM)) * ((p.N + BN - 1) / BN);
  int grid = sm_count() & ~1;           // CTA pairs
  if (grid > 2 * num_tiles) grid = 2 * num_tiles;
  CUtensorMap to = ta;                  // only the TMA-store epilogues read it
  if constexpr (tma_out<EPI>()) {
    const int rc = make_tma_out(&to, p.out, p.out_rows, p.N, p.ld_out, std::is_same<T, __half>::value ? DDG_F16 : DDG_BF16);
    if (rc) return rc;
  }
  CUtensorMap to2 = ta;                 // fused LayerNorm: the normalised 16-bit rows
  if constexpr (EPI == EPI_GATE_RESIDUAL_LN) {
    int rc = make_tma_out(&to, p.out, p.M, p.N, p.ld_out, DDG_F32);
    if (rc) return rc;
    rc = make_tma_out(&to2, p.ln_out, p.M, p.N, p.N, p.ln_dtype);
    if (rc) return rc;
  }
  gemm_kernel<T, EPI><<<grid, NUM_THREADS, smem_bytes<EPI>(), s>>>(ta, tb, to, to2, p);
  return check_cuda(cudaGetLastError(), "gemm launch");
}

template <typename T>
int dispatch_epi(int epi, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t s) {
  switch (epi) {
    case DDG_EPI_STORE: return launch_gemm<T, DDG_EPI_STORE>(ta, tb, p, s);
    case DDG_EPI_QKV_ROPE: return launch_gemm<T, DDG_EPI_QKV_ROPE>(ta, tb, p, s);
    case DDG_EPI_BIAS_GELU:
      if (p.aux) return launch_gemm<T, EPI_BIAS_GELU_AUX>(ta, tb, p, s);
      return launch_gemm<T, DDG_EPI_BIAS_GELU>(ta, tb, p, s);
    case DDG_EPI_GATE_RESIDUAL:
      if (p.ln_w) return launch_gemm<T, EPI_GATE_RESIDUAL_LN>(ta, tb, p, s);
      return launch_gemm<T, DDG_EPI_GATE_RESIDUAL>(ta, tb, p, s);
    case DDG_EPI_ACCUM_F32: return launch_gemm<T, DDG_EPI_ACCUM_F32>(ta, tb, p, s);
    case DDG_EPI_GELU_BWD: return launch_gemm<T, DDG_EPI_GELU_BWD>(ta, tb, p, s);
  }
  set_error("gemm: bad epilogue %d", epi);
  return DDG_ERR_INVALID;
}

}  // namespace
}  // namespace ddg

#ifdef DDG_LN_PROF
// out[0..4]: cycles waiting for the accumulator / pass 1 / publish + residual store / polling the
// neighbours' partials / pass 2; [5] tiles; [6] cycles the MMA thread waited for a free accumulator
extern "C" __attribute__((visibility("default"))) int ddg_debug_ln_prof(unsigned long long* out, int reset) {
  if (out) cudaMemcpyFromSymbol(out, ddg::g_ln_prof, sizeof(unsigned long long) * 8);
  if (reset) {
    unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    cudaMemcpyToSymbol(ddg::g_ln_prof, z, sizeof(z));
  }
  return 0;
}
#endif

extern "C" int64_t ddg_gemm_ln_work_bytes(int32_t M, int32_t N) {
  if (M < 0 || N <= 0) return 0;
  // per 128-row block: a 64-byte counter slot + 128 rows x ceil(N / 128) words of 8 bytes
  const int64_t n_blocks = 2 * (((int64_t)M + 2 * ddg::BM - 1) / (2 * ddg::BM));
  return n_blocks * (64 + (int64_t)ddg::BM * ((N + 127) / 128) * 8);
}

extern "C" int ddg_gemm(const ddg_gemm_args* args, ddg_stream_t stream) {
  using namespace ddg;
  DDG_REQUIRE(args != nullptr, DDG_ERR_INVALID, "gemm: null args");
  const ddg_gemm_args& a = *args;
  int rc = ddg_device_check();
  if (rc) return rc;
  DDG_REQUIRE(a.M >= 0 && a.N > 0 && a.K > 0, DDG_ERR_INVALID, "gemm: bad sizes M=%d N=%d K=%d", a.M, a.N, a.K);
  if (a.M == 0) return DDG_OK;
  DDG_REQUIRE(a.a && a.w && a.out, DDG_ERR_INVALID, "gemm: null a/w/out");
  DDG_REQUIRE(a.dtype == DDG_F16 || a.dtype == DDG_BF16, DDG_ERR_INVALID, "gemm: dtype must be f16 or bf16");
  DDG_REQUIRE(a.K % 8 == 0 && a.lda % 8 == 0 && a.ldw % 8 == 0 && a.lda >= a.K && a.ldw >= a.K,
              DDG_ERR_INVALID, "gemm: K, lda, ldw must be multiples of 8 (TMA 16-byte strides)");
  DDG_REQUIRE(reinterpret_cast<uintptr_t>(a.a) % 16 == 0 && reinterpret_cast<uintptr_t>(a.w) % 16 == 0 &&
              reinterpret_cast<uintptr_t>(a.out) % 16 == 0, DDG_ERR_INVALID, "gemm: a/w/out must be 16-byte aligned");
  DDG_REQUIRE(a.ld_out >= a.N, DDG_ERR_INVALID, "gemm: ld_out < N");
  if (a.epilogue == DDG_EPI_GATE_RESIDUAL)
    DDG_REQUIRE(a.ld_out % 4 == 0 && a.N % 4 == 0, DDG_ERR_INVALID,
                "gemm: fp32 residual N and ld_out must be multiples of 4");
  else if (a.epilogue == DDG_EPI_ACCUM_F32)   // columns up to the next multiple of 4 may be written
    DDG_REQUIRE(a.ld_out % 4 == 0 && a.ld_out >= (a.N + 3) / 4 * 4, DDG_ERR_INVALID,
                "gemm: fp32 output ld_out must be a multiple of 4 and cover N rounded up to 4");
  else
    DDG_REQUIRE(a.ld_out % 8 == 0, DDG_ERR_INVALID, "gemm: 16-bit ld_out must be a multiple of 8");
  if (a.bias) DDG_REQUIRE(reinterpret_cast<uintptr_t>(a.bias) % 16 == 0, DDG_ERR_INVALID, "gemm: bias alignment");
  if (a.epilogue == DDG_EPI_QKV_ROPE) {
    DDG_REQUIRE(a.rope_cos && a.rope_sin && a.seq_len > 0 && a.head_dim == 64 && a.rope_cols % 64 == 0,
                DDG_ERR_UNSUPPORTED, "gemm: rotary epilogue needs head_dim 64 and cos/sin tables");
  }
  if (a.epilogue == DDG_EPI_GATE_RESIDUAL && a.gate)
    DDG_REQUIRE(a.rows_per_mod > 0 && a.gate_stride % 8 == 0 &&
                reinterpret_cast<uintptr_t>(a.gate) % 16 == 0, DDG_ERR_INVALID, "gemm: gate layout");

  CUtensorMap ta, tb;
  rc = make_tma_2d(&ta, a.a, a.M, a.K, a.lda, BM, a.dtype);
  if (rc) return rc;
  rc = make_tma_2d(&tb, a.w, a.N, a.K, a.ldw, BNH, a.dtype);
  if (rc) return rc;
  GemmParams p;
  p.out = a.out; p.bias = a.bias; p.gate = a.gate; p.rope_cos = a.rope_cos; p.rope_sin = a.rope_sin;
  p.ld_out = a.ld_out; p.gate_stride = a.gate_stride; p.rows_per_mod = a.rows_per_mod;
  p.M = a.M; p.N = a.N; p.K = a.K; p.seq_len = a.seq_len; p.rope_cols = a.rope_cols;
  p.stats = reinterpret_cast<float2*>(a.stats); p.stats_ld = a.stats_ld; p.mask_col = a.stats_mask_col;
  p.accumulate = a.accumulate;
  p.rope_pos0 = a.rope_pos0;
  p.live_seqs = a.live_seqs;
  p.out_seq_map = a.out_seq_map;
  if (a.live_seqs || a.out_seq_map)
    DDG_REQUIRE(a.seq_len > 0 && a.M % a.seq_len == 0, DDG_ERR_INVALID,
                "gemm: live_seqs / out_seq_map need seq_len > 0 dividing M");
  if (a.out_seq_map)
    DDG_REQUIRE(a.epilogue == DDG_EPI_STORE && a.seq_len % 32 == 0, DDG_ERR_INVALID,
                "gemm: out_seq_map needs the STORE epilogue and seq_len %% 32 == 0");
  // a 16-bit output that cannot stay in the 126 MB L2 anyway (the vocab head: 8 GB of logits at
  // 131072 rows) is written with an evict-first policy and its operands are loaded evict-last:
  // without this the write-allocate stream kept evicting W (47 MB) and ncu showed 4.4 GB of DRAM
  // reads per launch against 0.25 GB algorithmic (profiles/r01_traffic.json)
  p.ln_w = nullptr; p.ln_shift = p.ln_scale = nullptr; p.ln_out = nullptr; p.ln_work = nullptr;
  p.residual = nullptr;
  if (a.residual) {
    DDG_REQUIRE(a.epilogue == DDG_EPI_GATE_RESIDUAL && reinterpret_cast<uintptr_t>(a.residual) % 16 == 0,
                DDG_ERR_INVALID, "gemm: residual needs the GATE_RESIDUAL epilogue and 16-byte alignment");
    p.residual = a.residual;
  }
  p.ln_mod_stride = 0; p.ln_dtype = 0;
  if (a.ln_weight) {
    // fused LayerNorm + modulate of the updated residual rows (see the epilogue)
    DDG_REQUIRE(a.epilogue == DDG_EPI_GATE_RESIDUAL, DDG_ERR_INVALID, "gemm: ln_weight needs the GATE_RESIDUAL epilogue");
    DDG_REQUIRE(a.N % 128 == 0 && a.N <= 1024 && a.ld_out == a.N, DDG_ERR_UNSUPPORTED,
                "gemm: fused LayerNorm needs N %% 128 == 0, N <= 1024 and a dense residual (N=%d ld_out=%lld)", a.N,
                (long long)a.ld_out);
    DDG_REQUIRE(a.ln_out && a.ln_work && reinterpret_cast<uintptr_t>(a.ln_out) % 16 == 0 &&
                reinterpret_cast<uintptr_t>(a.ln_work) % 256 == 0 && reinterpret_cast<uintptr_t>(a.ln_weight) % 16 == 0,
                DDG_ERR_INVALID, "gemm: fused LayerNorm needs ln_out (16-byte aligned) and ln_work (256-byte aligned)");
    DDG_REQUIRE(a.ln_dtype == DDG_F16 || a.ln_dtype == DDG_BF16, DDG_ERR_INVALID, "gemm: ln_dtype must be f16 or bf16");
    DDG_REQUIRE((a.ln_shift == nullptr) == (a.ln_scale == nullptr), DDG_ERR_INVALID,
                "gemm: ln_shift and ln_scale go together");
    if (a.gate || a.ln_shift)
      DDG_REQUIRE(a.rows_per_mod > 0 && a.rows_per_mod % 32 == 0, DDG_ERR_UNSUPPORTED,
                  "gemm: fused LayerNorm needs rows_per_mod %% 32 == 0");
    if (a.ln_shift)
      DDG_REQUIRE(a.ln_mod_stride % 8 == 0 && reinterpret_cast<uintptr_t>(a.ln_shift) % 16 == 0 &&
                  reinterpret_cast<uintptr_t>(a.ln_scale) % 16 == 0, DDG_ERR_INVALID, "gemm: ln_shift / ln_scale layout");
    p.ln_w = a.ln_weight; p.ln_shift = a.ln_shift; p.ln_scale = a.ln_scale; p.ln_out = a.ln_out;
    p.ln_mod_stride = a.ln_mod_stride; p.ln_dtype = a.ln_dtype;
    p.ln_work = reinterpret_cast<uint8_t*>(a.ln_work);
  }
  p.aux = nullptr; p.ld_aux = 0;
  // with a sequence scatter map the output rows are those of the full batch: M is its upper bound too
  p.out_rows = a.M;
  if (a.aux || a.epilogue == DDG_EPI_GELU_BWD) {
    DDG_REQUIRE(a.epilogue == DDG_EPI_BIAS_GELU || a.epilogue == DDG_EPI_GELU_BWD, DDG_ERR_INVALID,
                "gemm: aux goes with the BIAS_GELU (output) and GELU_BWD (input) epilogues");
    DDG_REQUIRE(a.aux && a.ld_aux >= a.N && a.ld_aux % 8 == 0 && a.N % 8 == 0 &&
                reinterpret_cast<uintptr_t>(a.aux) % 16 == 0, DDG_ERR_INVALID,
                "gemm: aux needs N %% 8 == 0, ld_aux >= N, ld_aux %% 8 == 0 and 16-byte alignment");
    p.aux = a.aux; p.ld_aux = a.ld_aux;
  }
  p.stream_out = (a.epilogue == DDG_EPI_STORE && (int64_t)a.M * a.N * 2 >= ((int64_t)256 << 20)) ? 1 : 0;
  if (a.stats) {
    DDG_REQUIRE(a.epilogue == DDG_EPI_STORE && a.stats_ld >= a.M &&
                reinterpret_cast<uintptr_t>(a.stats) % 8 == 0, DDG_ERR_INVALID,
                "gemm: softmax partials need the STORE epilogue and stats_ld >= M");
  }
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (a.dtype == DDG_F16) return dispatch_epi<__half>(a.epilogue, ta, tb, p, s);
  return dispatch_epi<__nv_bfloat16>(a.epilogue, ta, tb, p, s);
}
