// Hierarchical exponential race: the production sampler (in-kernel Philox).
//
// Same per-token mathematics as sample_step.cu / sample_fused.cu (reference
// diffusion.py:356-394, 1393-1420, 1456-1509) up to the guided, unnormalised
// posterior q~_j.  The categorical draw (reference diffusion.py:1748-1755,
// argmax_j q_j / E_j with one exponential E_j per vocabulary entry) is
// restructured with the two textbook properties of an exponential race:
//   min_j E_j / q_j  over a set G  ~  Exp(1) / sum_{j in G} q_j,   and
//   P(argmin = j | G wins) = q_j / sum_{j in G} q_j, independent of the minimum.
// So the row is cut into groups of 8 consecutive columns (one 16-byte vector
// of bf16 logits); each group gets ONE exponential, the winning group is the
// arg-max of  log2(sum_G q~) - log2(E_G),  and the column inside the winning
// group is drawn by inverse CDF from one more uniform.  The distribution of
// the result is exactly the reference's; the Philox work and the two lg2 per
// column of the noise drop by 8x, which is what moves this kernel from the
// MUFU / issue limit towards the HBM roofline.  (With injected noise the
// per-column race of sample_fused.cu / sample_step.cu is used: that is the
// path that is bit-comparable with the reference.)
//
// round_bf16 = 1 (the reference's bf16 log-softmax / exp roundings, SURVEY
// 8a' item 4): lp = bf16(l - lse) takes only ~2000 distinct values that matter,
// so the per-column factor of q~ is a 2048-entry shared-memory table per
// forward (built per CTA, rebuilt when the schedule scalars change):
//   linear tables  A[i] = (dA x_i + c0)^gamma / max,  B[i] = (..)^(1-gamma) / max
//                  -> q~_j = A[ia_j] * B[ib_j], group sum by FFMA, no MUFU per column
//                  (used when the dynamic range provably fits fp32),
//   log2 tables    otherwise (q~_j = 2^(A + B), one ex2 per column).
//
// One CTA of 128 threads walks one row at a time; thread t owns the groups
// G = t + 128 k.  The next row's ids and softmax partials are prefetched
// (register / cp.async) while the current row streams.
//
// Philox stream (so tests can regenerate it; G = j / 8 is the group index):
//   key     = (seed_lo, seed_hi)
//   group exponential: counter = ((G & 255) | ((G >> 10) << 8), row_lo, row_hi, step),
//                      word (G >> 8) & 3,  U = (word >> 8) * 2^-24,
//                      E = 1e-10 - ln(U + 1e-10)
//   in-group pick:     counter = (G, row_lo, row_hi | 0x80000000, step), word 0
//
// race_group = 64 (WIDE): a group is any SET of columns, it need not be contiguous.  Here
// it is the 8 vectors one thread touches in one sweep of its unrolled loop: with vector
// index v = j / 8, slot t = v % 128, kk = v / 128, sweep k8 = kk / 8, the group is
// G = k8 * 128 + t and holds columns 8 (t + 128 (8 k8 + s)) + e, s, e = 0..7 (64 columns,
// pick order (s, e)).  Memory access is unchanged; the per-group work (Philox, the two lg2
// of the exponential, the lg2 of the sum, the compare) is paid once per 64 columns instead
// of once per 8, which removes a third of the kernel's instructions.  The in-group pick is
// done by 64 threads (one column each) with a shuffle scan.
//   group exponential: counter = (t | ((k8 >> 2) << 8), row_lo, row_hi | 0x40000000, step),
//                      word k8 & 3
//   in-group pick:     counter = (G, row_lo, row_hi | 0x80000000, step), word 0
#include "ddg_common.cuh"

namespace ddg {
namespace {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr int kThreads = 128;
constexpr int kWarps = kThreads / 32;
constexpr int kCtasPerSm = 6;
// resident CTAs per SM of the 64-column-group kernels (measured: 72 registers / 7 CTAs is the
// best point for the two-forward D-CFG kernel, 64 / 8 for the one-forward ones)
__host__ __device__ constexpr int wide_ctas(int mode) { return mode == 1 /* G_UNI_CFG */ ? 7 : 8; }
constexpr int kLut = 2048;          // bf16 magnitudes 0x3B00 .. 0x42FF: |lp| in [2^-9, 128)
constexpr int kLutBase = 0x3B00;
constexpr int kMaxChunks = 256;     // softmax partials staged in smem (V <= 32768); else read from global
constexpr float kLinearRange = 120.f;   // log2 dynamic range allowed for the linear tables

enum GMode : int { G_UNI_NONE = 0, G_UNI_CFG = 1, G_ABS_NONE = 2 };

__device__ __forceinline__ float ex2a(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2a(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint4 ldg_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
               "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ float neg_log2_E(float u) {
  // E = 1e-10 - ln(u + 1e-10);  returns -log2(E)
  const float t = lg2a(u + 1e-10f);
  return -lg2a(fmaf(-kLn2, t, 1e-10f));
}
__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

struct RowConst {
  float Ka, Kb;            // natural-log lse of the two rows
  float ka2, kb2;          // the same in base 2
  float dA, c0;            // uniform: q~_j = dA x_j + c0 (+ zc at j = z)
  float zc_a, zc_b;
  float gam, omg;          // gamma, 1 - gamma
  float mask_score;        // absorbing: log2(r / (1 - r)) of the mask column
  int z, mask_index, V;
};

// x = exp(lp) with the reference's roundings when ROUND
template <bool ROUND>
__device__ __forceinline__ float x_of(float l, float K, float k2) {
  if (ROUND) {
    const float lp = bf16_round(l - K);
    return bf16_round(ex2a(lp * kLog2e));
  }
  return ex2a(fmaf(l, kLog2e, -k2));
}

// log2 of the unnormalised guided posterior of column j: every special case
// (j == z, the absorbing mask column, columns past the vocabulary).  Used for
// the group that holds z / the mask / the ragged tail and for the in-group
// pick; the hot loop has its own branch-free versions of the ordinary case.
template <int MODE, bool ROUND>
__device__ __forceinline__ float g_elem(const RowConst& c, float la, float lb, int j) {
  if (j >= c.V) return -INFINITY;
  if (MODE == G_ABS_NONE) {
    if (j == c.mask_index) return c.mask_score;
    if (!ROUND) return fmaf(la, kLog2e, -c.ka2);
    return lg2a(x_of<true>(la, c.Ka, c.ka2));
  }
  float na = fmaf(c.dA, x_of<ROUND>(la, c.Ka, c.ka2), c.c0);
  if (j == c.z) na += c.zc_a;
  if (MODE == G_UNI_NONE) return lg2a(na);
  float nb = fmaf(c.dA, x_of<ROUND>(lb, c.Kb, c.kb2), c.c0);
  if (j == c.z) nb += c.zc_b;
  return __fadd_rn(__fmul_rn(c.gam, lg2a(na)), __fmul_rn(c.omg, lg2a(nb)));
}

template <int MODE, bool ROUND>
__device__ __forceinline__ void group_careful(const RowConst& c, const uint4& ra, const uint4& rb, int j0,
                                              float (&g)[8]) {
  const uint32_t wa[4] = {ra.x, ra.y, ra.z, ra.w};
  const uint32_t wb[4] = {rb.x, rb.y, rb.z, rb.w};
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    g[2 * p] = g_elem<MODE, ROUND>(c, bf16_lo(wa[p]), bf16_lo(wb[p]), j0 + 2 * p);
    g[2 * p + 1] = g_elem<MODE, ROUND>(c, bf16_hi(wa[p]), bf16_hi(wb[p]), j0 + 2 * p + 1);
  }
}

// packed f32x2 arithmetic (one issue slot for two results)
__device__ __forceinline__ float2 sub2(float k, float l0, float l1) {
  float2 r;
  asm("{\n\t"
      ".reg .b64 a, b, d;\n\t"
      "mov.b64 a, {%2, %2};\n\t"
      "mov.b64 b, {%3, %4};\n\t"
      "sub.f32x2 d, a, b;\n\t"
      "mov.b64 {%0, %1}, d;\n\t"
      "}" : "=f"(r.x), "=f"(r.y) : "f"(k), "f"(l0), "f"(l1));
  return r;
}
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

// Table byte offsets of a pair of differences d = lse - l >= 0: bf16 round,
// (magnitude bits - 0x3B00) clamped to [0, 2047] with one DPX instruction for
// the pair, times 4 (both halves at once: 4 * 2047 still fits 16 bits).
constexpr uint32_t kNegBase2 = (uint32_t)(0x10000 - kLutBase) * 0x00010001u;   // (-0x3B00, -0x3B00) as s16x2
__device__ __forceinline__ uint32_t lut_offsets2(float d0, float d1, uint32_t neg_base2) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(d0, d1);
  const uint32_t w = *reinterpret_cast<const uint32_t*>(&h);
  constexpr uint32_t kMaxIdx = (uint32_t)(kLut - 1) * 0x00010001u;
  return __viaddmin_s16x2_relu(w, neg_base2, kMaxIdx) << 2;
}
__device__ __forceinline__ float lut_at(const float* T, uint32_t byte_off) {
  return *reinterpret_cast<const float*>(reinterpret_cast<const char*>(T) + byte_off);
}
// table values of the first NP column pairs of one 16-byte vector
template <int NP = 4>
__device__ __forceinline__ void lut_gather(const uint4& r, float K, const float* T, uint32_t neg_base2,
                                           float2 (&v)[4]) {
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    const float2 d = sub2(K, bf16_lo(w[p]), bf16_hi(w[p]));
    const uint32_t t = lut_offsets2(d.x, d.y, neg_base2);
    v[p] = make_float2(lut_at(T, t & 0xffffu), lut_at(T, t >> 16));
  }
}

__device__ __forceinline__ float max8(const float (&g)[8]) {
  return fmaxf(fmaxf(fmaxf(g[0], g[1]), fmaxf(g[2], g[3])), fmaxf(fmaxf(g[4], g[5]), fmaxf(g[6], g[7])));
}
// sum_e 2^(g_e - gm) with packed f32x2 subtracts / adds
__device__ __forceinline__ float sum_ex2_8(const float (&g)[8], float gm) {
  float2 acc = make_float2(0.f, 0.f);
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const float2 d = add2(make_float2(g[2 * p], g[2 * p + 1]), make_float2(-gm, -gm));
    acc = add2(acc, make_float2(ex2a(d.x), ex2a(d.y)));
  }
  return acc.x + acc.y;
}

// log2( sum over the 8 ordinary columns of one group of q~_j ) up to the
// row-constant `goff` (no z, no mask column, all < V): branch-free.
//   LIN  (ROUND only): tables hold linear factors, sum by FFMA
//   !LIN: tables (ROUND) or MUFU math give log2 q~_j, sum of 2^(g - max)
template <int MODE, bool ROUND, bool LIN>
__device__ __forceinline__ float group_hot(const RowConst& c, const uint4& ra, const uint4& rb, const float* A,
                                           const float* B, uint32_t neg_base2) {
  if (ROUND && LIN) {
    float2 va[4];
    lut_gather(ra, c.Ka, A, neg_base2, va);
    float2 acc;
    if (MODE == G_UNI_CFG) {
      float2 vb[4];
      lut_gather(rb, c.Kb, B, neg_base2, vb);
      acc = make_float2(va[0].x * vb[0].x, va[0].y * vb[0].y);
#pragma unroll
      for (int p = 1; p < 4; ++p) acc = fma2(va[p], vb[p], acc);
    } else {
      acc = add2(add2(va[0], va[1]), add2(va[2], va[3]));
    }
    return lg2a(acc.x + acc.y);
  }
  float g[8];
  if (ROUND) {
    float2 va[4];
    lut_gather(ra, c.Ka, A, neg_base2, va);
    if (MODE == G_UNI_CFG) {
      float2 vb[4];
      lut_gather(rb, c.Kb, B, neg_base2, vb);
#pragma unroll
      for (int p = 0; p < 4; ++p) va[p] = add2(va[p], vb[p]);
    }
#pragma unroll
    for (int p = 0; p < 4; ++p) { g[2 * p] = va[p].x; g[2 * p + 1] = va[p].y; }
  } else {
    const uint32_t wa[4] = {ra.x, ra.y, ra.z, ra.w};
    const uint32_t wb[4] = {rb.x, rb.y, rb.z, rb.w};
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      if (MODE == G_ABS_NONE) {
        g[2 * p] = fmaf(bf16_lo(wa[p]), kLog2e, -c.ka2);
        g[2 * p + 1] = fmaf(bf16_hi(wa[p]), kLog2e, -c.ka2);
      } else {
        float g0 = lg2a(fmaf(c.dA, ex2a(fmaf(bf16_lo(wa[p]), kLog2e, -c.ka2)), c.c0));
        float g1 = lg2a(fmaf(c.dA, ex2a(fmaf(bf16_hi(wa[p]), kLog2e, -c.ka2)), c.c0));
        if (MODE == G_UNI_CFG) {
          const float h0 = lg2a(fmaf(c.dA, ex2a(fmaf(bf16_lo(wb[p]), kLog2e, -c.kb2)), c.c0));
          const float h1 = lg2a(fmaf(c.dA, ex2a(fmaf(bf16_hi(wb[p]), kLog2e, -c.kb2)), c.c0));
          g0 = __fadd_rn(__fmul_rn(c.gam, g0), __fmul_rn(c.omg, h0));
          g1 = __fadd_rn(__fmul_rn(c.gam, g1), __fmul_rn(c.omg, h1));
        }
        g[2 * p] = g0;
        g[2 * p + 1] = g1;
      }
    }
  }
  const float gm = max8(g);
  return gm + lg2a(sum_ex2_8(g, gm));      // all -inf -> NaN, which never wins a '>' comparison
}

// group exponential words of (virtual thread vt, quad q): see the stream layout above
__device__ __forceinline__ uint4 group_noise_words(int vt, int quad, int64_t row, int step, uint2 key) {
  return philox4x32_10(
    make_uint4((uint32_t)(vt | (quad << 8)), (uint32_t)row, (uint32_t)((uint64_t)row >> 32), (uint32_t)step), key);
}
__device__ __forceinline__ uint32_t group_noise_word(int G, int64_t row, int step, uint2 key) {
  const uint4 r = group_noise_words(G & 255, G >> 10, row, step, key);
  const int w = (G >> 8) & 3;
  return w == 0 ? r.x : w == 1 ? r.y : w == 2 ? r.z : r.w;
}

// One pass over the full groups of a row (va / vb: the row's vectors, CTA-uniform):
// thread t takes G = t + 128 k.  Two
// vector buffers alternate (even / odd k), each reloaded for k + 2 as soon as it
// has been consumed, so every load has two group iterations to land.
template <int MODE, bool ROUND, bool LIN>
__device__ __forceinline__ void row_pass(const RowConst& c, const uint4* va, const uint4* vb, uint4 ra0, uint4 rb0,
                                         uint4 ra1, uint4 rb1, const float* A, const float* B, uint32_t neg_base2,
                                         int n_full, int special_grp, int64_t row, int step, uint2 key, float& best,
                                         int& bgrp) {
  const int tid = threadIdx.x;
  uint4 ra[2] = {ra0, ra1}, rb[2] = {rb0, rb1};
  // 8 consecutive k of this thread = virtual threads (tid, tid + 128) x one Philox quad each
  for (int k8 = 0; tid + k8 * (8 * kThreads) < n_full; ++k8) {
    const uint4 n0 = group_noise_words(tid, k8, row, step, key);
    const uint4 n1 = group_noise_words(tid + kThreads, k8, row, step, key);
    const uint32_t rw[8] = {n0.x, n1.x, n0.y, n1.y, n0.z, n1.z, n0.w, n1.w};
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      const int G = tid + (k8 * 8 + s) * kThreads;
      if (G >= n_full) break;
      const float lsum = group_hot<MODE, ROUND, LIN>(c, ra[s & 1], rb[s & 1], A, B, neg_base2);
      {                                               // reload this buffer for k + 2 (clamped: always in bounds)
        const uint32_t off = (uint32_t)min(G + 2 * kThreads, n_full - 1) << 4;
        ra[s & 1] = ldg_stream(reinterpret_cast<const char*>(va) + off);
        if (MODE == G_UNI_CFG) rb[s & 1] = ldg_stream(reinterpret_cast<const char*>(vb) + off);
      }
      const float sc = lsum + neg_log2_E(u01_from_bits(rw[s]));
      if (sc > best && G != special_grp) { best = sc; bgrp = G; }
    }
  }
}

// The first forward's factor of the linear tables by arithmetic instead of a lookup:
// x = bf16(2^(-d log2 e)) exactly as the table builder forms it, d = bf16(lse - l); then
// (dA x + c0)^2 (D-CFG with gamma == 2), dA x + c0 (UDLM without guidance) or x itself
// (MDLM: dA = 1, c0 = 0).  The data-dependent shared-memory lookups are the pipe this kernel
// is bound by: this moves half of them (all of them without guidance) to the XU / FMA pipes.
// dA2 / c02 carry the table's scale (2^(-mA / 2) when squared, 2^-mA otherwise).
template <bool SQUARE, int P0>
__device__ __forceinline__ void cond_factor(const uint4& r, float K, float2 dA2, float2 c02, float2 (&v)[4]) {
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
  const float2 nl2e = make_float2(-kLog2e, -kLog2e);
#pragma unroll
  for (int p = P0; p < 4; ++p) {                      // column pairs P0 .. 3 (the others come from the table)
    const float2 d = sub2(K, bf16_lo(w[p]), bf16_hi(w[p]));
    const __nv_bfloat162 h = __floats2bfloat162_rn(d.x, d.y);      // d >= 0: lse >= every logit
    const uint32_t hw = *reinterpret_cast<const uint32_t*>(&h);
    const float2 t = mul2(make_float2(bf16_lo(hw), bf16_hi(hw)), nl2e);
    const __nv_bfloat162 xh = __floats2bfloat162_rn(ex2a(t.x), ex2a(t.y));
    const uint32_t xw = *reinterpret_cast<const uint32_t*>(&xh);
    const float2 a = fma2(make_float2(bf16_lo(xw), bf16_hi(xw)), dA2, c02);
    v[p] = SQUARE ? mul2(a, a) : a;
  }
}

// ---- WIDE (race_group = 64): linear / log2 sum of the 8 ordinary columns of one vector
template <int MODE, bool ROUND, bool A2 = false>
__device__ __forceinline__ float vec_sum_lin(const RowConst& c, const uint4& ra, const uint4& rb, const float* A,
                                             const float* B, uint32_t neg_base2, float2 dA2 = make_float2(0.f, 0.f),
                                             float2 c02 = make_float2(0.f, 0.f)) {
  float2 va[4];
  if (A2) {
    // D-CFG: all four pairs by arithmetic (the second forward keeps the LSU busy); one forward
    // only: two pairs from the table, two by arithmetic, so that LSU and XU share the row
    constexpr int P0 = MODE == G_UNI_CFG ? 0 : 2;
    if (P0 > 0) lut_gather<P0>(ra, c.Ka, A, neg_base2, va);
    cond_factor<MODE == G_UNI_CFG, P0>(ra, c.Ka, dA2, c02, va);
  } else {
    lut_gather(ra, c.Ka, A, neg_base2, va);
  }
  float2 acc;
  if (MODE == G_UNI_CFG) {
    float2 vb[4];
    lut_gather(rb, c.Kb, B, neg_base2, vb);
    acc = make_float2(va[0].x * vb[0].x, va[0].y * vb[0].y);
#pragma unroll
    for (int p = 1; p < 4; ++p) acc = fma2(va[p], vb[p], acc);
  } else {
    acc = add2(add2(va[0], va[1]), add2(va[2], va[3]));
  }
  return acc.x + acc.y;
}

// log2-domain accumulator (max, sum of 2^(v - max)) for the non-linear paths
struct LogAcc {
  float m = -INFINITY, s = 0.f;
  __device__ __forceinline__ void add(float v) {        // v = log2 of a non-negative term (may be -inf)
    if (!(v > -INFINITY)) return;
    const float nm = fmaxf(m, v);
    s = s * ex2a(m - nm) + ex2a(v - nm);                // m = -inf: s = 0 * 0 + 1
    m = nm;
  }
  __device__ __forceinline__ float value() const { return m + lg2a(s); }   // -inf + lg2(0) = NaN never wins
};

__device__ __forceinline__ uint4 wide_noise_words(int t, int quad, int64_t row, int step, uint2 key) {
  return philox4x32_10(make_uint4((uint32_t)(t | (quad << 8)), (uint32_t)row,
                                  (uint32_t)((uint64_t)row >> 32) | 0x40000000u, (uint32_t)step), key);
}

// One pass over the row, one score per 64-column group (see the header).  extra0 / extra1:
// contribution of the careful vectors this thread owns (z / mask vector, ragged tail) to its
// sweep ek0 / ek1 -- linear units (LIN) or log2 (otherwise); neutral on the other threads.
template <int MODE, bool ROUND, bool LIN, bool A2 = false>
__device__ __forceinline__ void row_pass_wide(const RowConst& c, const uint4* va, const uint4* vb, uint4 ra0,
                                              uint4 rb0, uint4 ra1, uint4 rb1, const float* A, const float* B,
                                              uint32_t neg_base2, int n_full, int n_vec, int special_vec,
                                              int64_t row, int step, uint2 key, float extra0, int ek0, float extra1,
                                              int ek1, float& best, int& bgrp, float dA2s = 0.f, float c02s = 0.f) {
  const float2 dA2 = make_float2(dA2s, dA2s), c02 = make_float2(c02s, c02s);
  const int tid = threadIdx.x;
  uint4 ra[2] = {ra0, ra1}, rb[2] = {rb0, rb1};
  uint4 nw = make_uint4(0, 0, 0, 0);
  for (int k8 = 0; tid + k8 * (8 * kThreads) < n_vec; ++k8) {
    if ((k8 & 3) == 0) nw = wide_noise_words(tid, k8 >> 2, row, step, key);
    float lin_sum = 0.f;
    LogAcc lacc;
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      const int vi = tid + (k8 * 8 + s) * kThreads;
      if (vi >= n_full) break;
      float v;
      if (LIN) v = vec_sum_lin<MODE, ROUND, A2>(c, ra[s & 1], rb[s & 1], A, B, neg_base2, dA2, c02);
      else v = group_hot<MODE, ROUND, false>(c, ra[s & 1], rb[s & 1], A, B, neg_base2);
      {                                               // reload this buffer for two vectors ahead (clamped)
        const uint32_t off = (uint32_t)min(vi + 2 * kThreads, n_full - 1) << 4;
        ra[s & 1] = ldg_stream(reinterpret_cast<const char*>(va) + off);
        if (MODE == G_UNI_CFG) rb[s & 1] = ldg_stream(reinterpret_cast<const char*>(vb) + off);
      }
      if (vi != special_vec) {
        if (LIN) lin_sum += v; else lacc.add(v);
      }
    }
    float lsum;
    if (LIN) {
      lin_sum += (k8 == ek0 ? extra0 : 0.f) + (k8 == ek1 ? extra1 : 0.f);
      lsum = lg2a(lin_sum);                           // lg2(0) = -inf: an empty group never wins
    } else {
      if (k8 == ek0) lacc.add(extra0);
      if (k8 == ek1) lacc.add(extra1);
      lsum = lacc.value();
    }
    const uint32_t w = (k8 & 3) == 0 ? nw.x : (k8 & 3) == 1 ? nw.y : (k8 & 3) == 2 ? nw.z : nw.w;
    const float sc = lsum + neg_log2_E(u01_from_bits(w));
    if (sc > best) { best = sc; bgrp = k8 * kThreads + tid; }
  }
}

template <int MODE, bool ROUND, bool WIDE>
__global__ void __launch_bounds__(kThreads, WIDE ? wide_ctas(MODE) : kCtasPerSm)
sample_group_kernel(const ddg_sample_step_args a) {
  constexpr int NT = MODE == G_UNI_CFG ? 2 : 1;        // logits tensors
  __shared__ __align__(16) float s_A[ROUND ? kLut : 4];
  __shared__ __align__(16) float s_B[(ROUND && MODE == G_UNI_CFG) ? kLut : 4];
  __shared__ __align__(16) float2 s_st[2][NT][kMaxChunks];   // [buffer][tensor][chunk] softmax partials
  __shared__ float s_k[2];
  __shared__ float s_red[kWarps];
  __shared__ int s_idx[kWarps];
  __shared__ int s_win;
  __shared__ float s_pick_f[5];
  __shared__ int s_pick_i[4];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int V = a.vocab;
  const int step = a.step_ptr ? *a.step_ptr : a.step;
  const uint64_t seed64 = a.seed_ptr ? *a.seed_ptr : a.seed;
  const uint2 key = make_uint2((uint32_t)seed64, (uint32_t)(seed64 >> 32));
  const float2* st_mine = reinterpret_cast<const float2*>(warp == 0 ? a.stats_a : a.stats_b);  // warp w: tensor w
  const int n_full = V >> 3;                 // groups made of 8 valid columns
  const int n_groups = (V + 7) >> 3;
  const bool stage = a.stats_chunks <= kMaxChunks;
  // a value the compiler cannot fold (race_group == 8 on this path), so the constant
  // lives in one register instead of being re-materialised in front of every VIADDMNMX
  const uint32_t neg_base2 = kNegBase2 ^ (uint32_t)(a.race_group ^ (WIDE ? 64 : 8));
  float tab_mct = -1.f, tab_mcs = -1.f;     // schedule scalars the table was built for
  bool tab_ready = false, lin = true;
  float goff = 0.f;                          // row-constant offset of the linear tables (log2)
  float tab_mA = 0.f;                        // log2 scale of the conditional table

  // warp w < NT stages tensor w's partials of `row` into buffer `buf` (each lane copies what it reads later)
  auto prefetch_stats = [&](int64_t r, int b) {
    if (stage && warp < NT && r < a.rows) {
      const float2* st = st_mine + r;
      for (int ch = lane; ch < a.stats_chunks; ch += 32) cp_async8(&s_st[b][warp][ch], st + (int64_t)ch * a.stats_ld);
    }
    cp_async_commit();
  };

  int64_t row = blockIdx.x;
  int z_next = row < a.rows ? (int)a.xt[row] : 0;
  prefetch_stats(row, 0);
  int buf = 0;
  for (; row < a.rows; row += gridDim.x, buf ^= 1) {
    const int z = z_next;
    const int64_t row_n = row + gridDim.x;
    if (row_n < a.rows) z_next = (int)a.xt[row_n];
    prefetch_stats(row_n, buf ^ 1);
    if (MODE == G_ABS_NONE && z != a.mask_index) {      // carry-over (diffusion.py:1414-1418)
      if (tid == 0) a.xs[row] = z;
      continue;
    }
    const __nv_bfloat16* pa = reinterpret_cast<const __nv_bfloat16*>(a.logits_a) + row * a.ld_a;
    const __nv_bfloat16* pb = MODE == G_UNI_CFG ? reinterpret_cast<const __nv_bfloat16*>(a.logits_b) + row * a.ld_b : pa;
    const uint4* va = reinterpret_cast<const uint4*>(pa) + tid;   // this thread's groups: va[128 k]
    const uint4* vb = reinterpret_cast<const uint4*>(pb) + tid;
    // first vectors and the logits at z in flight while the normalisers / table are prepared
    uint4 ra0, rb0 = make_uint4(0, 0, 0, 0), ra1, rb1 = make_uint4(0, 0, 0, 0);
    {
      const int g0 = min(tid, n_full - 1) - tid, g1 = min(tid + kThreads, n_full - 1) - tid;
      ra0 = ldg_stream(va + g0);
      ra1 = ldg_stream(va + g1);
      if (MODE == G_UNI_CFG) {
        rb0 = ldg_stream(vb + g0);
        rb1 = ldg_stream(vb + g1);
      }
    }
    __nv_bfloat16 lz_a = pa[0], lz_b = pa[0];
    if (MODE != G_ABS_NONE) {
      lz_a = pa[z];
      if (MODE == G_UNI_CFG) lz_b = pb[z];
    }
    // ---- row normalisers from the GEMM partials (warp w < NT: tensor w)
    if (warp < NT) {
      float m = -INFINITY, s = 0.f;
      if (stage) {
        cp_async_wait<1>();                              // this row's copies (the next row's may be in flight)
        for (int ch = lane; ch < a.stats_chunks; ch += 32) {
          const float2 pr = s_st[buf][warp][ch];
          const float nm = fmaxf(m, pr.x);
          s = s * ex2a((m - nm) * kLog2e) + pr.y * ex2a((pr.x - nm) * kLog2e);
          m = nm;
        }
      } else {
        const float2* st = st_mine + row;
        for (int ch = lane; ch < a.stats_chunks; ch += 32) {
          const float2 pr = st[(int64_t)ch * a.stats_ld];
          const float nm = fmaxf(m, pr.x);
          s = s * ex2a((m - nm) * kLog2e) + pr.y * ex2a((pr.x - nm) * kLog2e);
          m = nm;
        }
      }
      const float M = warp_max(m);
      s = (m == -INFINITY) ? 0.f : s * ex2a((m - M) * kLog2e);
      const float S = warp_sum(s);
      // lse >= max exactly, so that lse - l >= 0 for every column (table index math)
      if (lane == 0) s_k[warp] = M + fmaxf(kLn2 * lg2a(S), 0.f);
    }
    RowConst c;
    c.z = z; c.mask_index = a.mask_index; c.V = V;
    c.gam = a.gamma; c.omg = 1.f - a.gamma;
    const int64_t seq = a.rows <= 0x7fffffff ? (int64_t)((uint32_t)row / (uint32_t)a.seq_len) : row / a.seq_len;
    const int64_t mc_off = seq * a.mc_seq_stride + (int64_t)step * a.mc_step_stride;
    const float mc_t = a.mc_t[mc_off], mc_s = a.mc_s[mc_off];
    float atV = 0.f, zoff = 0.f;
    if (MODE == G_ABS_NONE) {
      const float r = mc_s / mc_t;
      c.mask_score = lg2a(r) - lg2a(1.f - r);
      c.dA = 0.f; c.c0 = 0.f;
    } else {
      const float alpha_s = 1.f - mc_s, alpha_t = 1.f - mc_t;
      const float alpha_ts = alpha_t / alpha_s;
      c.dA = alpha_s - alpha_t;
      c.c0 = (1.f - alpha_ts) * (1.f - alpha_s) * (1.f / (float)V);
      atV = alpha_t * (float)V;
      zoff = alpha_ts - alpha_t;
      c.mask_score = 0.f;
    }
    if (ROUND) {
      // ---- per-column factor tables over the bf16 values of |lp| (see the header comment)
      const bool stale = MODE == G_ABS_NONE ? !tab_ready : (!tab_ready || mc_t != tab_mct || mc_s != tab_mcs);
      if (stale) {                                   // block-uniform
        float mA = 0.f, mB = 0.f;
        if (MODE == G_ABS_NONE) {
          lin = true; goff = 0.f;
        } else {
          const float lq_hi = lg2a(c.dA + c.c0), lq_lo = lg2a(c.c0);   // x = 1 / x = 0
          if (MODE == G_UNI_NONE) {
            lin = (lq_hi - lq_lo) <= kLinearRange; mA = lq_hi;
          } else {
            lin = (fabsf(c.gam) + fabsf(c.omg)) * (lq_hi - lq_lo) <= kLinearRange;
            mA = fmaxf(__fmul_rn(c.gam, lq_hi), __fmul_rn(c.gam, lq_lo));
            mB = fmaxf(__fmul_rn(c.omg, lq_hi), __fmul_rn(c.omg, lq_lo));
          }
          goff = lin ? mA + mB : 0.f;
          tab_mA = mA;
        }
        for (int i = tid; i < kLut; i += kThreads) {
          const float mag = __uint_as_float((uint32_t)(kLutBase + i) << 16);
          const float x = bf16_round(ex2a(-mag * kLog2e));
          if (MODE == G_ABS_NONE) {
            s_A[i] = x;                                             // q~ = x (bf16-exact), linear
          } else {
            const float lq = lg2a(fmaf(c.dA, x, c.c0));
            const float ta = MODE == G_UNI_CFG ? __fmul_rn(c.gam, lq) : lq;
            s_A[i] = lin ? ex2a(ta - mA) : ta;
            if (MODE == G_UNI_CFG) {
              const float tb = __fmul_rn(c.omg, lq);
              s_B[i] = lin ? ex2a(tb - mB) : tb;
            }
          }
        }
        tab_ready = true; tab_mct = mc_t; tab_mcs = mc_s;
      }
    }
    __syncthreads();                                     // B1: s_k and the tables are ready
    c.Ka = s_k[0];
    c.Kb = MODE == G_UNI_CFG ? s_k[1] : 0.f;
    c.ka2 = c.Ka * kLog2e;
    c.kb2 = c.Kb * kLog2e;
    int special;                                        // column needing the careful path
    if (MODE == G_ABS_NONE) {
      special = a.mask_index;
      c.zc_a = c.zc_b = 0.f;
    } else {
      c.zc_a = atV * x_of<ROUND>(__bfloat162float(lz_a), c.Ka, c.ka2) + zoff;
      c.zc_b = MODE == G_UNI_CFG ? atV * x_of<ROUND>(__bfloat162float(lz_b), c.Kb, c.kb2) + zoff : 0.f;
      special = z;
    }
    const int special_grp = special >> 3;

    float best = -INFINITY;
    int bgrp = 0x7fffffff;
    if (WIDE) {
      // ---- careful vectors (the one holding z / the mask column, the ragged tail) are left
      // out of the hot sums and added to their owner thread's group here
      const int n_vec = n_groups;
      const int special_vec = special_grp;
      float extra[2];
      int ek[2];
      const bool use_lin = ROUND && lin;
#pragma unroll
      for (int sp = 0; sp < 2; ++sp) {
        const int vec = sp == 0 ? special_vec : n_full;
        const bool exists = sp == 0 || (n_full != n_groups && special_vec != n_full);
        extra[sp] = use_lin ? 0.f : -INFINITY;
        ek[sp] = -1;
        const int owner = vec & (kThreads - 1);
        if (exists && warp == (owner >> 5)) {             // warp-uniform: the owner's warp, lane e % 8 = column e
          const int j = vec * 8 + (lane & 7);             // rows are padded to whole vectors: in bounds
          const float la = __bfloat162float(pa[j]);
          const float lb = MODE == G_UNI_CFG ? __bfloat162float(pb[j]) : 0.f;
          const float g = g_elem<MODE, ROUND>(c, la, lb, j);
          float gm = g;
#pragma unroll
          for (int o = 4; o > 0; o >>= 1) gm = fmaxf(gm, __shfl_xor_sync(0xffffffffu, gm, o));
          float sm = gm > -INFINITY ? ex2a(g - gm) : 0.f;
#pragma unroll
          for (int o = 4; o > 0; o >>= 1) sm += __shfl_xor_sync(0xffffffffu, sm, o);
          if (lane == (owner & 31) && gm > -INFINITY) {
            const float lg = (gm - goff) + lg2a(sm);
            extra[sp] = use_lin ? ex2a(lg) : lg;
            ek[sp] = (vec / kThreads) >> 3;
          }
        }
      }
      if (use_lin && (MODE != G_UNI_CFG || c.gam == 2.f)) {
        // table scale: (dA x + c0)^2 / 2^mA = ((dA x + c0) 2^(-mA / 2))^2;  (dA x + c0) / 2^mA;  x
        const float sc2 = MODE == G_UNI_CFG ? ex2a(-0.5f * tab_mA) : ex2a(-tab_mA);
        const float dAs = MODE == G_ABS_NONE ? 1.f : c.dA * sc2, c0s = MODE == G_ABS_NONE ? 0.f : c.c0 * sc2;
        row_pass_wide<MODE, ROUND, true, true>(c, reinterpret_cast<const uint4*>(pa), reinterpret_cast<const uint4*>(pb), ra0, rb0, ra1, rb1, s_A, s_B, neg_base2, n_full, n_vec, special_vec, row, step, key, extra[0], ek[0], extra[1], ek[1], best, bgrp, dAs, c0s);
      } else if (use_lin)
        row_pass_wide<MODE, ROUND, true>(c, reinterpret_cast<const uint4*>(pa), reinterpret_cast<const uint4*>(pb), ra0, rb0, ra1, rb1, s_A, s_B, neg_base2, n_full, n_vec, special_vec, row, step, key, extra[0], ek[0], extra[1], ek[1], best, bgrp);
      else
        row_pass_wide<MODE, ROUND, false>(c, reinterpret_cast<const uint4*>(pa), reinterpret_cast<const uint4*>(pb), ra0, rb0, ra1, rb1, s_A, s_B, neg_base2, n_full, n_vec, special_vec, row, step, key, extra[0], ek[0], extra[1], ek[1], best, bgrp);
    } else {
    // ---- one pass over the row: per full group of 8 columns, log2(sum q~) - log2(E).
    // The group holding z / the mask column is evaluated too (branch-free) but its
    // score is discarded there and redone on the careful path below.
    if (ROUND && lin)
      row_pass<MODE, ROUND, true>(c, reinterpret_cast<const uint4*>(pa), reinterpret_cast<const uint4*>(pb), ra0, rb0, ra1, rb1, s_A, s_B, neg_base2, n_full, special_grp, row, step, key, best, bgrp);
    else
      row_pass<MODE, ROUND, false>(c, reinterpret_cast<const uint4*>(pa), reinterpret_cast<const uint4*>(pb), ra0, rb0, ra1, rb1, s_A, s_B, neg_base2, n_full, special_grp, row, step, key, best, bgrp);
    // ---- the (at most two) groups that need the careful path: the one holding z /
    // the mask column and the ragged tail; each is done by the thread that owns it
#pragma unroll 1
    for (int sp = 0; sp < 2; ++sp) {
      const int G = sp == 0 ? special_grp : n_full;
      if (sp == 1 && (n_full == n_groups || special_grp == n_full)) break;
      if (tid == (G & (kThreads - 1))) {
        const uint4 ca = __ldg(va + (G - tid));
        const uint4 cb = MODE == G_UNI_CFG ? __ldg(vb + (G - tid)) : ca;
        float g[8];
        group_careful<MODE, ROUND>(c, ca, cb, G * 8, g);
        const float gm = max8(g);
        if (gm > -INFINITY) {
          const float sc = (gm - goff) + lg2a(sum_ex2_8(g, gm)) +
                           neg_log2_E(u01_from_bits(group_noise_word(G, row, step, key)));
          if (sc > best || (sc == best && G < bgrp)) { best = sc; bgrp = G; }
        }
      }
    }
    }
    // ---- block argmax over groups, lowest group index wins ties
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float os = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bgrp, o);
      if (os > best || (os == best && oi < bgrp)) { best = os; bgrp = oi; }
    }
    if (lane == 0) { s_red[warp] = best; s_idx[warp] = bgrp; }
    __syncthreads();                                     // B2
    if (warp == 0) {
      best = lane < kWarps ? s_red[lane] : -INFINITY;
      bgrp = lane < kWarps ? s_idx[lane] : 0x7fffffff;
#pragma unroll
      for (int o = kWarps / 2; o > 0; o >>= 1) {
        const float os = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bgrp, o);
        if (os > best || (os == best && oi < bgrp)) { best = os; bgrp = oi; }
      }
      if (lane == 0) s_win = bgrp;
    }
    __syncthreads();                                     // B3
    const int win = s_win;
    if (win == 0x7fffffff) {
      if (tid == 0) { a.xs[row] = 0; if (a.seq_changed && z != 0) a.seq_changed[seq] = 1; }   // all-zero row: cannot happen with finite logits
      if (WIDE) { __syncthreads(); __syncthreads(); }   // keep the barrier count of the pick below
    } else if (WIDE) {
      // ---- inverse-CDF pick inside the winning 64-column group: thread i < 64 takes column
      // (s, e) = (i / 8, i % 8); shuffle scan inside each of the two warps, offsets through smem
      const int wt = win & (kThreads - 1), wk8 = win / kThreads;
      float g = -INFINITY;
      int j = 0;
      if (tid < 64) {
        const int vi = wt + kThreads * (8 * wk8 + (tid >> 3));
        j = vi * 8 + (tid & 7);
        if (j < V) {
          const float la = __bfloat162float(pa[j]);
          const float lb = MODE == G_UNI_CFG ? __bfloat162float(pb[j]) : 0.f;
          g = g_elem<MODE, ROUND>(c, la, lb, j);
        }
      }
      const float wm = warp_max(g);
      if (lane == 0 && warp < 2) s_pick_f[warp] = wm;
      if (warp == 2) {                                   // idle in the pick: draws its uniform
        const uint4 r2 = philox4x32_10(
          make_uint4((uint32_t)win, (uint32_t)row, (uint32_t)((uint64_t)row >> 32) | 0x80000000u, (uint32_t)step), key);
        if (lane == 0) s_pick_f[4] = u01_from_bits(r2.x);
      }
      __syncthreads();                                   // P1
      const float gm = fmaxf(s_pick_f[0], s_pick_f[1]);
      const float pr = (tid < 64 && g > -INFINITY) ? ex2a(g - gm) : 0.f;
      float cum = pr;                                    // inclusive scan inside the warp
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float up = __shfl_up_sync(0xffffffffu, cum, o);
        if (lane >= o) cum += up;
      }
      if (lane == 31 && warp < 2) s_pick_f[2 + warp] = cum;
      __syncthreads();                                   // P2
      if (tid < 64) {
        const float S = s_pick_f[2] + s_pick_f[3];
        if (warp == 1) cum += s_pick_f[2];
        const float target = s_pick_f[4] * S;
        const uint32_t hit = __ballot_sync(0xffffffffu, cum > target);
        const uint32_t pos = __ballot_sync(0xffffffffu, pr > 0.f);
        // first column whose running sum passes the target, else (rounding) the last positive one
        if (lane == 0) {
          s_pick_i[warp] = hit ? __ffs(hit) - 1 : -1;
          s_pick_i[2 + warp] = pos ? 31 - __clz(pos) : -1;
        }
      }
      __syncthreads();                                   // P3
      if (tid == 0) {
        int i;
        if (s_pick_i[0] >= 0) i = s_pick_i[0];
        else if (s_pick_i[1] >= 0) i = 32 + s_pick_i[1];
        else if (s_pick_i[3] >= 0) i = 32 + s_pick_i[3];
        else i = s_pick_i[2] >= 0 ? s_pick_i[2] : 0;
        const int64_t pick_id = (int64_t)(wt + kThreads * (8 * wk8 + (i >> 3))) * 8 + (i & 7);
        a.xs[row] = pick_id;
        if (a.seq_changed && pick_id != z) a.seq_changed[seq] = 1;   // this sequence needs a new forward
      }
    } else if (tid == (win & (kThreads - 1))) {
      // ---- inverse-CDF pick inside the winning group
      const uint4 ca = __ldg(va + (win - tid));
      const uint4 cb = MODE == G_UNI_CFG ? __ldg(vb + (win - tid)) : ca;
      float g[8];
      group_careful<MODE, ROUND>(c, ca, cb, win * 8, g);
      const float gm = max8(g);
      float pr[8], S = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) { pr[e] = ex2a(g[e] - gm); S += pr[e]; }
      const uint4 r2 = philox4x32_10(
        make_uint4((uint32_t)win, (uint32_t)row, (uint32_t)((uint64_t)row >> 32) | 0x80000000u, (uint32_t)step), key);
      const float target = u01_from_bits(r2.x) * S;
      int pick = -1, last = 0;
      float cum = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        cum += pr[e];
        if (pr[e] > 0.f) last = e;
        if (pick < 0 && cum > target) pick = e;
      }
      const int pick_id = win * 8 + (pick < 0 ? last : pick);
      a.xs[row] = pick_id;
      if (a.seq_changed && pick_id != z) a.seq_changed[seq] = 1;
    }
    // Shared state is rewritten by the next row only before its B1 (s_k, tables, the other
    // s_st buffer), before its B2 (s_red / s_idx) and before its B3 (s_win); every read of
    // this row precedes the barrier after which those writes can happen.
  }
  cp_async_wait<0>();
}

template <int MODE, bool ROUND>
int launch(const ddg_sample_step_args& a, cudaStream_t s) {
  int64_t grid = (int64_t)sm_count() * (a.race_group == 64 ? wide_ctas(MODE) : kCtasPerSm) * 2;   // two waves of resident CTAs
  if (grid > a.rows) grid = a.rows;
  if (a.race_group == 64)
    sample_group_kernel<MODE, ROUND, true><<<(unsigned)grid, kThreads, 0, s>>>(a);
  else
    sample_group_kernel<MODE, ROUND, false><<<(unsigned)grid, kThreads, 0, s>>>(a);
  return check_cuda(cudaGetLastError(), "sample_group launch");
}

}  // namespace

// Called by try_sample_fused once alignment / stats have been validated (race_group 8 or 64).
int launch_sample_group(int mode, const ddg_sample_step_args& a, cudaStream_t s) {
  switch (mode * 2 + (a.round_bf16 ? 1 : 0)) {
    case 0: return launch<G_UNI_NONE, false>(a, s);
    case 1: return launch<G_UNI_NONE, true>(a, s);
    case 2: return launch<G_UNI_CFG, false>(a, s);
    case 3: return launch<G_UNI_CFG, true>(a, s);
    case 4: return launch<G_ABS_NONE, false>(a, s);
    default: return launch<G_ABS_NONE, true>(a, s);
  }
}

}  // namespace ddg
