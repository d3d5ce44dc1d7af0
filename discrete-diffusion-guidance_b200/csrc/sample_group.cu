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
// column of the noise drop by 8x, which is what keeps this kernel on the HBM
// roofline instead of the MUFU / issue limit.  (With injected noise the
// per-column race of sample_fused.cu / sample_step.cu is used: that is the
// path that is bit-comparable with the reference.)
//
// round_bf16 = 1 (the reference's bf16 log-softmax / exp roundings, SURVEY
// 8a' item 4): lp = bf16(l - lse) takes only ~2000 distinct values that matter,
// so  gamma*log2(dA*bf16(exp(lp)) + c0)  is a 2048-entry shared-memory table
// per forward (built per CTA, rebuilt if the schedule scalars change) and the
// hot loop has no transcendental per column at all except the exp of the
// group sum.
//
// Philox stream (so tests can regenerate it; G = j / 8 is the group index):
//   key     = (seed_lo, seed_hi)
//   group exponential: counter = ((G & 255) | ((G >> 10) << 8), row_lo, row_hi, step),
//                      word (G >> 8) & 3,  U = (word >> 8) * 2^-24,
//                      E = 1e-10 - ln(U + 1e-10)
//   in-group pick:     counter = (G, row_lo, row_hi | 0x80000000, step), word 0
#include "ddg_common.cuh"

namespace ddg {
namespace {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr int kThreads = 256;
constexpr int kLut = 2048;          // bf16 magnitudes 0x3B00 .. 0x42FF: |lp| in [2^-9, 128)
constexpr int kLutBase = 0x3B00;

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

// packed f32x2 subtract (one issue slot for two differences)
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

// Table byte offsets of a pair of differences d = lse - l >= 0: bf16 round,
// (magnitude bits - 0x3B00) clamped to [0, 2047] with one DPX instruction for
// the pair, times 8 = the stride of the interleaved (a-term, b-term) table
// (both halves at once: 8 * 2047 still fits 16 bits).
constexpr uint32_t kNegBase2 = (uint32_t)(0x10000 - kLutBase) * 0x00010001u;   // (-0x3B00, -0x3B00) as s16x2
__device__ __forceinline__ uint32_t lut_offsets2(float d0, float d1, uint32_t neg_base2) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(d0, d1);
  const uint32_t w = *reinterpret_cast<const uint32_t*>(&h);
  constexpr uint32_t kMaxIdx = (uint32_t)(kLut - 1) * 0x00010001u;
  return __viaddmin_s16x2_relu(w, neg_base2, kMaxIdx) << 3;
}
__device__ __forceinline__ float lut_at(const float* T, uint32_t byte_off) {
  return *reinterpret_cast<const float*>(reinterpret_cast<const char*>(T) + byte_off);
}

// log2 of the unnormalised guided posterior of 8 ordinary columns (no z, no
// mask column, all < V): branch-free
template <int MODE, bool ROUND>
__device__ __forceinline__ void group_hot(const RowConst& c, const uint4& ra, const uint4& rb, const float* T,
                                          uint32_t neg_base2, float (&g)[8]) {
  const uint32_t wa[4] = {ra.x, ra.y, ra.z, ra.w};
  const uint32_t wb[4] = {rb.x, rb.y, rb.z, rb.w};
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    if (ROUND) {
      const float2 da = sub2(c.Ka, bf16_lo(wa[p]), bf16_hi(wa[p]));
      const uint32_t ta = lut_offsets2(da.x, da.y, neg_base2);
      float2 gp = make_float2(lut_at(T, ta & 0xffffu), lut_at(T, ta >> 16));
      if (MODE == G_UNI_CFG) {
        const float2 db = sub2(c.Kb, bf16_lo(wb[p]), bf16_hi(wb[p]));
        const uint32_t tb = lut_offsets2(db.x, db.y, neg_base2);
        gp = add2(gp, make_float2(lut_at(T + 1, tb & 0xffffu), lut_at(T + 1, tb >> 16)));
      }
      g[2 * p] = gp.x;
      g[2 * p + 1] = gp.y;
    } else if (MODE == G_ABS_NONE) {
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

// the same for a group that may hold z / the mask column / columns >= V
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

// group exponential of group G (see the stream layout at the top of the file)
__device__ __forceinline__ uint4 group_noise_words(int tid, int quad, int64_t row, int step, uint2 key) {
  return philox4x32_10(
    make_uint4((uint32_t)(tid | (quad << 8)), (uint32_t)row, (uint32_t)((uint64_t)row >> 32), (uint32_t)step), key);
}

template <int MODE, bool ROUND>
__global__ void __launch_bounds__(kThreads, 3)
sample_group_kernel(const ddg_sample_step_args a) {
  __shared__ __align__(16) float s_T[ROUND ? 2 * kLut : 4];   // (a-term, b-term) interleaved
  __shared__ float s_k[2];
  __shared__ float s_red[kThreads / 32];
  __shared__ int s_idx[kThreads / 32];
  __shared__ int s_win;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int V = a.vocab;
  const int step = a.step_ptr ? *a.step_ptr : a.step;
  const uint64_t seed64 = a.seed_ptr ? *a.seed_ptr : a.seed;
  const uint2 key = make_uint2((uint32_t)seed64, (uint32_t)(seed64 >> 32));
  const float2* st_a = reinterpret_cast<const float2*>(a.stats_a);
  const float2* st_b = reinterpret_cast<const float2*>(a.stats_b);
  const int n_full = V >> 3;                 // groups made of 8 valid columns
  const int n_groups = (V + 7) >> 3;
  uint32_t neg_base2;                        // opaque to constant propagation: stays in one register
  asm volatile("mov.u32 %0, %1;" : "=r"(neg_base2) : "n"(kNegBase2));
  float tab_mct = -1.f, tab_mcs = -1.f;     // schedule scalars the table was built for
  bool tab_ready = false;

  for (int64_t row = blockIdx.x; row < a.rows; row += gridDim.x) {
    const int z = (int)a.xt[row];
    if (MODE == G_ABS_NONE && z != a.mask_index) {      // carry-over (diffusion.py:1414-1418)
      if (tid == 0) a.xs[row] = z;
      continue;
    }
    // this thread's groups: G = tid + 256 * it  ->  vector G of the row
    const uint4* va = reinterpret_cast<const uint4*>(
      reinterpret_cast<const __nv_bfloat16*>(a.logits_a) + row * a.ld_a) + tid;
    const uint4* vb = MODE == G_UNI_CFG ? reinterpret_cast<const uint4*>(
      reinterpret_cast<const __nv_bfloat16*>(a.logits_b) + row * a.ld_b) + tid : va;
    // first vectors in flight while the normalisers and the table are prepared
    uint4 ra = make_uint4(0, 0, 0, 0), rb = make_uint4(0, 0, 0, 0);
    if (tid < n_full) {
      ra = ldg_stream(va);
      if (MODE == G_UNI_CFG) rb = ldg_stream(vb);
    }
    // ---- row normalisers from the GEMM partials (warps 0 and 1)
    if (warp < (MODE == G_UNI_CFG ? 2 : 1)) {
      const float2* st = (warp == 0 ? st_a : st_b) + row;
      float m = -INFINITY, s = 0.f;
      for (int ch = lane; ch < a.stats_chunks; ch += 32) {
        const float2 pr = st[(int64_t)ch * a.stats_ld];
        const float nm = fmaxf(m, pr.x);
        s = s * ex2a((m - nm) * kLog2e) + pr.y * ex2a((pr.x - nm) * kLog2e);
        m = nm;
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
    const int64_t mc_off = (row / a.seq_len) * a.mc_seq_stride + (int64_t)step * a.mc_step_stride;
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
      // ---- table of gamma * log2(dA * bf16(exp(-|lp|)) + c0) over the bf16 values of |lp|
      const bool stale = MODE == G_ABS_NONE ? !tab_ready : (!tab_ready || mc_t != tab_mct || mc_s != tab_mcs);
      if (stale) {                                   // block-uniform
        for (int i = tid; i < kLut; i += kThreads) {
          const float mag = __uint_as_float((uint32_t)(kLutBase + i) << 16);
          const float x = bf16_round(ex2a(-mag * kLog2e));
          if (MODE == G_ABS_NONE) {
            s_T[2 * i] = lg2a(x);
          } else {
            const float lq = lg2a(fmaf(c.dA, x, c.c0));
            if (MODE == G_UNI_CFG) {
              s_T[2 * i] = __fmul_rn(c.gam, lq);
              s_T[2 * i + 1] = __fmul_rn(c.omg, lq);
            } else {
              s_T[2 * i] = lq;
            }
          }
        }
        tab_ready = true; tab_mct = mc_t; tab_mcs = mc_s;
      }
    }
    __syncthreads();
    c.Ka = s_k[0];
    c.Kb = MODE == G_UNI_CFG ? s_k[1] : 0.f;
    c.ka2 = c.Ka * kLog2e;
    c.kb2 = c.Kb * kLog2e;
    int special;                                        // column needing the careful path
    if (MODE == G_ABS_NONE) {
      special = a.mask_index;
      c.zc_a = c.zc_b = 0.f;
    } else {
      const __nv_bfloat16* pa = reinterpret_cast<const __nv_bfloat16*>(a.logits_a) + row * a.ld_a;
      const __nv_bfloat16* pb = reinterpret_cast<const __nv_bfloat16*>(a.logits_b) + row * a.ld_b;
      c.zc_a = atV * x_of<ROUND>(__bfloat162float(pa[z]), c.Ka, c.ka2) + zoff;
      c.zc_b = MODE == G_UNI_CFG ? atV * x_of<ROUND>(__bfloat162float(pb[z]), c.Kb, c.kb2) + zoff : 0.f;
      special = z;
    }
    const int special_grp = special >> 3;

    // ---- one pass over the row: per full group of 8 columns, log2(sum q~) - log2(E).
    // The group holding z / the mask column is evaluated too (branch-free) but its
    // score is discarded here and redone on the careful path below.
    float best = -INFINITY;
    int bgrp = 0x7fffffff;
    for (int q = 0; tid + q * (4 * kThreads) < n_full; ++q) {
      const uint4 rnd = group_noise_words(tid, q, row, step, key);
      const uint32_t rw[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
#pragma unroll
      for (int s4 = 0; s4 < 4; ++s4) {
        const int G = tid + (q * 4 + s4) * kThreads;
        if (G >= n_full) break;
        const uint4 ca = ra, cb = rb;
        {                                               // prefetch the next vectors (clamped: always in bounds)
          const int gn = min(G + kThreads, n_full - 1) - tid;
          ra = ldg_stream(va + gn);
          if (MODE == G_UNI_CFG) rb = ldg_stream(vb + gn);
        }
        float g[8];
        group_hot<MODE, ROUND>(c, ca, cb, s_T, neg_base2, g);
        const float gm = max8(g);
        const float S = sum_ex2_8(g, gm);
        const float sc = gm + lg2a(S) + neg_log2_E(u01_from_bits(rw[s4]));
        // (gm == -inf gives NaN, which never compares greater)
        if (sc > best && G != special_grp) { best = sc; bgrp = G; }
      }
    }
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
          const float S = sum_ex2_8(g, gm);
          const uint4 rnd = group_noise_words(tid, G >> 10, row, step, key);
          const int wsel = (G >> 8) & 3;
          const uint32_t word = wsel == 0 ? rnd.x : wsel == 1 ? rnd.y : wsel == 2 ? rnd.z : rnd.w;
          const float sc = gm + lg2a(S) + neg_log2_E(u01_from_bits(word));
          if (sc > best || (sc == best && G < bgrp)) { best = sc; bgrp = G; }
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
    __syncthreads();
    if (warp == 0) {
      best = lane < kThreads / 32 ? s_red[lane] : -INFINITY;
      bgrp = lane < kThreads / 32 ? s_idx[lane] : 0x7fffffff;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float os = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bgrp, o);
        if (os > best || (os == best && oi < bgrp)) { best = os; bgrp = oi; }
      }
      if (lane == 0) s_win = bgrp;
    }
    __syncthreads();
    const int win = s_win;
    if (win == 0x7fffffff) {
      if (tid == 0) a.xs[row] = 0;                      // all-zero row: cannot happen with finite logits
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
      a.xs[row] = win * 8 + (pick < 0 ? last : pick);
    }
    // s_k / s_red / the table are rewritten only after every thread has passed this
    // row's last read of them (the barrier above); s_win is rewritten after the next
    // row's second barrier.  No trailing barrier needed.
  }
}

template <int MODE, bool ROUND>
int launch(const ddg_sample_step_args& a, cudaStream_t s) {
  int64_t grid = (int64_t)sm_count() * 4 * 4;     // a few waves of resident CTAs
  if (grid > a.rows) grid = a.rows;
  sample_group_kernel<MODE, ROUND><<<(unsigned)grid, kThreads, 0, s>>>(a);
  return check_cuda(cudaGetLastError(), "sample_group launch");
}

}  // namespace

// Called by try_sample_fused once alignment / stats have been validated.
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
