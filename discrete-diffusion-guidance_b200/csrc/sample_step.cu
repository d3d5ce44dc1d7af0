// Fused per-token sampler: log-softmax -> posterior (MDLM absorbing / UDLM
// uniform) -> D-CFG / D-CBG mix -> exponential-race (Gumbel-max) selection.
//
// What it computes follows the reference (paths relative to the reference
// root): diffusion.py:321-337 & 356-365 (log-probs), :367-394 (UDLM
// posterior), :1399-1401 (MDLM posterior), :1475-1495 (D-CFG), :1587-1617
// (D-CBG), :1748-1755 (categorical draw), :1414-1418 (carry-over).  How it
// computes it is new: one launch, the selection is done in the log domain
//     xs = argmax_j [ g_j - log E_j ],   E_j = 1e-10 - log(U_j + 1e-10)
// where g_j is the log of the (unnormalised) guided posterior, so the final
// softmax of the reference is never materialised (it cancels in the argmax).
//
// This file holds the generic kernel: a group of threads (one warp for small
// vocabularies, one 256-thread CTA for large ones) owns a token row and walks
// it two times (normalisers, then selection); the second walk hits L1/L2.
#include "ddg_common.cuh"

namespace ddg {
int try_sample_fused(const ddg_sample_step_args& a, cudaStream_t s);   // sample_fused.cu
int launch_sample_f64(const ddg_sample_step_args& a, cudaStream_t s);  // sample_aux.cu
namespace {

constexpr float kNegInf = -1000000.0f;  // diffusion.py:141

enum Mode : int {
  ABS_NONE = 0, ABS_CFG = 1, ABS_CBG = 2,
  UNI_NONE = 3, UNI_CFG = 4, UNI_CBG = 5,
  AR_NONE = 6, AR_CFG = 7
};
__host__ __device__ constexpr bool mode_two(int m) { return m == ABS_CFG || m == UNI_CFG || m == AR_CFG; }
__host__ __device__ constexpr bool mode_clf(int m) { return m == ABS_CBG || m == UNI_CBG; }
__host__ __device__ constexpr bool mode_abs(int m) { return m <= ABS_CBG; }
__host__ __device__ constexpr bool mode_uni(int m) { return m >= UNI_NONE && m <= UNI_CBG; }
// modes whose q_xs the reference passes through a final softmax
// (diffusion.py:1495, 1617); the others return the posterior as formed.
__host__ __device__ constexpr bool mode_softmax(int m) { return m == UNI_CFG || m == ABS_CBG || m == UNI_CBG; }

template <bool FAST> struct Mth {
  static __device__ __forceinline__ float exp(float x) { return FAST ? __expf(x) : expf(x); }
  static __device__ __forceinline__ float log(float x) { return FAST ? __logf(x) : logf(x); }
  static __device__ __forceinline__ float mul(float a, float b) { return FAST ? a * b : __fmul_rn(a, b); }
  static __device__ __forceinline__ float add(float a, float b) { return FAST ? a + b : __fadd_rn(a, b); }
  static __device__ __forceinline__ float div(float a, float b) { return FAST ? __fdividef(a, b) : __fdiv_rn(a, b); }
};

// ------------------------------------------------------------ row loads
template <typename T, int VEC> struct RowLoad;
template <> struct RowLoad<__nv_bfloat16, 8> {
  static __device__ __forceinline__ void ld(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 r = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
};
template <> struct RowLoad<float, 4> {
  static __device__ __forceinline__ void ld(const float* p, float (&v)[4]) {
    const float4 r = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w;
  }
};
template <typename T> struct RowLoad<T, 1> {
  static __device__ __forceinline__ void ld(const T* p, float (&v)[1]) {
    v[0] = Cvt<T>::to_f32(__ldg(p));
  }
};

// ------------------------------------------------------------ reductions
template <int GROUP> struct Grp {
  // smem: GROUP/32 floats + GROUP/32 ints (only for GROUP > 32)
  static __device__ __forceinline__ float max(float v, float* sm) {
    v = warp_max(v);
    if (GROUP > 32) {
      const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
      __syncthreads();
      if (l == 0) sm[w] = v;
      __syncthreads();
      v = (l < GROUP / 32) ? sm[l] : -INFINITY;
      v = warp_max(v);
    }
    return v;
  }
  static __device__ __forceinline__ float sum(float v, float* sm) {
    v = warp_sum(v);
    if (GROUP > 32) {
      const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
      __syncthreads();
      if (l == 0) sm[w] = v;
      __syncthreads();
      v = (l < GROUP / 32) ? sm[l] : 0.f;
      v = warp_sum(v);
    }
    return v;
  }
  // first index wins on ties (torch.argmax semantics)
  static __device__ __forceinline__ void argmax(float& s, int& i, float* sm) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float os = __shfl_xor_sync(0xffffffffu, s, o);
      const int oi = __shfl_xor_sync(0xffffffffu, i, o);
      if (os > s || (os == s && oi < i)) { s = os; i = oi; }
    }
    if (GROUP > 32) {
      int* smi = reinterpret_cast<int*>(sm + GROUP / 32);
      const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
      __syncthreads();
      if (l == 0) { sm[w] = s; smi[w] = i; }
      __syncthreads();
      if (l < GROUP / 32) { s = sm[l]; i = smi[l]; } else { s = -INFINITY; i = 0x7fffffff; }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float os = __shfl_xor_sync(0xffffffffu, s, o);
        const int oi = __shfl_xor_sync(0xffffffffu, i, o);
        if (os > s || (os == s && oi < i)) { s = os; i = oi; }
      }
    }
  }
};

// online (max, sum-exp) accumulator
template <bool FAST> struct LseAcc {
  float m = -INFINITY, s = 0.f;
  __device__ __forceinline__ void push_max(float vm) {
    if (vm > m) { s *= Mth<FAST>::exp(m - vm); m = vm; }
  }
  __device__ __forceinline__ void add(float v) { s += Mth<FAST>::exp(v - m); }
};
template <int GROUP, bool FAST>
__device__ __forceinline__ void lse_finish(float m, float s, float* sm, float& M, float& lnS) {
  M = Grp<GROUP>::max(m, sm);
  const float sc = (m == -INFINITY) ? 0.f : s * Mth<FAST>::exp(m - M);
  const float S = Grp<GROUP>::sum(sc, sm);
  lnS = Mth<FAST>::log(S);
}

struct RowConst {
  float Ma, lnSa, Mb, lnSb;
  float ln_keep, ln_mask;                 // absorbing
  float d_alpha, c0, zc_a, zc_b;          // uniform numerators
  float den_a, den_b;                     // uniform denominators
  float mix_shift;                        // ABS_CFG: -lse(mix)
  int z;
};

// log of the unnormalised guided posterior at column j
template <int MODE, bool FAST>
__device__ __forceinline__ float g_value(const ddg_sample_step_args& a, const RowConst& rc,
                                         int j, float la, float lb, float clf) {
  using F = Mth<FAST>;
  const bool rnd = a.round_bf16 != 0;
  float lpa = F::add(F::add(la, -rc.Ma), -rc.lnSa);
  if (rnd) lpa = bf16_round(lpa);
  float lpb = 0.f;
  if (mode_two(MODE)) {
    lpb = F::add(F::add(lb, -rc.Mb), -rc.lnSb);
    if (rnd) lpb = bf16_round(lpb);
  }
  if (MODE == AR_NONE) return lpa;
  if (MODE == AR_CFG) return F::add(F::mul(a.gamma, lpa), F::mul(1.f - a.gamma, lpb));
  if (MODE == ABS_NONE) {
    if (j == a.mask_index) return rc.ln_mask;
    if (rnd) return F::log(bf16_round(F::exp(lpa))) + rc.ln_keep;
    return lpa + rc.ln_keep;
  }
  if (MODE == ABS_CFG) {
    // mixture of log-probs; its own normaliser arrives through rc.mix_shift
    float ta = F::mul(a.gamma, lpa), tb = F::mul(1.f - a.gamma, lpb);
    if (rnd) { ta = bf16_round(ta); tb = bf16_round(tb); }
    float mix = F::add(ta, tb);
    if (rnd) mix = bf16_round(mix);
    return mix;   // caller adds mix_shift / ln_keep (needs the lse of mix)
  }
  // gamma * classifier_log_prob is a bf16 product in the reference when the exact
  // classifier term (a bf16 tensor on a GPU) is used (diffusion.py:1607, 1612)
  float gc = 0.f;
  if (mode_clf(MODE)) {
    gc = F::mul(a.gamma, clf);
    if (a.clf_round_bf16) gc = bf16_round(gc);
  }
  if (MODE == ABS_CBG) {
    const float lq = (j == a.mask_index) ? rc.ln_mask : F::add(lpa, rc.ln_keep);
    return F::add(gc, lq);
  }
  // uniform: q_j = (d_alpha x_j + c0 + [j = z] zc) / den     diffusion.py:383-392
  float xa = F::exp(lpa);
  if (rnd) xa = bf16_round(xa);
  float na = (j == rc.z) ? F::add(F::add(rc.zc_a, F::mul(rc.d_alpha, xa)), rc.c0)
                         : F::add(F::mul(rc.d_alpha, xa), rc.c0);
  const float lqa = F::log(F::div(na, rc.den_a));
  if (MODE == UNI_NONE) return lqa;
  if (MODE == UNI_CBG) return F::add(gc, lqa);
  float xb = F::exp(lpb);
  if (rnd) xb = bf16_round(xb);
  float nb = (j == rc.z) ? F::add(F::add(rc.zc_b, F::mul(rc.d_alpha, xb)), rc.c0)
                         : F::add(F::mul(rc.d_alpha, xb), rc.c0);
  const float lqb = F::log(F::div(nb, rc.den_b));
  return F::add(F::mul(a.gamma, lqa), F::mul(1.f - a.gamma, lqb));
}

// ABS_CFG second stage: normalised mixture -> absorbing posterior (log)
template <bool FAST>
__device__ __forceinline__ float abs_cfg_finish(const ddg_sample_step_args& a, const RowConst& rc,
                                                int j, float mix) {
  using F = Mth<FAST>;
  if (j == a.mask_index) return rc.ln_mask;
  float lx = mix + rc.mix_shift;   // log softmax(mix)_j
  if (a.round_bf16) lx = F::log(bf16_round(F::exp(lx)));
  return lx + rc.ln_keep;
}

template <typename T, int VEC, int GROUP, int MODE, bool FAST>
__global__ void __launch_bounds__(GROUP == 32 ? 256 : GROUP, GROUP == 32 ? 4 : 4)
sample_step_kernel(const ddg_sample_step_args a) {
  using F = Mth<FAST>;
  __shared__ float sm_red[(GROUP > 32) ? 2 * (GROUP / 32) : 1];
  float* sm = sm_red;

  const int lane_in_group = (GROUP == 32) ? (threadIdx.x & 31) : threadIdx.x;
  const int64_t groups_per_cta = (GROUP == 32) ? (blockDim.x >> 5) : 1;
  const int64_t group0 = (GROUP == 32) ? (int64_t)blockIdx.x * groups_per_cta + (threadIdx.x >> 5)
                                       : (int64_t)blockIdx.x;
  const int64_t group_stride = (int64_t)gridDim.x * groups_per_cta;
  const int V = a.vocab;
  const int step = a.step_ptr ? *a.step_ptr : a.step;

  for (int64_t row = group0; row < a.rows; row += group_stride) {
    const int z = (int)a.xt[row];
    // ---- carry-over of unmasked tokens (diffusion.py:1414-1418, 1606-1621)
    if (mode_abs(MODE) && z != a.mask_index) {
      if (lane_in_group == 0) a.xs[row] = z;
      if (a.q_out) {
        float* q = a.q_out + row * a.ld_q;
        for (int j = lane_in_group; j < V; j += GROUP) q[j] = (j == z) ? 1.f : 0.f;
      }
      continue;
    }
    const int64_t seq = row / a.seq_len;
    const int64_t mc_off = seq * a.mc_seq_stride + (int64_t)step * a.mc_step_stride;
    const float mc_t = a.mc_t ? a.mc_t[mc_off] : 0.f;
    const float mc_s = a.mc_s ? a.mc_s[mc_off] : 0.f;

    const T* pa = reinterpret_cast<const T*>(a.logits_a) + row * a.ld_a;
    const T* pb = mode_two(MODE) ? reinterpret_cast<const T*>(a.logits_b) + row * a.ld_b : nullptr;
    const float* pc = mode_clf(MODE) ? a.clf + row * a.ld_clf : nullptr;
    const float* pn = a.noise ? a.noise + row * a.ld_noise : nullptr;

    RowConst rc;
    rc.z = z;
    // ---------------- pass 1: log-softmax normalisers
    // Throughput mode with the vocab GEMM's per-128-column (max, sum exp(v - max)) partials
    // at hand (ddg_gemm_args.stats): combine those instead of walking the row -- the D-CBG
    // and other guided modes this generic kernel serves then read each logit once.
    const bool have_stats = FAST && a.stats_a != nullptr && a.stats_chunks == (V + 127) / 128 &&
                            (!mode_two(MODE) || a.stats_b != nullptr);
    if (have_stats) {
      LseAcc<FAST> acc_a, acc_b;
      const float2* sa = reinterpret_cast<const float2*>(a.stats_a) + row;
      const float2* sb = mode_two(MODE) ? reinterpret_cast<const float2*>(a.stats_b) + row : nullptr;
      for (int ch = lane_in_group; ch < a.stats_chunks; ch += GROUP) {
        const float2 pa2 = sa[(int64_t)ch * a.stats_ld];
        if (pa2.x > -INFINITY) { acc_a.push_max(pa2.x); acc_a.s += pa2.y * F::exp(pa2.x - acc_a.m); }
        if (mode_two(MODE)) {
          const float2 pb2 = sb[(int64_t)ch * a.stats_ld];
          if (pb2.x > -INFINITY) { acc_b.push_max(pb2.x); acc_b.s += pb2.y * F::exp(pb2.x - acc_b.m); }
        }
      }
      lse_finish<GROUP, FAST>(acc_a.m, acc_a.s, sm, rc.Ma, rc.lnSa);
      if (mode_two(MODE)) lse_finish<GROUP, FAST>(acc_b.m, acc_b.s, sm, rc.Mb, rc.lnSb);
      else { rc.Mb = 0.f; rc.lnSb = 0.f; }
    } else {
      LseAcc<FAST> acc_a, acc_b;
      for (int j0 = lane_in_group * VEC; j0 < V; j0 += GROUP * VEC) {
        float va[VEC], vb[VEC];
        RowLoad<T, VEC>::ld(pa + j0, va);
        if (mode_two(MODE)) RowLoad<T, VEC>::ld(pb + j0, vb);
        float ma = -INFINITY, mb = -INFINITY;
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          const int j = j0 + e;
          if (a.mask_logit && j == a.mask_index) { va[e] += kNegInf; if (mode_two(MODE)) vb[e] += kNegInf; }
          if (j >= V) { va[e] = -INFINITY; if (mode_two(MODE)) vb[e] = -INFINITY; }
          ma = fmaxf(ma, va[e]);
          if (mode_two(MODE)) mb = fmaxf(mb, vb[e]);
        }
        acc_a.push_max(ma);
        if (mode_two(MODE)) acc_b.push_max(mb);
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          acc_a.add(va[e]);
          if (mode_two(MODE)) acc_b.add(vb[e]);
        }
      }
      lse_finish<GROUP, FAST>(acc_a.m, acc_a.s, sm, rc.Ma, rc.lnSa);
      if (mode_two(MODE)) lse_finish<GROUP, FAST>(acc_b.m, acc_b.s, sm, rc.Mb, rc.lnSb);
      else { rc.Mb = 0.f; rc.lnSb = 0.f; }
    }
    // ---------------- row constants
    if (mode_abs(MODE)) {
      const float r = F::div(mc_s, mc_t);
      rc.ln_keep = F::log(1.f - r);
      rc.ln_mask = F::log(r);
    }
    if (mode_uni(MODE)) {
      const float alpha_s = 1.f - mc_s, alpha_t = 1.f - mc_t;
      const float alpha_ts = F::div(alpha_t, alpha_s);
      rc.d_alpha = alpha_s - alpha_t;
      const float limiting = __fdiv_rn(1.f, (float)V);
      rc.c0 = F::mul(F::mul(1.f - alpha_ts, 1.f - alpha_s), limiting);
      const float atV = F::mul(alpha_t, (float)V);
      float lz = Cvt<T>::to_f32(pa[z]);
      if (a.mask_logit && z == a.mask_index) lz += kNegInf;
      float lpz = F::add(F::add(lz, -rc.Ma), -rc.lnSa);
      if (a.round_bf16) lpz = bf16_round(lpz);
      float xz = F::exp(lpz);
      if (a.round_bf16) xz = bf16_round(xz);
      rc.zc_a = F::add(F::mul(atV, xz), alpha_ts - alpha_t);
      rc.den_a = F::add(F::mul(atV, xz), 1.f - alpha_t);
      if (mode_two(MODE)) {
        float lzb = Cvt<T>::to_f32(pb[z]);
        if (a.mask_logit && z == a.mask_index) lzb += kNegInf;
        float lpzb = F::add(F::add(lzb, -rc.Mb), -rc.lnSb);
        if (a.round_bf16) lpzb = bf16_round(lpzb);
        float xzb = F::exp(lpzb);
        if (a.round_bf16) xzb = bf16_round(xzb);
        rc.zc_b = F::add(F::mul(atV, xzb), alpha_ts - alpha_t);
        rc.den_b = F::add(F::mul(atV, xzb), 1.f - alpha_t);
      }
    }
    // ---------------- pass 1b (ABS_CFG only): normaliser of the mixture
    if (MODE == ABS_CFG) {
      LseAcc<FAST> acc;
      for (int j0 = lane_in_group * VEC; j0 < V; j0 += GROUP * VEC) {
        float va[VEC], vb[VEC], mx[VEC];
        RowLoad<T, VEC>::ld(pa + j0, va);
        RowLoad<T, VEC>::ld(pb + j0, vb);
        float vm = -INFINITY;
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          const int j = j0 + e;
          if (a.mask_logit && j == a.mask_index) { va[e] += kNegInf; vb[e] += kNegInf; }
          mx[e] = (j < V) ? g_value<MODE, FAST>(a, rc, j, va[e], vb[e], 0.f) : -INFINITY;
          vm = fmaxf(vm, mx[e]);
        }
        acc.push_max(vm);
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc.add(mx[e]);
      }
      float M, lnS;
      lse_finish<GROUP, FAST>(acc.m, acc.s, sm, M, lnS);
      rc.mix_shift = -(M + lnS);
    }
    // ---------------- pass 2: selection (+ normaliser of g when q_out)
    float best = -INFINITY;
    int bidx = 0x7fffffff;
    LseAcc<FAST> gacc;
    const uint64_t seed64 = a.seed_ptr ? *a.seed_ptr : a.seed;
    const uint2 key = make_uint2((uint32_t)seed64, (uint32_t)(seed64 >> 32));
    for (int j0 = lane_in_group * VEC; j0 < V; j0 += GROUP * VEC) {
      float va[VEC], vb[VEC], g[VEC];
      RowLoad<T, VEC>::ld(pa + j0, va);
      if (mode_two(MODE)) RowLoad<T, VEC>::ld(pb + j0, vb);
      float u[VEC];
      if (a.select != DDG_SELECT_ARGMAX) {
        if (pn) {
#pragma unroll
          for (int e = 0; e < VEC; ++e) u[e] = (j0 + e < V) ? __ldg(pn + j0 + e) : 0.5f;
        } else {
#pragma unroll
          for (int e0 = 0; e0 < VEC; e0 += 4) {
            const uint4 w = philox4x32_10(
              make_uint4((uint32_t)((j0 + e0) >> 2), (uint32_t)row, (uint32_t)(row >> 32), (uint32_t)step), key);
            const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
            if (VEC >= 4) {
#pragma unroll
              for (int e = 0; e < 4; ++e) u[e0 + e] = u01_from_bits(ww[e]);
            } else {
              u[0] = u01_from_bits(ww[(j0 + e0) & 3]);
            }
          }
        }
      }
      float vm = -INFINITY;
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        const int j = j0 + e;
        if (a.mask_logit && j == a.mask_index) { va[e] += kNegInf; if (mode_two(MODE)) vb[e] += kNegInf; }
        if (j < V) {
          const float cf = mode_clf(MODE) ? __ldg(pc + j) : 0.f;
          float gj = g_value<MODE, FAST>(a, rc, j, va[e], mode_two(MODE) ? vb[e] : 0.f, cf);
          if (MODE == ABS_CFG) gj = abs_cfg_finish<FAST>(a, rc, j, gj);
          g[e] = gj;
          float score = gj;
          if (a.select == DDG_SELECT_RACE || (a.select == DDG_SELECT_GUMBEL && !pn)) {
            const float E = F::add(1e-10f, -F::log(F::add(u[e], 1e-10f)));
            score = gj - F::log(E);
          } else if (a.select == DDG_SELECT_GUMBEL) {
            score = gj + u[e];
          }
          if (score > best) { best = score; bidx = j; }
        } else {
          g[e] = -INFINITY;
        }
        vm = fmaxf(vm, g[e]);
      }
      if (a.q_out && mode_softmax(MODE)) {
        gacc.push_max(vm);
#pragma unroll
        for (int e = 0; e < VEC; ++e) gacc.add(g[e]);
      }
    }
    Grp<GROUP>::argmax(best, bidx, sm);
    if (bidx == 0x7fffffff) bidx = 0;   // every score was -inf / NaN
    if (lane_in_group == 0) {
      a.xs[row] = (int64_t)bidx;
      if (a.seq_changed && bidx != z) a.seq_changed[seq] = 1;
    }
    // ---------------- pass 3 (debug / API): q_xs as the reference returns it
    if (a.q_out) {
      float lse = 0.f;
      if (mode_softmax(MODE)) {
        float M, lnS;
        lse_finish<GROUP, FAST>(gacc.m, gacc.s, sm, M, lnS);
        lse = M + lnS;
      }
      float* q = a.q_out + row * a.ld_q;
      for (int j0 = lane_in_group * VEC; j0 < V; j0 += GROUP * VEC) {
        float va[VEC], vb[VEC];
        RowLoad<T, VEC>::ld(pa + j0, va);
        if (mode_two(MODE)) RowLoad<T, VEC>::ld(pb + j0, vb);
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          const int j = j0 + e;
          if (j >= V) continue;
          if (a.mask_logit && j == a.mask_index) { va[e] += kNegInf; if (mode_two(MODE)) vb[e] += kNegInf; }
          const float cf = mode_clf(MODE) ? __ldg(pc + j) : 0.f;
          float gj = g_value<MODE, FAST>(a, rc, j, va[e], mode_two(MODE) ? vb[e] : 0.f, cf);
          if (MODE == ABS_CFG) gj = abs_cfg_finish<FAST>(a, rc, j, gj);
          q[j] = F::exp(gj - lse);
        }
      }
    }
  }
}

template <typename T, int VEC, int GROUP, int MODE, bool FAST>
int launch(const ddg_sample_step_args& a, cudaStream_t s) {
  const int sms = sm_count();
  int64_t grid;
  int block;
  if (GROUP == 32) {
    block = 256;
    const int64_t ctas = (a.rows + 7) / 8;
    grid = ctas < (int64_t)sms * 8 ? ctas : (int64_t)sms * 8;
  } else {
    block = GROUP;
    grid = a.rows < (int64_t)sms * 4 ? a.rows : (int64_t)sms * 4;
  }
  if (grid < 1) grid = 1;
  sample_step_kernel<T, VEC, GROUP, MODE, FAST><<<(unsigned)grid, block, 0, s>>>(a);
  return check_cuda(cudaGetLastError(), "sample_step launch");
}

template <typename T, int VEC, int GROUP, int MODE>
int launch_fast(const ddg_sample_step_args& a, cudaStream_t s) {
  if (a.fast_math) {
    if (VEC == 1) {
      // the scalar-load fallback is a parity / odd-alignment path: keep one variant
      return launch<T, VEC, GROUP, MODE, false>(a, s);
    }
    return launch<T, VEC, GROUP, MODE, true>(a, s);
  }
  return launch<T, VEC, GROUP, MODE, false>(a, s);
}

template <typename T, int VEC, int GROUP>
int launch_mode(int mode, const ddg_sample_step_args& a, cudaStream_t s) {
  switch (mode) {
    case ABS_NONE: return launch_fast<T, VEC, GROUP, ABS_NONE>(a, s);
    case ABS_CFG: return launch_fast<T, VEC, GROUP, ABS_CFG>(a, s);
    case ABS_CBG: return launch_fast<T, VEC, GROUP, ABS_CBG>(a, s);
    case UNI_NONE: return launch_fast<T, VEC, GROUP, UNI_NONE>(a, s);
    case UNI_CFG: return launch_fast<T, VEC, GROUP, UNI_CFG>(a, s);
    case UNI_CBG: return launch_fast<T, VEC, GROUP, UNI_CBG>(a, s);
    case AR_NONE: return launch_fast<T, VEC, GROUP, AR_NONE>(a, s);
    case AR_CFG: return launch_fast<T, VEC, GROUP, AR_CFG>(a, s);
  }
  set_error("sample_step: bad mode %d", mode);
  return DDG_ERR_INVALID;
}

template <typename T, int VECN>
int launch_group(int mode, const ddg_sample_step_args& a, bool vec_ok, cudaStream_t s) {
  const bool big = a.vocab >= 2048;
  if (vec_ok) {
    return big ? launch_mode<T, VECN, 256>(mode, a, s) : launch_mode<T, VECN, 32>(mode, a, s);
  }
  return big ? launch_mode<T, 1, 256>(mode, a, s) : launch_mode<T, 1, 32>(mode, a, s);
}

// Standalone log-probs (API / parity use only): one CTA per row.
template <typename T>
__global__ void __launch_bounds__(256)
log_probs_kernel(const T* __restrict__ logits, int64_t ld, const int64_t* __restrict__ xt, int64_t rows, int V,
                 int subs, int mask_logit, int mask_index, int round_bf16, float* __restrict__ out,
                 int64_t ld_out) {
  __shared__ float sm[16];
  for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
    const T* p = logits + row * ld;
    float* o = out + row * ld_out;
    const int z = xt ? (int)xt[row] : -1;
    if (subs && z != mask_index) {
      // carry-over row: -1e6 everywhere, 0 at z, then log_softmax (diffusion.py:330-337)
      for (int j = threadIdx.x; j < V; j += 256) o[j] = (j == z) ? 0.f : kNegInf;
      continue;
    }
    LseAcc<false> acc;
    for (int j = threadIdx.x; j < V; j += 256) {
      float v = Cvt<T>::to_f32(p[j]);
      if (mask_logit && j == mask_index) v += kNegInf;
      acc.push_max(v);
      acc.add(v);
    }
    float M, lnS;
    lse_finish<256, false>(acc.m, acc.s, sm, M, lnS);
    for (int j = threadIdx.x; j < V; j += 256) {
      float v = Cvt<T>::to_f32(p[j]);
      if (mask_logit && j == mask_index) v += kNegInf;
      float lp = __fadd_rn(__fadd_rn(v, -M), -lnS);
      if (round_bf16) lp = bf16_round(lp);
      o[j] = lp;
    }
  }
}

__global__ void prior_kernel(int64_t* xt, int64_t n, int uniform, int vocab, int mask_index, uint64_t seed) {
  const int64_t i4 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i4 * 4 >= n) return;
  uint32_t ww[4] = {0, 0, 0, 0};
  if (uniform) {
    const uint4 w = philox4x32_10(make_uint4((uint32_t)i4, (uint32_t)(i4 >> 32), 0u, 0xFFFFFFFFu),
                                  make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    ww[0] = w.x; ww[1] = w.y; ww[2] = w.z; ww[3] = w.w;
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int64_t i = i4 * 4 + e;
    if (i < n) xt[i] = uniform ? (int64_t)(((uint64_t)ww[e] * (uint64_t)vocab) >> 32) : (int64_t)mask_index;
  }
}

__global__ void step_advance_kernel(int32_t* p) { *p += 1; }

}  // namespace
}  // namespace ddg

extern "C" int ddg_sample_step(const ddg_sample_step_args* args, ddg_stream_t stream) {
  using namespace ddg;
  DDG_REQUIRE(args != nullptr, DDG_ERR_INVALID, "sample_step: null args");
  const ddg_sample_step_args& a = *args;
  int rc = ddg_device_check();
  if (rc) return rc;
  DDG_REQUIRE(a.rows >= 0 && a.vocab > 0 && a.seq_len > 0, DDG_ERR_INVALID,
              "sample_step: bad sizes rows=%lld vocab=%d seq_len=%d", (long long)a.rows, a.vocab, a.seq_len);
  if (a.rows == 0) return DDG_OK;
  DDG_REQUIRE(a.logits_a && a.xt && a.xs, DDG_ERR_INVALID, "sample_step: null logits_a/xt/xs");
  DDG_REQUIRE(a.ld_a >= a.vocab, DDG_ERR_INVALID, "sample_step: ld_a < vocab");
  DDG_REQUIRE(a.logits_dtype == DDG_F32 || a.logits_dtype == DDG_BF16, DDG_ERR_INVALID,
              "sample_step: logits dtype must be f32 or bf16");
  int mode;
  const bool ar = a.diffusion == DDG_AR;
  if (a.diffusion == DDG_ABSORBING) mode = ABS_NONE;
  else if (a.diffusion == DDG_UNIFORM) mode = UNI_NONE;
  else if (ar) mode = AR_NONE;
  else { set_error("Diffusion type %d not implemented.", a.diffusion); return DDG_ERR_INVALID; }
  if (a.guidance == DDG_GUIDE_CFG) {
    DDG_REQUIRE(a.logits_b && a.ld_b >= a.vocab, DDG_ERR_INVALID, "sample_step: cfg needs logits_b");
    mode += 1;
  } else if (a.guidance == DDG_GUIDE_CBG) {
    DDG_REQUIRE(!ar, DDG_ERR_INVALID, "sample_step: cbg is a diffusion guidance");
    DDG_REQUIRE(a.clf && a.ld_clf >= a.vocab, DDG_ERR_INVALID, "sample_step: cbg needs clf");
    mode += 2;
  } else if (a.guidance != DDG_GUIDE_NONE) {
    set_error("Guidance method %d not implemented.", a.guidance);
    return DDG_ERR_INVALID;
  }
  if (!ar) DDG_REQUIRE(a.mc_t && a.mc_s, DDG_ERR_INVALID, "sample_step: null move chances");
  if (a.diffusion == DDG_ABSORBING || a.mask_logit)
    DDG_REQUIRE(a.mask_index >= 0 && a.mask_index < a.vocab, DDG_ERR_INVALID,
                "sample_step: mask_index %d outside vocab %d", a.mask_index, a.vocab);
  DDG_REQUIRE(a.select >= DDG_SELECT_RACE && a.select <= DDG_SELECT_GUMBEL, DDG_ERR_INVALID,
              "sample_step: bad select %d", a.select);
  if (a.noise) DDG_REQUIRE(a.ld_noise >= a.vocab, DDG_ERR_INVALID, "sample_step: ld_noise < vocab");
  if (a.q_out) DDG_REQUIRE(a.ld_q >= a.vocab, DDG_ERR_INVALID, "sample_step: ld_q < vocab");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const bool two = mode_two(mode);
  if (a.compute_f64) return launch_sample_f64(a, s);
  {
    // single-pass production kernel when the GEMM supplied its softmax partials
    const int took = try_sample_fused(a, s);
    if (took != 0) return took < 0 ? took : DDG_OK;
  }
  if (a.logits_dtype == DDG_BF16) {
    const int64_t vpad = (a.vocab + 7) / 8 * 8;
    bool ok = (reinterpret_cast<uintptr_t>(a.logits_a) % 16 == 0) && a.ld_a % 8 == 0 && a.ld_a >= vpad;
    if (two) ok = ok && (reinterpret_cast<uintptr_t>(a.logits_b) % 16 == 0) && a.ld_b % 8 == 0 && a.ld_b >= vpad;
    return launch_group<__nv_bfloat16, 8>(mode, a, ok, s);
  }
  const int64_t vpad = (a.vocab + 3) / 4 * 4;
  bool ok = (reinterpret_cast<uintptr_t>(a.logits_a) % 16 == 0) && a.ld_a % 4 == 0 && a.ld_a >= vpad;
  if (two) ok = ok && (reinterpret_cast<uintptr_t>(a.logits_b) % 16 == 0) && a.ld_b % 4 == 0 && a.ld_b >= vpad;
  return launch_group<float, 4>(mode, a, ok, s);
}

extern "C" int ddg_log_probs(const void* logits, int64_t ld, int32_t logits_dtype, const int64_t* xt, int64_t rows,
                             int32_t vocab, int32_t subs, int32_t mask_logit, int32_t mask_index,
                             int32_t round_bf16, float* out, int64_t ld_out, ddg_stream_t stream) {
  using namespace ddg;
  int rc = ddg_device_check();
  if (rc) return rc;
  DDG_REQUIRE(logits && out && rows >= 0 && vocab > 0 && ld >= vocab && ld_out >= vocab, DDG_ERR_INVALID,
              "log_probs: bad args");
  if (subs) DDG_REQUIRE(xt != nullptr, DDG_ERR_INVALID, "log_probs: subs needs xt");
  if (rows == 0) return DDG_OK;
  const int64_t cap = (int64_t)sm_count() * 8;
  const unsigned grid = (unsigned)(rows < cap ? rows : cap);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (logits_dtype == DDG_BF16)
    log_probs_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(logits), ld, xt, rows,
                                                         vocab, subs, mask_logit, mask_index, round_bf16, out, ld_out);
  else if (logits_dtype == DDG_F32)
    log_probs_kernel<float><<<grid, 256, 0, s>>>(reinterpret_cast<const float*>(logits), ld, xt, rows, vocab, subs,
                                                 mask_logit, mask_index, round_bf16, out, ld_out);
  else {
    set_error("log_probs: logits dtype must be f32 or bf16");
    return DDG_ERR_INVALID;
  }
  return check_cuda(cudaGetLastError(), "log_probs launch");
}

extern "C" int ddg_sample_prior(int64_t* xt, int64_t n, int32_t diffusion, int32_t vocab,
                                int32_t mask_index, uint64_t seed, ddg_stream_t stream) {
  using namespace ddg;
  int rc = ddg_device_check();
  if (rc) return rc;
  DDG_REQUIRE(xt && n >= 0 && vocab > 0, DDG_ERR_INVALID, "sample_prior: bad args");
  DDG_REQUIRE(diffusion == DDG_ABSORBING || diffusion == DDG_UNIFORM, DDG_ERR_INVALID,
              "Diffusion type %d not implemented.", diffusion);
  if (n == 0) return DDG_OK;
  const int64_t n4 = (n + 3) / 4;
  prior_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
    xt, n, diffusion == DDG_UNIFORM, vocab, mask_index, seed);
  return check_cuda(cudaGetLastError(), "sample_prior launch");
}

extern "C" int ddg_step_advance(int32_t* step_ptr, ddg_stream_t stream) {
  using namespace ddg;
  int rc = ddg_device_check();
  if (rc) return rc;
  DDG_REQUIRE(step_ptr, DDG_ERR_INVALID, "step_advance: null");
  step_advance_kernel<<<1, 1, 0, reinterpret_cast<cudaStream_t>(stream)>>>(step_ptr);
  return check_cuda(cudaGetLastError(), "step_advance launch");
}
