// Single-pass fused sampler (production path).
//
// Same mathematics as sample_step.cu (reference diffusion.py:356-394,
// 1393-1420, 1456-1509, 1748-1755), specialised for what the DiT head emits:
// bf16 logits in 16-byte-aligned rows plus the log-softmax partials that the
// vocab GEMM epilogue already produced (ddg_gemm_args.stats).  With the row
// normaliser known up front, one CTA walks a token row exactly once: every
// logit is read from HBM one time and nothing is staged or re-read.
//
// All transcendental work is in base 2 (ex2/lg2.approx), selection scores are
//     s_j = log2 q~_j - log2 E_j,      E_j = 1e-10 - ln(U_j + 1e-10)
// (the reference's p_j / E_j race in the log domain; common factors dropped).
// Modes: UDLM posterior (null / D-CFG / D-CBG with the fp32 classifier term of
// diffusion.py:1587-1621 read alongside the logits) and MDLM absorbing posterior (null).
#include "ddg_common.cuh"

namespace ddg {
namespace {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr int kThreads = 256;

enum FMode : int { F_UNI_NONE = 0, F_UNI_CFG = 1, F_ABS_NONE = 2, F_UNI_CBG = 3 };

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
// bf16 round trip of a pair with one packed conversion
__device__ __forceinline__ void bf16_round2(float& a, float& b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  const uint32_t w = *reinterpret_cast<const uint32_t*>(&h);
  a = __uint_as_float(w << 16);
  b = __uint_as_float(w & 0xffff0000u);
}
__device__ __forceinline__ void unpack8(const uint4& r, float (&v)[8]) {
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(w[i] << 16);
    v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ uint4 ldg_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

struct FusedConst {
  float Ka, Kb;            // natural-log lse of the two rows
  float dA, c0;            // uniform: q~_j = dA x_j + c0 (+ zc at j = z)
  float zc_a, zc_b;
  float gam, omg;          // gamma, 1 - gamma
  bool clf_round;          // D-CBG: gamma * clf is a bf16 product (exact classifier term on a GPU)
  float mask_score;        // absorbing: log2(r / (1 - r)) of the mask column
  int z, mask_index, V;
};

// log2 of the unnormalised (guided) posterior at one column.  `extra` selects
// the j == z (uniform) handling; the absorbing mask column is handled by the
// caller.
// D-CBG: gamma * classifier term in log2 units (diffusion.py:1607, 1612)
__device__ __forceinline__ float clf2(const FusedConst& c, float cf) {
  float gc = c.gam * cf;
  if (c.clf_round) gc = bf16_round(gc);
  return gc * kLog2e;
}

template <int MODE, bool ROUND>
__device__ __forceinline__ float g2_one(const FusedConst& c, float la, float lb, bool is_z) {
  float lpa = la - c.Ka;
  if (ROUND) lpa = bf16_round(lpa);
  if (MODE == F_ABS_NONE) {
    if (!ROUND) return lpa * kLog2e;
    return lg2a(bf16_round(ex2a(lpa * kLog2e)));
  }
  float xa = ex2a(lpa * kLog2e);
  if (ROUND) xa = bf16_round(xa);
  float na = fmaf(c.dA, xa, c.c0);
  if (is_z) na += c.zc_a;
  const float lqa = lg2a(na);
  if (MODE == F_UNI_NONE) return lqa;
  if (MODE == F_UNI_CBG) return lqa + clf2(c, lb);      // lb carries the classifier term
  float lpb = lb - c.Kb;
  if (ROUND) lpb = bf16_round(lpb);
  float xb = ex2a(lpb * kLog2e);
  if (ROUND) xb = bf16_round(xb);
  float nb = fmaf(c.dA, xb, c.c0);
  if (is_z) nb += c.zc_b;
  return fmaf(c.gam, lqa, c.omg * lg2a(nb));
}

__device__ __forceinline__ float neg_log2_E(float u) {
  // E = 1e-10 - ln(u + 1e-10);  returns -log2(E)
  const float t = lg2a(u + 1e-10f);
  return -lg2a(fmaf(-kLn2, t, 1e-10f));
}

template <int MODE, bool ROUND, bool INJ>
__global__ void __launch_bounds__(kThreads, 4)
sample_fused_kernel(const ddg_sample_step_args a) {
  __shared__ float s_k[2];
  __shared__ float s_red[kThreads / 32];
  __shared__ int s_idx[kThreads / 32];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int V = a.vocab;
  const int step = a.step_ptr ? *a.step_ptr : a.step;
  const uint64_t seed64 = a.seed_ptr ? *a.seed_ptr : a.seed;
  const uint2 key = make_uint2((uint32_t)seed64, (uint32_t)(seed64 >> 32));
  const float2* st_a = reinterpret_cast<const float2*>(a.stats_a);
  const float2* st_b = reinterpret_cast<const float2*>(a.stats_b);

  for (int64_t row = blockIdx.x; row < a.rows; row += gridDim.x) {
    const int z = (int)a.xt[row];
    if (MODE == F_ABS_NONE && z != a.mask_index) {      // carry-over (diffusion.py:1414-1418)
      if (tid == 0) a.xs[row] = z;
      continue;
    }
    const __nv_bfloat16* pa = reinterpret_cast<const __nv_bfloat16*>(a.logits_a) + row * a.ld_a;
    const __nv_bfloat16* pb =
      MODE == F_UNI_CFG ? reinterpret_cast<const __nv_bfloat16*>(a.logits_b) + row * a.ld_b : nullptr;
    // first vectors in flight while warp 0 folds the partials
    const int jfirst = tid * 8;
    // D-CBG: the row of the fp32 classifier term (16-byte aligned, padded to whole vectors of 8)
    const float* pc = MODE == F_UNI_CBG ? a.clf + row * a.ld_clf : nullptr;
    uint4 ra = make_uint4(0, 0, 0, 0), rb = make_uint4(0, 0, 0, 0);
    uint4 rc0 = make_uint4(0, 0, 0, 0), rc1 = make_uint4(0, 0, 0, 0);
    if (jfirst < V) {
      ra = ldg_stream(pa + jfirst);
      if (MODE == F_UNI_CFG) rb = ldg_stream(pb + jfirst);
      if (MODE == F_UNI_CBG) { rc0 = ldg_stream(pc + jfirst); rc1 = ldg_stream(pc + jfirst + 4); }
    }
    // ---- row normalisers from the GEMM partials (warps 0 and 1)
    if (warp < (MODE == F_UNI_CFG ? 2 : 1)) {
      const float2* st = (warp == 0 ? st_a : st_b) + row;
      float m = -INFINITY, s = 0.f;
      for (int c = lane; c < a.stats_chunks; c += 32) {
        const float2 pr = st[(int64_t)c * a.stats_ld];
        const float nm = fmaxf(m, pr.x);
        s = s * ex2a((m - nm) * kLog2e) + pr.y * ex2a((pr.x - nm) * kLog2e);
        m = nm;
      }
      const float M = warp_max(m);
      s = (m == -INFINITY) ? 0.f : s * ex2a((m - M) * kLog2e);
      const float S = warp_sum(s);
      if (lane == 0) s_k[warp] = M + kLn2 * lg2a(S);
    }
    __syncthreads();
    FusedConst c;
    c.Ka = s_k[0];
    c.Kb = MODE == F_UNI_CFG ? s_k[1] : 0.f;
    c.z = z; c.mask_index = a.mask_index; c.V = V;
    c.gam = a.gamma; c.omg = 1.f - a.gamma;
    c.clf_round = a.clf_round_bf16 != 0;
    const int64_t mc_off = (row / a.seq_len) * a.mc_seq_stride + (int64_t)step * a.mc_step_stride;
    const float mc_t = a.mc_t[mc_off], mc_s = a.mc_s[mc_off];
    int special;                                        // column needing the careful path
    if (MODE == F_ABS_NONE) {
      const float r = mc_s / mc_t;
      c.mask_score = lg2a(r) - lg2a(1.f - r);
      special = a.mask_index;
    } else {
      const float alpha_s = 1.f - mc_s, alpha_t = 1.f - mc_t;
      const float alpha_ts = alpha_t / alpha_s;
      c.dA = alpha_s - alpha_t;
      c.c0 = (1.f - alpha_ts) * (1.f - alpha_s) * (1.f / (float)V);
      const float atV = alpha_t * (float)V;
      float lz = __bfloat162float(pa[z]) - c.Ka;
      if (ROUND) lz = bf16_round(lz);
      float xz = ex2a(lz * kLog2e);
      if (ROUND) xz = bf16_round(xz);
      c.zc_a = atV * xz + (alpha_ts - alpha_t);
      if (MODE == F_UNI_CFG) {
        float lzb = __bfloat162float(pb[z]) - c.Kb;
        if (ROUND) lzb = bf16_round(lzb);
        float xzb = ex2a(lzb * kLog2e);
        if (ROUND) xzb = bf16_round(xzb);
        c.zc_b = atV * xzb + (alpha_ts - alpha_t);
      }
      special = z;
    }
    const int special_vec = special & ~7;
    const float* pn = INJ ? a.noise + row * a.ld_noise : nullptr;

    float best = -INFINITY;
    int bidx = 0x7fffffff;
    for (int j0 = jfirst; j0 < V; j0 += kThreads * 8) {
      float va[8], vb[8], nl[8];
      unpack8(ra, va);
      if (MODE == F_UNI_CFG) unpack8(rb, vb);
      if (MODE == F_UNI_CBG) {                          // vb = the classifier term of these 8 columns
        vb[0] = __uint_as_float(rc0.x); vb[1] = __uint_as_float(rc0.y);
        vb[2] = __uint_as_float(rc0.z); vb[3] = __uint_as_float(rc0.w);
        vb[4] = __uint_as_float(rc1.x); vb[5] = __uint_as_float(rc1.y);
        vb[6] = __uint_as_float(rc1.z); vb[7] = __uint_as_float(rc1.w);
      }
      const int jn = j0 + kThreads * 8;                 // prefetch the next vectors
      if (jn < V) {
        ra = ldg_stream(pa + jn);
        if (MODE == F_UNI_CFG) rb = ldg_stream(pb + jn);
        if (MODE == F_UNI_CBG) { rc0 = ldg_stream(pc + jn); rc1 = ldg_stream(pc + jn + 4); }
      }
      if (INJ) {
#pragma unroll
        for (int e = 0; e < 8; ++e) nl[e] = neg_log2_E(j0 + e < V ? __ldg(pn + j0 + e) : 0.5f);
      } else {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const uint4 w = philox4x32_10(
            make_uint4((uint32_t)((j0 >> 2) + h), (uint32_t)row, (uint32_t)(row >> 32), (uint32_t)step), key);
          nl[4 * h + 0] = neg_log2_E(u01_from_bits(w.x));
          nl[4 * h + 1] = neg_log2_E(u01_from_bits(w.y));
          nl[4 * h + 2] = neg_log2_E(u01_from_bits(w.z));
          nl[4 * h + 3] = neg_log2_E(u01_from_bits(w.w));
        }
      }
      if (j0 != special_vec && j0 + 8 <= V) {
        // ---- hot path: 8 ordinary columns
        {
#pragma unroll
          for (int e = 0; e < 8; e += 2) {
            float g0, g1;
            if (ROUND) {
              float l0 = va[e] - c.Ka, l1 = va[e + 1] - c.Ka;
              bf16_round2(l0, l1);
              float x0 = ex2a(l0 * kLog2e), x1 = ex2a(l1 * kLog2e);
              bf16_round2(x0, x1);
              if (MODE == F_ABS_NONE) {
                g0 = lg2a(x0); g1 = lg2a(x1);
              } else {
                g0 = lg2a(fmaf(c.dA, x0, c.c0)); g1 = lg2a(fmaf(c.dA, x1, c.c0));
                if (MODE == F_UNI_CFG) {
                  float m0 = vb[e] - c.Kb, m1 = vb[e + 1] - c.Kb;
                  bf16_round2(m0, m1);
                  float y0 = ex2a(m0 * kLog2e), y1 = ex2a(m1 * kLog2e);
                  bf16_round2(y0, y1);
                  g0 = fmaf(c.gam, g0, c.omg * lg2a(fmaf(c.dA, y0, c.c0)));
                  g1 = fmaf(c.gam, g1, c.omg * lg2a(fmaf(c.dA, y1, c.c0)));
                }
                if (MODE == F_UNI_CBG) { g0 += clf2(c, vb[e]); g1 += clf2(c, vb[e + 1]); }
              }
            } else {
              const float ka2 = c.Ka * kLog2e;
              if (MODE == F_ABS_NONE) {
                g0 = fmaf(va[e], kLog2e, -ka2);
                g1 = fmaf(va[e + 1], kLog2e, -ka2);
              } else {
                g0 = lg2a(fmaf(c.dA, ex2a(fmaf(va[e], kLog2e, -ka2)), c.c0));
                g1 = lg2a(fmaf(c.dA, ex2a(fmaf(va[e + 1], kLog2e, -ka2)), c.c0));
              }
              if (MODE == F_UNI_CFG) {
                const float kb2 = c.Kb * kLog2e;
                g0 = fmaf(c.gam, g0, c.omg * lg2a(fmaf(c.dA, ex2a(fmaf(vb[e], kLog2e, -kb2)), c.c0)));
                g1 = fmaf(c.gam, g1, c.omg * lg2a(fmaf(c.dA, ex2a(fmaf(vb[e + 1], kLog2e, -kb2)), c.c0)));
              }
              if (MODE == F_UNI_CBG) { g0 += clf2(c, vb[e]); g1 += clf2(c, vb[e + 1]); }
            }
            const float s0 = g0 + nl[e], s1 = g1 + nl[e + 1];
            if (s0 > best) { best = s0; bidx = j0 + e; }
            if (s1 > best) { best = s1; bidx = j0 + e + 1; }
          }
        }
      } else {
        // ---- careful path: the vector holding z / the mask column, or the ragged tail
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int j = j0 + e;
          if (j >= V) continue;
          float g;
          if (MODE == F_ABS_NONE) {
            if (j == a.mask_index) g = c.mask_score;
            else g = g2_one<MODE, ROUND>(c, va[e], 0.f, false);
          } else {
            g = g2_one<MODE, ROUND>(c, va[e], MODE != F_UNI_NONE ? vb[e] : 0.f, j == z);
          }
          const float sc = g + nl[e];
          if (sc > best) { best = sc; bidx = j; }
        }
      }
    }
    // ---- block argmax, first index wins ties
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float os = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
      if (os > best || (os == best && oi < bidx)) { best = os; bidx = oi; }
    }
    if (lane == 0) { s_red[warp] = best; s_idx[warp] = bidx; }
    __syncthreads();
    if (warp == 0) {
      best = lane < kThreads / 32 ? s_red[lane] : -INFINITY;
      bidx = lane < kThreads / 32 ? s_idx[lane] : 0x7fffffff;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float os = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
        if (os > best || (os == best && oi < bidx)) { best = os; bidx = oi; }
      }
      if (lane == 0) {
        const int pick_id = bidx == 0x7fffffff ? 0 : bidx;
        a.xs[row] = pick_id;
        if (a.seq_changed && pick_id != z) a.seq_changed[row / a.seq_len] = 1;
      }
    }
    __syncthreads();
  }
}

template <int MODE, bool ROUND>
int launch_inj(const ddg_sample_step_args& a, cudaStream_t s) {
  int64_t grid = (int64_t)sm_count() * 4 * 4;     // a few waves of resident CTAs
  if (grid > a.rows) grid = a.rows;
  if (a.noise)
    sample_fused_kernel<MODE, ROUND, true><<<(unsigned)grid, kThreads, 0, s>>>(a);
  else
    sample_fused_kernel<MODE, ROUND, false><<<(unsigned)grid, kThreads, 0, s>>>(a);
  return check_cuda(cudaGetLastError(), "sample_fused launch");
}

}  // namespace

int launch_sample_group(int mode, const ddg_sample_step_args& a, cudaStream_t s);   // sample_group.cu

// Returns 1 if the fused single-pass kernel took the call, 0 if the generic
// kernel must, negative on error.
int try_sample_fused(const ddg_sample_step_args& a, cudaStream_t s) {
  if (a.race_group != 0 && a.race_group != 8 && a.race_group != 64) {
    set_error("sample_step: race_group must be 0, 8 or 64 (got %d)", a.race_group);
    return DDG_ERR_INVALID;
  }
  if (!a.stats_a || a.stats_chunks <= 0 || a.logits_dtype != DDG_BF16 || a.q_out) return 0;
  if ((a.clf != nullptr) != (a.guidance == DDG_GUIDE_CBG)) return 0;
  if (a.select != DDG_SELECT_RACE || !a.fast_math || a.vocab < 2048) return 0;
  if (a.stats_chunks != (a.vocab + 127) / 128) return 0;
  const int64_t vpad = (a.vocab + 7) / 8 * 8;
  auto aligned = [&](const void* p, int64_t ld) {
    return p && reinterpret_cast<uintptr_t>(p) % 16 == 0 && ld % 8 == 0 && ld >= vpad;
  };
  if (!aligned(a.logits_a, a.ld_a)) return 0;
  int mode;
  if (a.diffusion == DDG_UNIFORM && a.guidance == DDG_GUIDE_NONE && !a.mask_logit) mode = F_UNI_NONE;
  else if (a.diffusion == DDG_UNIFORM && a.guidance == DDG_GUIDE_CFG && !a.mask_logit) {
    if (!a.stats_b || !aligned(a.logits_b, a.ld_b)) return 0;
    mode = F_UNI_CFG;
  } else if (a.diffusion == DDG_ABSORBING && a.guidance == DDG_GUIDE_NONE && a.subs && a.mask_logit)
    mode = F_ABS_NONE;   // stats must have been built with stats_mask_col = mask_index
  else if (a.diffusion == DDG_UNIFORM && a.guidance == DDG_GUIDE_CBG && !a.mask_logit) {
    // the fp32 classifier term is read as two 16-byte vectors per 8 columns: rows padded to whole vectors
    if (reinterpret_cast<uintptr_t>(a.clf) % 16 != 0 || a.ld_clf % 4 != 0 || a.ld_clf < vpad) return 0;
    mode = F_UNI_CBG;
  } else return 0;
  if (a.race_group != 0 && !a.noise && mode != F_UNI_CBG) {
    const int rcg = launch_sample_group(mode, a, s);   // hierarchical race (Philox only)
    return rcg ? rcg : 1;
  }
  int rc;
  switch (mode * 2 + (a.round_bf16 ? 1 : 0)) {
    case 0: rc = launch_inj<F_UNI_NONE, false>(a, s); break;
    case 1: rc = launch_inj<F_UNI_NONE, true>(a, s); break;
    case 2: rc = launch_inj<F_UNI_CFG, false>(a, s); break;
    case 3: rc = launch_inj<F_UNI_CFG, true>(a, s); break;
    case 4: rc = launch_inj<F_ABS_NONE, false>(a, s); break;
    case 5: rc = launch_inj<F_ABS_NONE, true>(a, s); break;
    case 6: rc = launch_inj<F_UNI_CBG, false>(a, s); break;
    default: rc = launch_inj<F_UNI_CBG, true>(a, s); break;
  }
  return rc ? rc : 1;
}

}  // namespace ddg
