// Input-gradient ("dgrad") pieces of the classifier backward that first-order
// D-CBG needs (reference diffusion.py:1532-1548: one classifier forward +
// backward w.r.t. a one-hot input; the classifier is models/dit.py:470-556).
// Only activations' gradients are propagated -- no weight gradients.  The
// linear layers' dgrads are ddg_gemm calls on transposed weights; this file has
// the rest: head, gate, GELU', LayerNorm+modulate, attention (+ inverse rotary)
// and the final first-order term.  Everything is a first version aimed at
// correctness (fp32 CUDA-core math on 16-bit tensors); the classifier is small
// next to the denoiser.
#include <cstdlib>

#include "ddg_common.cuh"

namespace ddg {
// attention_bwd.cu: tensor-core path for seq_len <= 128
bool attention_bwd_tc_supported(int batch, int seq_len, int causal);   // attention_bwd_tc.cu
int launch_attention_bwd_tc(const void* qkv, const void* g_o, const void* rope_cos, const void* rope_sin, int batch,
                            int seq_len, int n_heads, int dtype, void* g_qkv, cudaStream_t s);
int launch_attention_bwd_mma(const void* qkv, const void* g_o, const void* rope_cos, const void* rope_sin, int batch,
                             int seq_len, int n_heads, int causal, int dtype, void* g_qkv, cudaStream_t s);
namespace {

// ---------------------------------------------------------------- head backward
// log p = log_softmax(Linear(mean_l x)) (models/dit.py:538-555, classifier.py:316-327).
// d log p[y] / d x[l, :] = W^T (onehot(y) - softmax) / L.  One CTA per sequence.
__global__ void __launch_bounds__(256)
classifier_head_bwd_kernel(const float* __restrict__ x, int seq_len, int hidden, const float* __restrict__ w,
                           const float* __restrict__ b, int n_classes, int target, int round_bf16,
                           float* __restrict__ log_probs, float* __restrict__ g_x) {
  extern __shared__ float sm[];
  float* mean = sm;               // hidden
  float* logit = sm + hidden;     // n_classes, then g_logit
  const int64_t n = blockIdx.x;
  const float* xs = x + n * (int64_t)seq_len * hidden;
  for (int col = threadIdx.x; col < hidden; col += blockDim.x) {
    float acc = 0.f;
    for (int l = 0; l < seq_len; ++l) acc += xs[(int64_t)l * hidden + col];
    mean[col] = acc / (float)seq_len;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  for (int c = warp; c < n_classes; c += nwarp) {
    float acc = 0.f;
    for (int k = lane; k < hidden; k += 32) {
      float a = mean[k], ww = w[(int64_t)c * hidden + k];
      if (round_bf16) { a = bf16_round(a); ww = bf16_round(ww); }
      acc += a * ww;
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      float bb = b[c];
      if (round_bf16) bb = bf16_round(bb);
      float v = acc + bb;
      if (round_bf16) v = bf16_round(v);
      logit[c] = v;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = -INFINITY;
    for (int c = 0; c < n_classes; ++c) m = fmaxf(m, logit[c]);
    float s = 0.f;
    for (int c = 0; c < n_classes; ++c) s += expf(logit[c] - m);
    const float ls = logf(s);
    for (int c = 0; c < n_classes; ++c) {
      const float lp = (logit[c] - m) - ls;
      log_probs[n * n_classes + c] = lp;   // fp32 log_softmax of the bf16 logits (classifier.py:325-327)
      logit[c] = (c == target ? 1.f : 0.f) - expf(lp);       // d log p[target] / d logit[c]
    }
  }
  __syncthreads();
  const float inv_l = 1.f / (float)seq_len;
  float* gs = g_x + n * (int64_t)seq_len * hidden;
  for (int col = threadIdx.x; col < hidden; col += blockDim.x) {
    float acc = 0.f;
    for (int c = 0; c < n_classes; ++c) {
      float ww = w[(int64_t)c * hidden + col];
      if (round_bf16) ww = bf16_round(ww);
      acc += logit[c] * ww;
    }
    acc *= inv_l;
    for (int l = 0; l < seq_len; ++l) gs[(int64_t)l * hidden + col] = acc;
  }
}

// ---------------------------------------------------------------- gate (x) g -> 16-bit
template <typename T>
__global__ void gate_mul_kernel(const float* __restrict__ g, const T* __restrict__ gate, int64_t gate_stride,
                                int64_t rows_per_mod, int64_t rows, int hidden, T* __restrict__ out) {
  const int64_t n = rows * hidden;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / hidden;
    const int col = (int)(i % hidden);
    float v = g[i];
    if (gate) v *= Cvt<T>::to_f32(gate[(row / rows_per_mod) * gate_stride + col]);
    out[i] = Cvt<T>::from_f32(v);
  }
}

// 16-bit pairs <-> fp32
template <typename T> __device__ __forceinline__ float2 unpack2(uint32_t w);
template <> __device__ __forceinline__ float2 unpack2<__half>(uint32_t w) {
  return __half22float2(*reinterpret_cast<__half2*>(&w));
}
template <> __device__ __forceinline__ float2 unpack2<__nv_bfloat16>(uint32_t w) {
  return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}
template <typename T> __device__ __forceinline__ uint32_t pack2(float a, float b);
template <> __device__ __forceinline__ uint32_t pack2<__half>(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
template <> __device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// 8 columns per thread (hidden % 8 == 0, 16-byte aligned rows): two 128-bit loads of g, one
// of the gate row, one 128-bit store; one row division per 8 elements.  Same per-element
// arithmetic as the scalar kernel above.
template <typename T>
__global__ void __launch_bounds__(256)
gate_mul_vec_kernel(const float* __restrict__ g, const T* __restrict__ gate, int64_t gate_stride,
                    int64_t rows_per_mod, int64_t rows, int hidden, T* __restrict__ out) {
  const int vec_per_row = hidden >> 3;
  const int64_t n8 = rows * vec_per_row;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / vec_per_row;
    const int col = (int)(i - row * vec_per_row) << 3;
    const float4 a = __ldcs(reinterpret_cast<const float4*>(g + row * hidden + col));
    const float4 b = __ldcs(reinterpret_cast<const float4*>(g + row * hidden + col) + 1);
    float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    if (gate) {
      const uint4 gw = __ldg(reinterpret_cast<const uint4*>(gate + (row / rows_per_mod) * gate_stride + col));
      const uint32_t w[4] = {gw.x, gw.y, gw.z, gw.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = unpack2<T>(w[k]);
        v[2 * k] *= f.x;
        v[2 * k + 1] *= f.y;
      }
    }
    uint4 o;
    o.x = pack2<T>(v[0], v[1]); o.y = pack2<T>(v[2], v[3]); o.z = pack2<T>(v[4], v[5]); o.w = pack2<T>(v[6], v[7]);
    *reinterpret_cast<uint4*>(out + row * hidden + col) = o;
  }
}

// ---------------------------------------------------------------- GELU'(pre) * g
// gelu_tanh_grad(): ddg_common.cuh (shared with the GEMM's GELU_BWD epilogue)
template <typename T>
__global__ void gelu_bwd_kernel(const T* __restrict__ g_h, const T* __restrict__ pre, T* __restrict__ out, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = Cvt<T>::from_f32(Cvt<T>::to_f32(g_h[i]) * gelu_tanh_grad(Cvt<T>::to_f32(pre[i])));
}
// 8 elements per thread, 128-bit loads / stores (n % 8 == 0, 16-byte aligned pointers)
template <typename T>
__global__ void __launch_bounds__(256)
gelu_bwd_vec_kernel(const T* __restrict__ g_h, const T* __restrict__ pre, T* __restrict__ out, int64_t n8) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    const uint4 gw = __ldcs(reinterpret_cast<const uint4*>(g_h) + i);
    const uint4 pw = __ldcs(reinterpret_cast<const uint4*>(pre) + i);
    const uint32_t a[4] = {gw.x, gw.y, gw.z, gw.w}, b[4] = {pw.x, pw.y, pw.z, pw.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 gg = unpack2<T>(a[k]), pp = unpack2<T>(b[k]);
      o[k] = pack2<T>(__fmul_rn(gg.x, gelu_tanh_grad(pp.x)), __fmul_rn(gg.y, gelu_tanh_grad(pp.y)));
    }
    reinterpret_cast<uint4*>(out)[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// ---------------------------------------------------------------- LayerNorm + modulate backward
// a = LN(x) * w * ops + shift  (models/dit.py:131-139, 47-50);  g_resid += d a / d x ^T g_a.
// EXACT: hidden == 128 * MAXV (the loops unroll fully and no register is held for unused slots)
template <typename T, int MAXV, bool EXACT = false>
__global__ void __launch_bounds__(256)
ln_modulate_bwd_kernel(const float* __restrict__ x, const T* __restrict__ g_a, int64_t rows, int hidden,
                       const float* __restrict__ ln_w, const T* __restrict__ ops, int64_t mod_stride,
                       int64_t rows_per_mod, float* __restrict__ g_resid) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int nv = EXACT ? MAXV : hidden >> 7;
  float4 v[MAXV], gh[MAXV];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i)
    if (i < nv) {
      v[i] = *reinterpret_cast<const float4*>(x + row * (int64_t)hidden + (i * 32 + lane) * 4);
      sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  const float mean = warp_sum(sum) / (float)hidden;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i)
    if (i < nv) {
      v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
      sq += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
  const float rstd = rsqrtf(warp_sum(sq) / (float)hidden + 1e-5f);
  const T* sc = ops ? ops + (rows_per_mod > 0 ? row / rows_per_mod : 0) * mod_stride : nullptr;
  float s1 = 0.f, s2 = 0.f;            // sum(ghat), sum(ghat * nhat)
#pragma unroll
  for (int i = 0; i < MAXV; ++i)
    if (i < nv) {
      const int col = (i * 32 + lane) * 4;
      const float4 w4 = *reinterpret_cast<const float4*>(ln_w + col);
      const uint2 gw = *reinterpret_cast<const uint2*>(g_a + row * (int64_t)hidden + col);   // 4 x 16-bit
      const float2 g01 = unpack2<T>(gw.x), g23 = unpack2<T>(gw.y);
      float gv[4] = {g01.x * w4.x, g01.y * w4.y, g23.x * w4.z, g23.y * w4.w};
      if (sc) {
        const uint2 sw = *reinterpret_cast<const uint2*>(sc + col);
        const float2 s01 = unpack2<T>(sw.x), s23 = unpack2<T>(sw.y);
        gv[0] *= s01.x; gv[1] *= s01.y; gv[2] *= s23.x; gv[3] *= s23.y;
      }
      v[i].x *= rstd; v[i].y *= rstd; v[i].z *= rstd; v[i].w *= rstd;      // nhat
      gh[i] = make_float4(gv[0], gv[1], gv[2], gv[3]);
      s1 += (gv[0] + gv[1]) + (gv[2] + gv[3]);
      s2 += (gv[0] * v[i].x + gv[1] * v[i].y) + (gv[2] * v[i].z + gv[3] * v[i].w);
    }
  s1 = warp_sum(s1) / (float)hidden;
  s2 = warp_sum(s2) / (float)hidden;
#pragma unroll
  for (int i = 0; i < MAXV; ++i)
    if (i < nv) {
      float4* gr = reinterpret_cast<float4*>(g_resid + row * (int64_t)hidden + (i * 32 + lane) * 4);
      float4 o = *gr;
      o.x += rstd * (gh[i].x - s1 - v[i].x * s2);
      o.y += rstd * (gh[i].y - s1 - v[i].y * s2);
      o.z += rstd * (gh[i].z - s1 - v[i].z * s2);
      o.w += rstd * (gh[i].w - s1 - v[i].w * s2);
      *gr = o;
    }
}

// ---------------------------------------------------------------- attention backward (seq_len > 128)
// CUDA-core fallback for long sequences; seq_len <= 128 runs attention_bwd.cu's tensor-core kernel.
// One CTA per (sequence, head); q, k (post-rotary), v and dO of the head in shared
// memory (row stride 66 elements: conflict-free column walks).  Pass A: one warp per
// query row -> softmax stats, D_i = dO_i . o_i, dq_i.  Pass B: one warp per key row ->
// dk_j, dv_j.  dq / dk go back through the rotary (transpose of the rotation) so that
// the result is the gradient of the QKV GEMM output (models/dit.py:284-307).
constexpr int kRowPad = 66;
template <typename T>
__global__ void __launch_bounds__(256)
attention_bwd_kernel(const T* __restrict__ qkv, const T* __restrict__ g_o, const T* __restrict__ rope_cos,
                     const T* __restrict__ rope_sin, int L, int H, int causal, T* __restrict__ g_qkv) {
  extern __shared__ __align__(16) uint8_t smraw[];
  T* sq = reinterpret_cast<T*>(smraw);
  T* sk = sq + (size_t)L * kRowPad;
  T* sv = sk + (size_t)L * kRowPad;
  T* sd = sv + (size_t)L * kRowPad;
  float* st_m = reinterpret_cast<float*>(sd + (size_t)L * kRowPad);
  float* st_l = st_m + L;
  float* st_D = st_l + L;
  float* scratch = st_D + L;                     // 8 warps x 2 x L floats
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t row_stride = (int64_t)3 * H * 64;
  const T* base = qkv + (int64_t)b * L * row_stride + (int64_t)h * 64;
  const T* gbase = g_o + (int64_t)b * L * ((int64_t)H * 64) + (int64_t)h * 64;
  for (int i = tid; i < L * 64; i += blockDim.x) {
    const int r = i >> 6, d = i & 63;
    sq[r * kRowPad + d] = base[(int64_t)r * row_stride + d];
    sk[r * kRowPad + d] = base[(int64_t)r * row_stride + (int64_t)H * 64 + d];
    sv[r * kRowPad + d] = base[(int64_t)r * row_stride + (int64_t)2 * H * 64 + d];
    sd[r * kRowPad + d] = gbase[(int64_t)r * H * 64 + d];
  }
  __syncthreads();
  const float scale = 0.125f;                    // 1 / sqrt(64)
  const int nt = (L + 31) >> 5;                  // keys (queries) per lane
  float* my = scratch + (size_t)warp * 2 * L;
  auto dot64 = [](const T* a, const T* bb) {
    float acc = 0.f;
#pragma unroll 16
    for (int d = 0; d < 64; ++d) acc += Cvt<T>::to_f32(a[d]) * Cvt<T>::to_f32(bb[d]);
    return acc;
  };
  // inverse rotary of the (2 lane, 2 lane + 1) dims of a row at position pos
  auto unrotate = [&](float& g0, float& g1, int pos) {
    const float o0 = __shfl_xor_sync(0xffffffffu, g0, 16), o1 = __shfl_xor_sync(0xffffffffu, g1, 16);
    const int idx = 2 * (lane & 15);
    const float c0 = Cvt<T>::to_f32(rope_cos[pos * 32 + idx]), c1 = Cvt<T>::to_f32(rope_cos[pos * 32 + idx + 1]);
    const float s0 = Cvt<T>::to_f32(rope_sin[pos * 32 + idx]), s1 = Cvt<T>::to_f32(rope_sin[pos * 32 + idx + 1]);
    if (lane < 16) { g0 = g0 * c0 + o0 * s0; g1 = g1 * c1 + o1 * s1; }      // g_x1 = g_o1 cos + g_o2 sin
    else           { g0 = g0 * c0 - o0 * s0; g1 = g1 * c1 - o1 * s1; }      // g_x2 = g_o2 cos - g_o1 sin
  };
  // ---- pass A: per query row
  for (int i = warp; i < L; i += 8) {
    float s[8], dp[8];
    float m = -INFINITY;
    for (int t = 0; t < nt; ++t) {
      const int j = lane + 32 * t;
      const bool ok = j < L && !(causal && j > i);
      s[t] = ok ? scale * dot64(sq + i * kRowPad, sk + j * kRowPad) : -INFINITY;
      dp[t] = ok ? dot64(sd + i * kRowPad, sv + j * kRowPad) : 0.f;
      m = fmaxf(m, s[t]);
    }
    m = warp_max(m);
    float l = 0.f;
    for (int t = 0; t < nt; ++t) { s[t] = expf(s[t] - m); l += s[t]; }
    l = warp_sum(l);
    float D = 0.f;
    for (int t = 0; t < nt; ++t) { s[t] /= l; D += s[t] * dp[t]; }
    D = warp_sum(D);
    if (lane == 0) { st_m[i] = m; st_l[i] = l; st_D[i] = D; }
    __syncwarp();
    for (int t = 0; t < nt; ++t) {
      const int j = lane + 32 * t;
      if (j < L) my[j] = s[t] * (dp[t] - D) * scale;        // dS_ij
    }
    __syncwarp();
    float g0 = 0.f, g1 = 0.f;
    for (int j = 0; j < L; ++j) {
      const float ds = my[j];
      g0 += ds * Cvt<T>::to_f32(sk[j * kRowPad + 2 * lane]);
      g1 += ds * Cvt<T>::to_f32(sk[j * kRowPad + 2 * lane + 1]);
    }
    unrotate(g0, g1, i);
    T* o = g_qkv + ((int64_t)b * L + i) * row_stride + (int64_t)h * 64 + 2 * lane;
    o[0] = Cvt<T>::from_f32(g0);
    o[1] = Cvt<T>::from_f32(g1);
    __syncwarp();
  }
  __syncthreads();
  // ---- pass B: per key row
  for (int j = warp; j < L; j += 8) {
    for (int t = 0; t < nt; ++t) {
      const int i = lane + 32 * t;
      if (i < L) {
        const bool ok = !(causal && j > i);
        float p = 0.f, ds = 0.f;
        if (ok) {
          p = expf(scale * dot64(sq + i * kRowPad, sk + j * kRowPad) - st_m[i]) / st_l[i];
          ds = p * (dot64(sd + i * kRowPad, sv + j * kRowPad) - st_D[i]) * scale;
        }
        my[i] = p;
        my[L + i] = ds;
      }
    }
    __syncwarp();
    float k0 = 0.f, k1 = 0.f, v0 = 0.f, v1 = 0.f;
    for (int i = 0; i < L; ++i) {
      const float p = my[i], ds = my[L + i];
      k0 += ds * Cvt<T>::to_f32(sq[i * kRowPad + 2 * lane]);
      k1 += ds * Cvt<T>::to_f32(sq[i * kRowPad + 2 * lane + 1]);
      v0 += p * Cvt<T>::to_f32(sd[i * kRowPad + 2 * lane]);
      v1 += p * Cvt<T>::to_f32(sd[i * kRowPad + 2 * lane + 1]);
    }
    unrotate(k0, k1, j);
    T* o = g_qkv + ((int64_t)b * L + j) * row_stride + (int64_t)h * 64 + 2 * lane;
    o[(int64_t)H * 64] = Cvt<T>::from_f32(k0);
    o[(int64_t)H * 64 + 1] = Cvt<T>::from_f32(k1);
    o[(int64_t)2 * H * 64] = Cvt<T>::from_f32(v0);
    o[(int64_t)2 * H * 64 + 1] = Cvt<T>::from_f32(v1);
    __syncwarp();
  }
}

// ---------------------------------------------------------------- fp32 -> (hi, lo) 16-bit split
template <typename T>
__global__ void split16_kernel(const float* __restrict__ src, T* __restrict__ hi, T* __restrict__ lo, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = src[i];
    const T h = Cvt<T>::from_f32(v);
    hi[i] = h;
    lo[i] = Cvt<T>::from_f32(v - Cvt<T>::to_f32(h));
  }
}

// ---------------------------------------------------------------- first-order term
// term[row, v] = grad[row, v] - grad[row, xt[row]] + log p(y | xt)[row / L]   (diffusion.py:1541-1548)
__global__ void __launch_bounds__(256)
cbg_first_order_kernel(const float* grad, int64_t ld_grad, const int64_t* __restrict__ xt,
                       const float* __restrict__ log_probs, int n_classes, int target, int seq_len,
                       int vocab, int64_t rows, float* term, int64_t ld_term) {
  const bool vec = ((ld_grad | ld_term) & 3) == 0 && (reinterpret_cast<uintptr_t>(grad) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(term) & 15) == 0;
  for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
    const float* g = grad + row * ld_grad;
    float* t = term + row * ld_term;
    const float base = log_probs[(row / seq_len) * n_classes + target] - g[xt[row]];
    __syncthreads();                               // term may alias grad: everyone has read g[xt] by now
    if (vec) {                                     // 128-bit loads / stores; the tail (vocab % 4) scalar
      const int n4 = vocab >> 2;
      for (int v = threadIdx.x; v < n4; v += blockDim.x) {
        float4 x = __ldcs(reinterpret_cast<const float4*>(g) + v);
        x.x += base; x.y += base; x.z += base; x.w += base;
        reinterpret_cast<float4*>(t)[v] = x;
      }
      for (int v = (n4 << 2) + threadIdx.x; v < vocab; v += blockDim.x) t[v] = g[v] + base;
    } else {
      for (int v = threadIdx.x; v < vocab; v += blockDim.x) t[v] = g[v] + base;
    }
    __syncthreads();                               // the next row's g[xt] is read after this row's stores
  }
}

int64_t grid_for(int64_t n) {
  int64_t g = (n + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  return g > cap ? cap : (g < 1 ? 1 : g);
}

}  // namespace
}  // namespace ddg

using namespace ddg;

extern "C" int ddg_classifier_head_backward(const float* x, int32_t n, int32_t seq_len, int32_t hidden,
                                            const float* w, const float* b, int32_t n_classes, int32_t target,
                                            int32_t round_bf16, float* log_probs, float* g_x,
                                            ddg_stream_t stream) {
  int rc = ddg_device_check();
  if (rc) return rc;
  DDG_REQUIRE(x && w && b && log_probs && g_x && n >= 0 && seq_len > 0 && hidden > 0 && n_classes > 0 &&
              target >= 0 && target < n_classes, DDG_ERR_INVALID, "classifier_head_backward: bad args");
  if (n == 0) return DDG_OK;
  const size_t smem = sizeof(float) * (size_t)(hidden + n_classes);
  classifier_head_bwd_kernel<<<n, 256, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
    x, seq_len, hidden, w, b, n_classes, target, round_bf16, log_probs, g_x);
  return check_cuda(cudaGetLastError(), "classifier_head_backward launch");
}

extern "C" int ddg_gate_mul(const float* g, const void* gate, int64_t gate_stride, int64_t rows_per_mod,
                            int64_t rows, int32_t hidden, int32_t dtype, void* out, ddg_stream_t stream) {
  int rc = ddg_device_check();
  if (rc) return rc;
  DDG_REQUIRE(g && out && rows >= 0 && hidden > 0 && (!gate || rows_per_mod > 0), DDG_ERR_INVALID, "gate_mul: bad args");
  if (rows == 0) return DDG_OK;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const bool vec = hidden % 8 == 0 && gate_stride % 8 == 0 && reinterpret_cast<uintptr_t>(g) % 16 == 0 &&
                   reinterpret_cast<uintptr_t>(gate) % 16 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0;
  if (vec && (dtype == DDG_F16 || dtype == DDG_BF16)) {
    const unsigned vgrid = (unsigned)grid_for(rows * (hidden / 8));
    if (dtype == DDG_F16)
      gate_mul_vec_kernel<__half><<<vgrid, 256, 0, s>>>(g, reinterpret_cast<const __half*>(gate), gate_stride,
                                                        rows_per_mod, rows, hidden, reinterpret_cast<__half*>(out));
    else
      gate_mul_vec_kernel<__nv_bfloat16><<<vgrid, 256, 0, s>>>(g, reinterpret_cast<const __nv_bfloat16*>(gate),
                                                               gate_stride, rows_per_mod, rows, hidden,
                                                               reinterpret_cast<__nv_bfloat16*>(out));
    return check_cuda(cudaGetLastError(), "gate_mul launch");
  }
  const unsigned grid = (unsigned)grid_for(rows * hidden);
  if (dtype == DDG_F16)
    gate_mul_kernel<__half><<<grid, 256, 0, s>>>(g, reinterpret_cast<const __half*>(gate), gate_stride, rows_per_mod,
                                                 rows, hidden, reinterpret_cast<__half*>(out));
  else if (dtype == DDG_BF16)
    gate_mul_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(g, reinterpret_cast<const __nv_bfloat16*>(gate), gate_stride,
                                                        rows_per_mod, rows, hidden,
                                                        reinterpret_cast<__nv_bfloat16*>(out));
  else { set_error("gate_mul: dtype must be f16 or bf16"); return DDG_ERR_INVALID; }
  return check_cuda(cudaGetLastError(), "gate_mul launch");
}

extern "C" int ddg_gelu_backward(const void* g_h, const void* pre, void* out, int64_t n, int32_t dtype,
                                 ddg_stream_t stream) {
  int rc = ddg_device_check();
  if (rc) return rc;
  DDG_REQUIRE(g_h && pre && out && n >= 0, DDG_ERR_INVALID, "gelu_backward: bad args");
  if (n == 0) return DDG_OK;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const bool vec = n % 8 == 0 && reinterpret_cast<uintptr_t>(g_h) % 16 == 0 &&
                   reinterpret_cast<uintptr_t>(pre) % 16 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0;
  if (vec && (dtype == DDG_F16 || dtype == DDG_BF16)) {
    const unsigned vgrid = (unsigned)grid_for(n / 8);
    if (dtype == DDG_F16)
      gelu_bwd_vec_kernel<__half><<<vgrid, 256, 0, s>>>(reinterpret_cast<const __half*>(g_h),
                                                        reinterpret_cast<const __half*>(pre),
                                                        reinterpret_cast<__half*>(out), n / 8);
    else
      gelu_bwd_vec_kernel<__nv_bfloat16><<<vgrid, 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(g_h),
                                                               reinterpret_cast<const __nv_bfloat16*>(pre),
                                                               reinterpret_cast<__nv_bfloat16*>(out), n / 8);
    return check_cuda(cudaGetLastError(), "gelu_backward launch");
  }
  const unsigned grid = (unsigned)grid_for(n);
  if (dtype == DDG_F16)
    gelu_bwd_kernel<__half><<<grid, 256, 0, s>>>(reinterpret_cast<const __half*>(g_h),
                                                 reinterpret_cast<const __half*>(pre), reinterpret_cast<__half*>(out), n);
  else if (dtype == DDG_BF16)
    gelu_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(g_h),
                                                        reinterpret_cast<const __nv_bfloat16*>(pre),
                                                        reinterpret_cast<__nv_bfloat16*>(out), n);
  else { set_error("gelu_backward: dtype must be f16 or bf16"); return DDG_ERR_INVALID; }
  return check_cuda(cudaGetLastError(), "gelu_backward launch");
}

extern "C" int ddg_ln_modulate_backward(const float* x, const void* g_a, int64_t rows, int32_t hidden,
                                        const float* ln_weight, const void* one_plus_scale, int64_t mod_stride,
                                        int64_t rows_per_mod, int32_t dtype, float* g_resid, ddg_stream_t stream) {
  int rc = ddg_device_check();
  if (rc) return rc;
  DDG_REQUIRE(x && g_a && ln_weight && g_resid && rows >= 0, DDG_ERR_INVALID, "ln_modulate_backward: bad args");
  DDG_REQUIRE(hidden % 128 == 0 && hidden >= 128 && hidden <= 2048, DDG_ERR_UNSUPPORTED,
              "ln_modulate_backward: hidden=%d must be a multiple of 128 in [128, 2048]", hidden);
  DDG_REQUIRE(reinterpret_cast<uintptr_t>(x) % 16 == 0 && reinterpret_cast<uintptr_t>(g_resid) % 16 == 0 &&
              reinterpret_cast<uintptr_t>(g_a) % 8 == 0 && reinterpret_cast<uintptr_t>(one_plus_scale) % 8 == 0 &&
              mod_stride % 4 == 0, DDG_ERR_INVALID, "ln_modulate_backward: pointer alignment / mod_stride");
  if (rows == 0) return DDG_OK;
  const unsigned grid = (unsigned)((rows + 7) / 8);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
#define DDG_LNB_LAUNCH(T, MAXV)                                                                              \
  ln_modulate_bwd_kernel<T, MAXV><<<grid, 256, 0, s>>>(x, reinterpret_cast<const T*>(g_a), rows, hidden,     \
                                                       ln_weight, reinterpret_cast<const T*>(one_plus_scale), \
                                                       mod_stride, rows_per_mod, g_resid)
#define DDG_LNB_EXACT(T, NV)                                                                                  \
  ln_modulate_bwd_kernel<T, NV, true><<<grid, 256, 0, s>>>(x, reinterpret_cast<const T*>(g_a), rows, hidden,   \
                                                           ln_weight, reinterpret_cast<const T*>(one_plus_scale), \
                                                           mod_stride, rows_per_mod, g_resid)
  if (dtype == DDG_F16) {
    if (hidden == 512) DDG_LNB_EXACT(__half, 4);
    else if (hidden == 768) DDG_LNB_EXACT(__half, 6);
    else if (hidden <= 1024) DDG_LNB_LAUNCH(__half, 8); else DDG_LNB_LAUNCH(__half, 16);
  } else if (dtype == DDG_BF16) {
    if (hidden == 512) DDG_LNB_EXACT(__nv_bfloat16, 4);
    else if (hidden == 768) DDG_LNB_EXACT(__nv_bfloat16, 6);
    else if (hidden <= 1024) DDG_LNB_LAUNCH(__nv_bfloat16, 8); else DDG_LNB_LAUNCH(__nv_bfloat16, 16);
  } else { set_error("ln_modulate_backward: dtype must be f16 or bf16"); return DDG_ERR_INVALID; }
#undef DDG_LNB_LAUNCH
#undef DDG_LNB_EXACT
  return check_cuda(cudaGetLastError(), "ln_modulate_backward launch");
}

extern "C" int ddg_attention_backward(const void* qkv, const void* g_out, const void* rope_cos, const void* rope_sin,
                                      int32_t batch, int32_t seq_len, int32_t n_heads, int32_t head_dim,
                                      int32_t causal, int32_t dtype, void* g_qkv, ddg_stream_t stream) {
  int rc = ddg_device_check();
  if (rc) return rc;
  DDG_REQUIRE(qkv && g_out && rope_cos && rope_sin && g_qkv && batch >= 0 && n_heads > 0, DDG_ERR_INVALID,
              "attention_backward: bad args");
  DDG_REQUIRE(head_dim == 64 && seq_len >= 1 && seq_len <= 256, DDG_ERR_UNSUPPORTED,
              "attention_backward: head_dim must be 64 and seq_len <= 256");
  if (batch == 0) return DDG_OK;
  DDG_REQUIRE(dtype == DDG_F16 || dtype == DDG_BF16, DDG_ERR_INVALID, "attention_backward: dtype must be f16 or bf16");
  // tcgen05 / tensor-memory kernel for the non-causal 32 / 64 / 128-token shapes (DDG_ATT_BWD=mma
  // keeps the mma.sync kernel: A/B measurements and the parity test of the two kernels)
  static const bool force_mma = [] {
    const char* e = getenv("DDG_ATT_BWD");
    return e != nullptr && e[0] == 'm';
  }();
  if (!force_mma && attention_bwd_tc_supported(batch, seq_len, causal) &&
      reinterpret_cast<uintptr_t>(qkv) % 16 == 0 && reinterpret_cast<uintptr_t>(g_out) % 16 == 0 &&
      reinterpret_cast<uintptr_t>(g_qkv) % 16 == 0)
    return launch_attention_bwd_tc(qkv, g_out, rope_cos, rope_sin, batch, seq_len, n_heads, dtype, g_qkv,
                                   reinterpret_cast<cudaStream_t>(stream));
  if (seq_len <= 128)
    return launch_attention_bwd_mma(qkv, g_out, rope_cos, rope_sin, batch, seq_len, n_heads, causal, dtype, g_qkv,
                                    reinterpret_cast<cudaStream_t>(stream));
  const size_t smem = (size_t)4 * seq_len * kRowPad * 2 + (size_t)3 * seq_len * 4 + (size_t)8 * 2 * seq_len * 4;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  static bool attr_done[64][2] = {{false}};
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned grid = (unsigned)(batch * n_heads);
  if (dtype == DDG_F16) {
    if (dev >= 0 && dev < 64 && !attr_done[dev][0]) {
      DDG_CUDA(cudaFuncSetAttribute(attention_bwd_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr_done[dev][0] = true;
    }
    attention_bwd_kernel<__half><<<grid, 256, smem, s>>>(
      reinterpret_cast<const __half*>(qkv), reinterpret_cast<const __half*>(g_out),
      reinterpret_cast<const __half*>(rope_cos), reinterpret_cast<const __half*>(rope_sin), seq_len, n_heads, causal,
      reinterpret_cast<__half*>(g_qkv));
  } else if (dtype == DDG_BF16) {
    if (dev >= 0 && dev < 64 && !attr_done[dev][1]) {
      DDG_CUDA(cudaFuncSetAttribute(attention_bwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    200 * 1024));
      attr_done[dev][1] = true;
    }
    attention_bwd_kernel<__nv_bfloat16><<<grid, 256, smem, s>>>(
      reinterpret_cast<const __nv_bfloat16*>(qkv), reinterpret_cast<const __nv_bfloat16*>(g_out),
      reinterpret_cast<const __nv_bfloat16*>(rope_cos), reinterpret_cast<const __nv_bfloat16*>(rope_sin), seq_len,
      n_heads, causal, reinterpret_cast<__nv_bfloat16*>(g_qkv));
  } else { set_error("attention_backward: dtype must be f16 or bf16"); return DDG_ERR_INVALID; }
  return check_cuda(cudaGetLastError(), "attention_backward launch");
}

extern "C" int ddg_split16(const float* src, void* hi, void* lo, int64_t n, int32_t dtype, ddg_stream_t stream) {
  int rc = ddg_device_check();
  if (rc) return rc;
  DDG_REQUIRE(src && hi && lo && n >= 0, DDG_ERR_INVALID, "split16: bad args");
  if (n == 0) return DDG_OK;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const unsigned grid = (unsigned)grid_for(n);
  if (dtype == DDG_F16)
    split16_kernel<__half><<<grid, 256, 0, s>>>(src, reinterpret_cast<__half*>(hi), reinterpret_cast<__half*>(lo), n);
  else if (dtype == DDG_BF16)
    split16_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(src, reinterpret_cast<__nv_bfloat16*>(hi),
                                                       reinterpret_cast<__nv_bfloat16*>(lo), n);
  else { set_error("split16: dtype must be f16 or bf16"); return DDG_ERR_INVALID; }
  return check_cuda(cudaGetLastError(), "split16 launch");
}

extern "C" int ddg_cbg_first_order(const float* grad, int64_t ld_grad, const int64_t* xt, const float* log_probs,
                                   int32_t n_classes, int32_t target, int32_t seq_len, int32_t vocab, int64_t rows,
                                   float* term, int64_t ld_term, ddg_stream_t stream) {
  int rc = ddg_device_check();
  if (rc) return rc;
  DDG_REQUIRE(grad && xt && log_probs && term && rows >= 0 && vocab > 0 && seq_len > 0 && ld_grad >= vocab &&
              ld_term >= vocab && target >= 0 && target < n_classes, DDG_ERR_INVALID, "cbg_first_order: bad args");
  if (rows == 0) return DDG_OK;
  const int64_t cap = (int64_t)ddg::sm_count() * 8;
  cbg_first_order_kernel<<<(unsigned)(rows < cap ? rows : cap), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
    grad, ld_grad, xt, log_probs, n_classes, target, seq_len, vocab, rows, term, ld_term);
  return check_cuda(cudaGetLastError(), "cbg_first_order launch");
}
