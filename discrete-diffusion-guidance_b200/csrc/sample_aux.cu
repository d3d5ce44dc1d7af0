// Two rarely used sampling flags of the reference, each as its own kernel so
// that the production kernels stay lean:
//
//  * sampling.use_float64 (reference diffusion.py:1393-1394, 1110-1111,
//    1134-1135): the log-probs keep the forward's precision (fp32, or bf16 with
//    round_bf16) and are then promoted; exp, posterior, the race and the arg-max
//    run in double.  Honoured by the reference only in `_ddpm_denoise` and the
//    AR sampler, so: absorbing / uniform / AR without guidance.
//
//  * eval.low_confidence_sampling (reference diffusion.py:1738-1746): keep the
//    entries whose cumulative mass in ascending-probability order stays
//    <= eval.low_confidence_threshold, zero the rest, then race / arg-max.
//    torch's sort + cumsum + scatter over (B, L, V) is replaced by a bisection
//    on the probability value (the row sits in shared memory) plus an
//    index-ordered prefix count inside the tie group at the threshold (stable
//    sort order, which is what torch.sort gives for equal keys).
#include "ddg_common.cuh"

namespace ddg {
namespace {

constexpr float kNegInf = -1000000.0f;  // diffusion.py:141
constexpr int kT = 256;

template <typename V> __device__ __forceinline__ V block_reduce_sum(V v, V* sm) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (l == 0) sm[w] = v;
  __syncthreads();
  v = l < nw ? sm[l] : V(0);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float block_reduce_max(float v, float* sm) {
  v = warp_max(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (l == 0) sm[w] = v;
  __syncthreads();
  v = l < nw ? sm[l] : -INFINITY;
  return warp_max(v);
}
// first index wins ties (torch.argmax)
__device__ __forceinline__ void block_argmax(double& s, int& i, double* sm_s, int* sm_i) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double os = __shfl_xor_sync(0xffffffffu, s, o);
    const int oi = __shfl_xor_sync(0xffffffffu, i, o);
    if (os > s || (os == s && oi < i)) { s = os; i = oi; }
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (l == 0) { sm_s[w] = s; sm_i[w] = i; }
  __syncthreads();
  if (l < nw) { s = sm_s[l]; i = sm_i[l]; } else { s = -INFINITY; i = 0x7fffffff; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double os = __shfl_xor_sync(0xffffffffu, s, o);
    const int oi = __shfl_xor_sync(0xffffffffu, i, o);
    if (os > s || (os == s && oi < i)) { s = os; i = oi; }
  }
}

__device__ __forceinline__ float philox_u(int j, int64_t row, int step, uint2 key) {
  const uint4 w = philox4x32_10(make_uint4((uint32_t)(j >> 2), (uint32_t)row, (uint32_t)((uint64_t)row >> 32),
                                           (uint32_t)step), key);
  const int s = j & 3;
  return u01_from_bits(s == 0 ? w.x : s == 1 ? w.y : s == 2 ? w.z : w.w);
}

// ------------------------------------------------------------------ float64
template <typename T>
__global__ void __launch_bounds__(kT)
sample_f64_kernel(const ddg_sample_step_args a) {
  __shared__ double sm_d[kT / 32];
  __shared__ float sm_f[kT / 32];
  __shared__ int sm_i[kT / 32];
  const int tid = threadIdx.x;
  const int V = a.vocab;
  const int step = a.step_ptr ? *a.step_ptr : a.step;
  const uint64_t seed64 = a.seed_ptr ? *a.seed_ptr : a.seed;
  const uint2 key = make_uint2((uint32_t)seed64, (uint32_t)(seed64 >> 32));
  const bool absorbing = a.diffusion == DDG_ABSORBING, ar = a.diffusion == DDG_AR;

  for (int64_t row = blockIdx.x; row < a.rows; row += gridDim.x) {
    const int z = (int)a.xt[row];
    if (absorbing && z != a.mask_index) {            // carry-over (diffusion.py:1414-1418)
      if (tid == 0) a.xs[row] = z;
      if (a.q_out) {
        float* q = a.q_out + row * a.ld_q;
        for (int j = tid; j < V; j += kT) q[j] = (j == z) ? 1.f : 0.f;
      }
      continue;
    }
    const T* pl = reinterpret_cast<const T*>(a.logits_a) + row * a.ld_a;
    const float* pn = a.noise ? a.noise + row * a.ld_noise : nullptr;
    auto logit = [&](int j) {
      float v = Cvt<T>::to_f32(pl[j]);
      if (a.mask_logit && j == a.mask_index) v += kNegInf;
      return v;
    };
    // ---- log-softmax at the forward's precision (fp32; bf16 result with round_bf16)
    float m = -INFINITY;
    for (int j = tid; j < V; j += kT) m = fmaxf(m, logit(j));
    const float M = block_reduce_max(m, sm_f);
    float s = 0.f;
    for (int j = tid; j < V; j += kT) s += expf(logit(j) - M);
    const float lnS = logf(block_reduce_sum(s, sm_f));
    auto log_prob = [&](int j) {
      float lp = __fadd_rn(__fadd_rn(logit(j), -M), -lnS);
      if (a.round_bf16) lp = bf16_round(lp);
      return (double)lp;
    };
    // ---- schedule scalars are fp32 tensors in the reference; only x is double
    const int64_t mc_off = (row / a.seq_len) * a.mc_seq_stride + (int64_t)step * a.mc_step_stride;
    const float mc_t = a.mc_t ? a.mc_t[mc_off] : 0.f, mc_s = a.mc_s ? a.mc_s[mc_off] : 0.f;
    double dk = 0., q_mask = 0., d_alpha = 0., c0 = 0., zc = 0., den = 1.;
    if (absorbing) {
      dk = (double)__fadd_rn(mc_t, -mc_s) / (double)mc_t;      // x (mc_t - mc_s) / mc_t
      q_mask = (double)mc_s / (double)mc_t;
    } else if (!ar) {
      const float alpha_s = 1.f - mc_s, alpha_t = 1.f - mc_t;
      const float alpha_ts = __fdiv_rn(alpha_t, alpha_s);
      d_alpha = (double)(alpha_s - alpha_t);
      c0 = (double)__fmul_rn(__fmul_rn(1.f - alpha_ts, 1.f - alpha_s), __fdiv_rn(1.f, (float)V));
      const double xz = exp(log_prob(z));
      const double atV = (double)__fmul_rn(alpha_t, (float)V);
      zc = atV * xz + (double)(alpha_ts - alpha_t);
      den = atV * xz + (double)(1.f - alpha_t);
    }
    double best = -INFINITY;
    int bidx = 0x7fffffff;
    float* qo = a.q_out ? a.q_out + row * a.ld_q : nullptr;
    for (int j = tid; j < V; j += kT) {
      const double lp = log_prob(j);
      double score;
      if (ar) {
        // argmax(log_probs + Gumbel) (diffusion.py:1038); injected G or Philox -log(E)
        const double g = pn ? (double)pn[j] : -log(1e-10 - log((double)philox_u(j, row, step, key) + 1e-10));
        score = lp + g;
        if (qo) qo[j] = (float)exp(lp);
      } else {
        double q;
        if (absorbing) q = (j == a.mask_index) ? q_mask : exp(lp) * dk;
        else q = (d_alpha * exp(lp) + c0 + (j == z ? zc : 0.)) / den;
        if (qo) qo[j] = (float)q;
        if (a.select == DDG_SELECT_ARGMAX) {
          score = q;
        } else {
          const double u = pn ? (double)pn[j] : (double)philox_u(j, row, step, key);
          score = q / (1e-10 - log(u + 1e-10));       // diffusion.py:1752-1755 in float64
        }
      }
      if (score > best) { best = score; bidx = j; }
    }
    block_argmax(best, bidx, sm_d, sm_i);
    if (tid == 0) a.xs[row] = bidx == 0x7fffffff ? 0 : bidx;
    __syncthreads();
  }
}

// ------------------------------------------------------------------ low confidence
// One CTA per row; the row of probabilities lives in shared memory.
__global__ void __launch_bounds__(1024)
low_confidence_kernel(const float* __restrict__ probs, int64_t ld, int64_t rows, int V, float threshold,
                      const float* __restrict__ noise, int64_t ld_noise, int select, uint64_t seed,
                      const uint64_t* seed_ptr, int step_arg, const int32_t* step_ptr, const int64_t* xt,
                      int carry_mask_index, int64_t* xs) {
  extern __shared__ float p[];                    // V floats
  __shared__ double sm_d[32];
  __shared__ int sm_i[32];
  __shared__ int sm_cnt[32];
  const int tid = threadIdx.x, nt = blockDim.x;
  const int step = step_ptr ? *step_ptr : step_arg;
  const uint64_t seed64 = seed_ptr ? *seed_ptr : seed;
  const uint2 key = make_uint2((uint32_t)seed64, (uint32_t)(seed64 >> 32));
  const int per = (V + nt - 1) / nt;              // contiguous chunk per thread (index-ordered tie ranks)
  const int j_lo = min(tid * per, V), j_hi = min(j_lo + per, V);

  for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
    if (xt && carry_mask_index >= 0) {            // absorbing carry-over happens after the draw
      const int z = (int)xt[row];
      if (z != carry_mask_index) {
        if (tid == 0) xs[row] = z;
        continue;
      }
    }
    const float* pr = probs + row * ld;
    __syncthreads();
    for (int j = tid; j < V; j += nt) p[j] = pr[j];
    __syncthreads();
    // ---- smallest value v* whose inclusion (with everything below it) exceeds the threshold
    auto mass_le = [&](float v) {
      double s = 0.;
      for (int j = tid; j < V; j += nt) s += p[j] <= v ? (double)p[j] : 0.;
      return block_reduce_sum(s, sm_d);
    };
    const double tau = (double)threshold;
    uint32_t lo = 0u, hi = 0x7F800000u;           // bit patterns of non-negative floats are ordered
    const bool all_kept = mass_le(__uint_as_float(hi)) <= tau;
    if (!all_kept) {
      while (lo < hi) {                           // lower bound of {b : mass_le(b) > tau}
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if (mass_le(__uint_as_float(mid)) > tau) hi = mid; else lo = mid + 1;
      }
    }
    const float vstar = all_kept ? INFINITY : __uint_as_float(lo);
    // ---- ties at v*: the first k by index stay, k = floor((tau - mass below) / v*)
    int k_ties = 0;
    if (!all_kept) {
      double s = 0.;
      for (int j = tid; j < V; j += nt) s += p[j] < vstar ? (double)p[j] : 0.;
      const double below = block_reduce_sum(s, sm_d);
      const double room = (tau - below) / (double)vstar;
      k_ties = vstar > 0.f ? (room >= 2147483647. ? 2147483647 : (int)floor(room)) : 0x7fffffff;
      if (k_ties < 0) k_ties = 0;
    }
    int my_ties = 0;
    if (!all_kept)
      for (int j = j_lo; j < j_hi; ++j) my_ties += p[j] == vstar;
    // exclusive prefix of tie counts over threads (thread order == index order)
    int tie_rank0 = 0;
    if (!all_kept) {
      int v = my_ties;
      const int lane = tid & 31, w = tid >> 5;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
      }
      __syncthreads();
      if (lane == 31) sm_cnt[w] = v;
      __syncthreads();
      int base = 0;
      for (int i = 0; i < w; ++i) base += sm_cnt[i];
      tie_rank0 = base + v - my_ties;
    }
    // ---- race / arg-max over the kept entries (probs / E in fp32, diffusion.py:1748-1755);
    // AR (`_sample_token`, diffusion.py:1030-1038): argmax of log p + G over the kept entries
    const bool gumbel = select == DDG_SELECT_GUMBEL && noise != nullptr;
    double best = gumbel ? -INFINITY : 0.;        // dropped entries score 0 / -inf: none kept -> index 0
    int bidx = 0x7fffffff;
    int rank = tie_rank0;
    for (int j = j_lo; j < j_hi; ++j) {
      const float pj = p[j];
      bool keep = all_kept || pj < vstar;
      if (!all_kept && pj == vstar) { keep = rank < k_ties; ++rank; }
      if (!keep || !(pj > 0.f)) continue;
      float score = pj;
      if (gumbel) {
        score = __fadd_rn(logf(pj), noise[row * ld_noise + j]);
      } else if (select != DDG_SELECT_ARGMAX) {   // Gumbel-max without injected noise == the race
        const float u = noise ? noise[row * ld_noise + j] : philox_u(j, row, step, key);
        score = __fdiv_rn(pj, __fadd_rn(1e-10f, -logf(__fadd_rn(u, 1e-10f))));
      }
      if ((double)score > best) { best = (double)score; bidx = j; }
    }
    block_argmax(best, bidx, sm_d, sm_i);
    if (tid == 0) xs[row] = bidx == 0x7fffffff ? 0 : bidx;
  }
}

}  // namespace

int launch_sample_f64(const ddg_sample_step_args& a, cudaStream_t s) {
  DDG_REQUIRE(a.guidance == DDG_GUIDE_NONE, DDG_ERR_UNSUPPORTED,
              "use_float64 is only honoured without guidance (reference diffusion.py:1393); not implemented");
  DDG_REQUIRE(a.seq_changed == nullptr, DDG_ERR_UNSUPPORTED, "seq_changed is not supported with compute_f64");
  int64_t grid = (int64_t)sm_count() * 8;
  if (grid > a.rows) grid = a.rows;
  if (a.logits_dtype == DDG_BF16)
    sample_f64_kernel<__nv_bfloat16><<<(unsigned)grid, kT, 0, s>>>(a);
  else
    sample_f64_kernel<float><<<(unsigned)grid, kT, 0, s>>>(a);
  return check_cuda(cudaGetLastError(), "sample_f64 launch");
}

}  // namespace ddg

extern "C" int ddg_low_confidence_select(const float* probs, int64_t ld, int64_t rows, int32_t vocab,
                                         float threshold, const float* noise, int64_t ld_noise, int32_t select,
                                         uint64_t seed, const uint64_t* seed_ptr, int32_t step,
                                         const int32_t* step_ptr, const int64_t* xt, int32_t carry_mask_index,
                                         int64_t* xs, ddg_stream_t stream) {
  using namespace ddg;
  int rc = ddg_device_check();
  if (rc) return rc;
  DDG_REQUIRE(probs && xs && rows >= 0 && vocab > 0 && ld >= vocab, DDG_ERR_INVALID, "low_confidence: bad args");
  DDG_REQUIRE(select == DDG_SELECT_RACE || select == DDG_SELECT_ARGMAX || select == DDG_SELECT_GUMBEL,
              DDG_ERR_INVALID, "low_confidence: select must be race, argmax or gumbel");
  if (noise) DDG_REQUIRE(ld_noise >= vocab, DDG_ERR_INVALID, "low_confidence: ld_noise < vocab");
  const size_t smem = (size_t)vocab * sizeof(float);
  DDG_REQUIRE(smem <= 200 * 1024, DDG_ERR_UNSUPPORTED, "low_confidence: vocab %d does not fit shared memory", vocab);
  if (rows == 0) return DDG_OK;
  static bool attr_done[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_done[dev]) {
    DDG_CUDA(cudaFuncSetAttribute(low_confidence_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_done[dev] = true;
  }
  const int threads = vocab >= 8192 ? 1024 : 256;
  const int per_sm = smem > 100 * 1024 ? 1 : 2;
  int64_t grid = (int64_t)sm_count() * per_sm;
  if (grid > rows) grid = rows;
  low_confidence_kernel<<<(unsigned)grid, threads, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
    probs, ld, rows, vocab, threshold, noise, ld_noise, select, seed, seed_ptr, step, step_ptr, xt,
    carry_mask_index, xs);
  return check_cuda(cudaGetLastError(), "low_confidence launch");
}
