/* ddg_b200 -- C ABI of the B200-native reverse-diffusion sampling hot path.
 *
 * Drop-in boundary for the per-step update of denaxen/discrete-diffusion-guidance
 * (`Diffusion._ddpm_denoise / _cfg_denoise / _cbg_denoise`, `_sample_categorical`
 * and the `models/dit.py` forward they call).  The reference is pure Python and
 * has no FFI of its own; each entry point below names the reference code it
 * replaces (paths relative to the reference root).  INTEGRATION.md shows the
 * ctypes binding a reference maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless the
 *     name ends in `_host`; tensors are row-major with explicit leading dims.
 *   - every call returns 0 on success or a negative DDG_ERR_* code and never
 *     throws; ddg_last_error() returns a thread-local message.
 *   - every launch is asynchronous on the caller's stream and graph-capturable:
 *     no allocation, no synchronisation, no host reads of device memory.
 *   - there is no CPU fallback: on a device that is not sm_100 every compute
 *     entry point returns DDG_ERR_UNSUPPORTED.
 */
#ifndef DDG_B200_H_
#define DDG_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DDG_ABI_VERSION 6
#define DDG_API __attribute__((visibility("default")))

typedef void* ddg_stream_t; /* cudaStream_t */

enum ddg_status {
  DDG_OK = 0,
  DDG_ERR_INVALID = -1,     /* bad argument (shape, alignment, enum)   */
  DDG_ERR_CUDA = -2,        /* CUDA runtime / driver error             */
  DDG_ERR_UNSUPPORTED = -3  /* not an sm_100 device / unsupported size */
};

enum ddg_dtype { DDG_F32 = 0, DDG_F16 = 1, DDG_BF16 = 2 };

/* config.diffusion (reference configs/config.yaml:14) */
enum ddg_diffusion {
  DDG_ABSORBING = 0,
  DDG_UNIFORM = 1,
  DDG_AR = 2 /* parameterization=ar: select straight from the log-probs */
};
/* config.guidance.method (reference configs/guidance/{cfg,cbg}.yaml) */
enum ddg_guidance { DDG_GUIDE_NONE = 0, DDG_GUIDE_CFG = 1, DDG_GUIDE_CBG = 2 };
/* categorical selection (reference diffusion.py:1748-1755, 1024-1038) */
enum ddg_select {
  DDG_SELECT_RACE = 0,   /* argmax_j p_j / E_j, E = 1e-10 - log(U + 1e-10) */
  DDG_SELECT_ARGMAX = 1, /* config.eval.argmax_sampling                    */
  DDG_SELECT_GUMBEL = 2  /* AR: argmax_j logp_j + G_j with G injected      */
};

DDG_API int ddg_abi_version(void);
DDG_API const char* ddg_last_error(void);
/* 0 if the current device is sm_100 (B200); DDG_ERR_UNSUPPORTED otherwise. */
DDG_API int ddg_device_check(void);
DDG_API int ddg_sm_count(void);

/* ------------------------------------------------------------------------
 * Fused per-token sampler.
 * Replaces, for one reverse step, everything after the backbone forward in
 *   Diffusion.forward            diffusion.py:356-365 (+ _subs_parameterization :321-337)
 *   Diffusion._compute_posterior diffusion.py:367-394
 *   Diffusion._ddpm_denoise      diffusion.py:1393-1420
 *   Diffusion._cfg_denoise       diffusion.py:1456-1509
 *   Diffusion._cbg_denoise       diffusion.py:1587-1621
 *   Diffusion._sample_categorical diffusion.py:1748-1755
 * One launch reads each logits row once from HBM and writes the next ids.
 * ---------------------------------------------------------------------- */
typedef struct ddg_sample_step_args {
  /* inputs */
  const void* logits_a;   /* (rows, ld_a) raw backbone logits: the conditional
                             (or only) forward                               */
  const void* logits_b;   /* (rows, ld_b) unconditional forward; CFG only     */
  const float* clf;       /* (rows, ld_clf) classifier log-prob term; CBG only */
  const float* noise;     /* (rows, ld_noise) injected U in [0,1) (RACE) or
                             Gumbel G (GUMBEL); NULL -> in-kernel Philox      */
  const int64_t* xt;      /* (rows,) current ids z_t                          */
  const float* mc_t;      /* move_chance_t, see mc_* strides                  */
  const float* mc_s;      /* move_chance_s                                    */
  const int32_t* step_ptr; /* optional device step counter (graph replay)     */
  const uint64_t* seed_ptr; /* optional device Philox key (overrides `seed`)   */
  /* outputs */
  int64_t* xs;            /* (rows,) next ids z_s                             */
  float* q_out;           /* optional (rows, ld_q): q_xs, the reference's 2nd
                             return (softmax-normalised for D-CFG on uniform
                             and for D-CBG, the posterior as formed otherwise) */
  /* sizes */
  int64_t rows;           /* B * L                                            */
  int64_t ld_a, ld_b, ld_clf, ld_noise, ld_q;
  int32_t seq_len;        /* L: row -> sequence = row / L                     */
  int32_t vocab;          /* V                                                */
  int32_t mc_seq_stride;  /* element stride per sequence (0: one scalar)      */
  int32_t mc_step_stride; /* element stride per step (0: no step table);
                             step = *step_ptr if step_ptr else `step`         */
  /* modes */
  int32_t logits_dtype;   /* DDG_F32 | DDG_BF16                               */
  int32_t diffusion;      /* ddg_diffusion                                    */
  int32_t guidance;       /* ddg_guidance                                     */
  int32_t subs;           /* 1: parameterization=subs (zero-mask + carry-over) */
  int32_t mask_logit;     /* 1: logits[mask_index] += -1e6 (subs or
                             config.subs_masking)                             */
  int32_t mask_index;
  int32_t round_bf16;     /* 1: round log-softmax and exp to bf16 the way the
                             reference does on a GPU (SURVEY 8a' item 4)      */
  int32_t select;         /* ddg_select                                       */
  int32_t fast_math;      /* 1: ex2/lg2.approx intrinsics (production);
                             0: expf/logf, unfused mul/add (parity)           */
  float gamma;            /* guidance.gamma                                   */
  uint64_t seed;          /* Philox key                                       */
  int32_t step;           /* Philox counter word 3 when step_ptr == NULL      */
  int32_t stats_chunks;   /* number of 128-column chunks in stats_a/b         */
  /* optional log-softmax partials written by ddg_gemm (see ddg_gemm_args):
   * with them (bf16 logits, 16-byte-aligned rows, no q_out) the kernel reads
   * every logit exactly once.                                               */
  const void* stats_a;
  const void* stats_b;
  int64_t stats_ld;
  /* Philox noise only (noise == NULL), select == DDG_SELECT_RACE:
   *   0  one exponential per vocabulary entry, argmax_j q_j / E_j -- the
   *      reference's race as written (diffusion.py:1752-1755);
   *   8  hierarchical race: one exponential per group of 8 consecutive
   *      columns, argmax over groups of (sum_G q) / E_G, then an inverse-CDF
   *      pick inside the winning group.  Identical distribution (min of
   *      independent exponentials), 8x less noise work; needs the stats_*
   *      partials, bf16 logits and fast_math.
   *  64  the same with 64-column groups made of the 8 vectors (of 8 columns)
   *      one thread of the kernel visits in a sweep: columns
   *      8 (t + 128 (8 k + s)) + e, s, e = 0..7, for group k * 128 + t.  Same
   *      distribution again (a group may be any set of columns); the per-group
   *      work is paid once per 64 columns.  Other values: DDG_ERR_INVALID.    */
  int32_t race_group;
  /* 1: sampling.use_float64 (diffusion.py:1393-1394, 1110-1111, 1134-1135): the
   * log-probs keep the forward's precision, everything after them runs in
   * double.  Like the reference only without guidance (absorbing / uniform /
   * AR); with guidance -> DDG_ERR_UNSUPPORTED ("not implemented").            */
  int32_t compute_f64;
  /* ABI v3.  D-CBG with the EXACT classifier term on a GPU: get_log_probs returns
   * a bf16 tensor (classifier.py:325-327), so `gamma * classifier_log_prob`
   * (python float x bf16 tensor, diffusion.py:1607, 1612) is rounded to bf16 before
   * the fp32 posterior is added.  1 = apply that rounding to gamma * clf; 0 = the
   * first-order term (fp32 in the reference, :1540-1547) or fp32 parity runs.  */
  int32_t clf_round_bf16;
  int32_t reserved0;
  /* ABI v4.  Optional (n_seqs,) flags: set to 1 for the sequence of every row whose
   * token changes in this step (xs != xt).  The MDLM loop (time_conditioning =
   * false) recomputes the denoiser only for the flagged sequences: an unchanged
   * sequence has unchanged logits -- the per-sequence form of the reference's
   * `sampling.use_cache` (diffusion.py:1371-1374), result-identical.           */
  int32_t* seq_changed;
} ddg_sample_step_args;

DDG_API int ddg_sample_step(const ddg_sample_step_args* args, ddg_stream_t stream);

/* eval.low_confidence_sampling (diffusion.py:1738-1755): from the posterior
 * `probs` (rows, ld) fp32 -- what _sample_categorical receives, e.g. the q_out
 * of ddg_sample_step -- keep the entries whose cumulative mass in ascending
 * probability order is <= threshold (ties in index order, i.e. a stable sort),
 * zero the rest, then draw by the exponential race (noise: injected U or NULL
 * for the Philox stream of ddg_sample_step) or take the arg-max; select =
 * DDG_SELECT_GUMBEL with injected Gumbel noise is the AR sampler's
 * `_sample_token` (diffusion.py:1024-1038): argmax of log p + G over the kept
 * entries (without injected noise it is the same race).  xt != NULL
 * and carry_mask_index >= 0: rows with xt != carry_mask_index are copied
 * through afterwards (absorbing carry-over, diffusion.py:1414-1418).          */
DDG_API int ddg_low_confidence_select(const float* probs, int64_t ld, int64_t rows, int32_t vocab,
                              float threshold, const float* noise, int64_t ld_noise,
                              int32_t select, uint64_t seed, const uint64_t* seed_ptr,
                              int32_t step, const int32_t* step_ptr, const int64_t* xt,
                              int32_t carry_mask_index, int64_t* xs, ddg_stream_t stream);

/* Log-probs the way Diffusion.forward returns them (diffusion.py:356-365,
 * 321-337): optional mask-logit, subs carry-over rows, log_softmax; fp32 out.
 * Not on the sampling path (ddg_sample_step fuses it); provided for the
 * `forward()` API and for logits-parity tests. */
DDG_API int ddg_log_probs(const void* logits, int64_t ld, int32_t logits_dtype,
                  const int64_t* xt, int64_t rows, int32_t vocab,
                  int32_t subs, int32_t mask_logit, int32_t mask_index,
                  int32_t round_bf16, float* out, int64_t ld_out,
                  ddg_stream_t stream);

/* Prior sample (diffusion.py:1006-1013): absorbing -> mask_index everywhere,
 * uniform -> Philox randint(0, V).  Counter word 3 = 0xFFFFFFFF. */
DDG_API int ddg_sample_prior(int64_t* xt, int64_t n, int32_t diffusion, int32_t vocab,
                     int32_t mask_index, uint64_t seed, ddg_stream_t stream);

/* *step_ptr += 1 (single thread); closes one captured step of the loop. */
DDG_API int ddg_step_advance(int32_t* step_ptr, ddg_stream_t stream);

/* ------------------------------------------------------------------------
 * DiT denoiser forward (models/dit.py:426-468) building blocks.
 * ---------------------------------------------------------------------- */

/* Conditioning vector c = silu(sigma_map(sigma)) [+ silu(cond_map[label])],
 * fp32 (models/dit.py:433-446, 155-213).  n rows; label may be NULL (no
 * cond_map term); sigma may be NULL (AR: c = silu(cond_map[label])).
 * Row r reads sigma[r * sigma_row_stride + step * sigma_step_stride] with
 * step = *step_ptr (0 when step_ptr is NULL). */
DDG_API int ddg_cond_embed(const float* sigma, int32_t sigma_row_stride,
                   const int32_t* step_ptr, int32_t sigma_step_stride,
                   const int64_t* label, int32_t n,
                   int32_t cond_dim, int32_t freq_dim,
                   const float* w0, const float* b0,  /* (cond_dim,freq_dim) */
                   const float* w2, const float* b2,  /* (cond_dim,cond_dim) */
                   const float* label_table,          /* (classes+1,cond_dim) */
                   float* c_out, ddg_stream_t stream);

/* adaLN_modulation linears (models/dit.py:269-271, 361-362) evaluated the way
 * autocast(dtype) does: operands rounded to `dtype`, fp32 accumulate, one
 * output rounding.  out[r, :] = round(c[r,:] @ W^T + b), W (n_out, cond_dim)
 * fp32 master weights.  Outputs whose bit (col / chunk) % n_chunks of
 * `one_plus_mask` is set store round(1 + value) instead
 * (the `1 + scale` of modulate, formed in `dtype`, models/dit.py:47-50). */
DDG_API int ddg_adaln(const float* c, int32_t n_rows, int32_t cond_dim,
              const float* w, const float* b, int32_t n_out,
              int32_t dtype, int32_t chunk, uint32_t one_plus_mask,
              int32_t n_chunks, void* out, int64_t ld_out,
              ddg_stream_t stream);

/* LayerNorm (fp32 stats, no bias) * weight, then x*(1+scale)+shift, rounded
 * once to the GEMM operand dtype (models/dit.py:131-139, 47-50, 280, 321,
 * 363).  `ids` != NULL additionally gathers x = embedding[ids[row %
 * ids_period]] and writes it to `x` first (models/dit.py:337, 449; the period
 * lets the cond/uncond halves of a batched CFG forward share one id buffer).  shift / one_plus_scale are rows of
 * the adaLN table: row = token_row / rows_per_mod.  shift == NULL: no adaLN. */
DDG_API int ddg_ln_modulate(float* x, int64_t rows, int32_t hidden,
                    const int64_t* ids, int64_t ids_period,
                    const float* embedding,
                    const float* ln_weight,
                    const void* shift, const void* one_plus_scale,
                    int64_t mod_stride, int64_t rows_per_mod,
                    int32_t dtype, void* out, ddg_stream_t stream);

/* The same over a dynamic batch: only the first *live_seqs sequences of
 * `seq_len` rows exist (`rows` is the upper bound); with the id gather,
 * seq_map[s] is the sequence of the id buffer that compact sequence s holds.  */
DDG_API int ddg_ln_modulate_live(float* x, int64_t rows, int32_t hidden,
                         const int64_t* ids, int64_t ids_period,
                         const float* embedding, const float* ln_weight,
                         const void* shift, const void* one_plus_scale,
                         int64_t mod_stride, int64_t rows_per_mod,
                         int32_t dtype, void* out, const int32_t* live_seqs,
                         int32_t seq_len, const int32_t* seq_map, ddg_stream_t stream);

/* seq_map[0..*count) = the indices s with flags[s] != 0, ascending; the flags are
 * cleared.  Builds the dynamic batch of the MDLM loop from the `seq_changed` flags
 * of ddg_sample_step (no host round trip: graph-capturable).  all_total > 0: the
 * whole-batch form of `sampling.use_cache` (diffusion.py:1371-1374) instead --
 * *count = all_total if any flag is set, else 0; seq_map is not written.       */
DDG_API int ddg_seq_compact(int32_t* flags, int32_t n_seqs, int32_t* seq_map, int32_t* count,
                    int32_t all_total, ddg_stream_t stream);

/* GEMM epilogues (models/dit.py:284-293, 312-316, 242-245/324-325, 359-366) */
enum ddg_epilogue {
  DDG_EPI_STORE = 0,        /* out = round(acc [+ bias])            16-bit   */
  DDG_EPI_QKV_ROPE = 1,     /* round, then NeoX rotary on q,k heads 16-bit   */
  DDG_EPI_BIAS_GELU = 2,    /* round(gelu_tanh(round(acc + bias)))  16-bit   */
  DDG_EPI_GATE_RESIDUAL = 3, /* x += round(gate * round(acc [+ bias])) fp32  */
  DDG_EPI_ACCUM_F32 = 4,    /* out = (accumulate ? out : 0) + acc, fp32, no
                               rounding (classifier dgrad outputs)            */
  DDG_EPI_GELU_BWD = 5      /* round(round(acc) * gelu_tanh'(aux))  16-bit:
                               the fc2 input-gradient GEMM of the first-order
                               D-CBG backward times GELU'(pre-activation)     */
};

typedef struct ddg_gemm_args {
  const void* a;        /* (M, K) activations, 16-bit `dtype`, row-major     */
  const void* w;        /* (N, K) nn.Linear weight, 16-bit `dtype`           */
  void* out;            /* (M, ld_out) 16-bit, or the fp32 residual stream.
                           16-bit outputs leave through TMA stores that clip at
                           16-byte units: with N % 8 != 0 the columns up to the
                           next multiple of 8 are written too (ld_out covers
                           them: it is a multiple of 8)                        */
  const void* bias;     /* (N,) `dtype` or NULL; readable up to the next
                           multiple of 256 columns                           */
  const void* gate;     /* GATE_RESIDUAL: adaLN gate rows (`dtype`) or NULL  */
  const void* rope_cos; /* QKV_ROPE: (L, head_dim/2) `dtype`                 */
  const void* rope_sin;
  int64_t lda, ldw, ld_out;
  int64_t gate_stride;  /* elements between gate rows                        */
  int64_t rows_per_mod; /* gate row = m / rows_per_mod                       */
  int32_t M, N, K;
  int32_t dtype;        /* DDG_F16 | DDG_BF16                                */
  int32_t epilogue;     /* ddg_epilogue                                      */
  int32_t seq_len;      /* QKV_ROPE: position = m % seq_len                  */
  int32_t head_dim;     /* QKV_ROPE: 64                                      */
  int32_t rope_cols;    /* QKV_ROPE: columns [0, rope_cols) are q,k heads    */
  /* STORE only, optional: log-softmax partials of the stored (rounded) rows.
   * stats[(n / 128) * stats_ld + m] = (max_j v_mj, sum_j 2^((v_mj - max) log2 e))
   * over the 128-column chunk, as float2; column stats_mask_col (or -1) is
   * left out.  ddg_sample_step uses them to skip its normaliser pass.       */
  void* stats;
  int64_t stats_ld;
  int32_t stats_mask_col;
  int32_t accumulate;   /* ACCUM_F32: 1 = add to the existing output         */
  int32_t rope_pos0;    /* QKV_ROPE: position = rope_pos0 + m % seq_len (the
                           incremental AR decode passes the current position) */
  int32_t reserved;
  /* ABI v4, dynamic batch (both optional, need seq_len > 0 dividing M):
   * live_seqs  device scalar: only the first *live_seqs sequences (seq_len rows
   *            each) of `a` exist in this launch; M stays the upper bound the
   *            buffers and the grid are sized for (graph-capturable).
   * out_seq_map (STORE, seq_len % 32 == 0): rows of compact sequence s (and
   *            their `stats`) are written at output sequence out_seq_map[s].  */
  const int32_t* live_seqs;
  const int32_t* out_seq_map;
  /* ABI v5, GATE_RESIDUAL only, optional (ln_weight != NULL): the LayerNorm + adaLN
   * modulate that follows every residual add of a DiT block (models/dit.py:312-321,
   * 256-259: `modulate(norm(x), shift, scale)`) fused into this launch, so the
   * updated fp32 rows are normalised while they are still on chip instead of being
   * read back from HBM by ddg_ln_modulate:
   *   ln_out[m, :] = round_ln_dtype(((x_new[m] - mean) * rstd) * ln_weight
   *                                  [* ln_scale[m / rows_per_mod] + ln_shift[...]])
   * with x_new the row this launch has just written to `out` (same values, same
   * operation order as ddg_ln_modulate; mean / variance are combined from per-128-
   * column partials, so rstd may differ from the stand-alone kernel in the last ulp).
   * Needs N % 128 == 0, N <= 1024, ld_out == N, rows_per_mod % 32 == 0.
   * ln_work: device scratch of ddg_gemm_ln_work_bytes(M, N) bytes, 256-byte aligned,
   * ZERO before its first use.  Its layout depends on N only (per 128-row block: a
   * counter word and the block's per-row partials), and every launch leaves it
   * ready for the next one, so one buffer sized for the largest M serves every
   * launch with the same N on a stream (not concurrent launches).              */
  const float* ln_weight;  /* (N,) fp32                                        */
  const void* ln_shift;    /* `ln_dtype` rows, ln_mod_stride apart, or NULL    */
  const void* ln_scale;    /* 1 + scale rows, same layout (NULL with ln_shift) */
  void* ln_out;            /* (M, N) `ln_dtype`, dense                         */
  void* ln_work;
  int64_t ln_mod_stride;
  int32_t ln_dtype;        /* DDG_F16 | DDG_BF16                               */
  int32_t reserved1;
  /* GATE_RESIDUAL only, optional: (M, ld_out) fp32 rows the delta is added to,
   * out = residual + round(gate * round(acc [+ bias])); NULL = `out` in place.
   * Lets a caller keep every block's input rows (the first-order D-CBG backward
   * needs them) without copying the residual stream.                          */
  const float* residual;
  /* ABI v6, optional 16-bit side tensor (M, ld_aux), `dtype`, 16-byte aligned rows:
   *   BIAS_GELU  output: aux = round(acc + bias), the pre-activation the GELU is
   *              applied to (models/dit.py:242-245) -- kept for the backward of
   *              first-order D-CBG instead of recomputing the fc1 GEMM there;
   *   GELU_BWD   input (required): the pre-activation saved by the forward.     */
  void* aux;
  int64_t ld_aux;
} ddg_gemm_args;

/* tcgen05 / TMEM / TMA GEMM: out = epilogue(a @ w^T). */
DDG_API int ddg_gemm(const ddg_gemm_args* args, ddg_stream_t stream);

/* Bytes of `ln_work` scratch a fused-LayerNorm ddg_gemm over up to M rows needs. */
DDG_API int64_t ddg_gemm_ln_work_bytes(int32_t M, int32_t N);

/* Softmax attention over whole sequences (models/dit.py:303-307):
 * qkv (B*L, 3*H*D) 16-bit with column order (three, head, dim)
 * (models/dit.py:285-289) -> out (B*L, H*D).  D must be 64, L <= 256 and a
 * multiple of 16. */
DDG_API int ddg_attention(const void* qkv, void* out, int32_t batch, int32_t seq_len,
                  int32_t n_heads, int32_t head_dim, int32_t causal,
                  int32_t dtype, ddg_stream_t stream);

/* ddg_attention over a dynamic batch: only the first *live_seqs sequences exist
 * (`batch` is the upper bound).                                                */
DDG_API int ddg_attention_live(const void* qkv, void* out, int32_t batch, int32_t seq_len,
                       int32_t n_heads, int32_t head_dim, int32_t causal,
                       int32_t dtype, const int32_t* live_seqs, ddg_stream_t stream);

/* One decode step of causal attention with a KV cache (AR sampler,
 * diffusion.py:1085-1186: the reference re-encodes the whole prefix for the DiT
 * backbone, :1132).  qkv_cache is the (batch*cache_len, 3*H*D) tensor ddg_gemm
 * QKV_ROPE fills row by row (row b*cache_len + t = token t of sequence b);
 * the query is row `pos`, the keys/values rows 0..pos.  out (batch, H*D). */
DDG_API int ddg_attention_decode(const void* qkv_cache, void* out, int32_t batch, int32_t cache_len,
                         int32_t pos, int32_t n_heads, int32_t head_dim, int32_t dtype,
                         ddg_stream_t stream);

/* fp32 -> 16-bit cast of a weight tensor (what autocast does per call;
 * done once at load time here). */
DDG_API int ddg_cast(const float* src, void* dst, int64_t n, int32_t dtype,
             ddg_stream_t stream);

/* Mean over the sequence + Linear(h -> classes) + log_softmax
 * (models/dit.py:538-555, classifier.py:316-327). x (n*L, hidden) fp32. */
DDG_API int ddg_classifier_head(const float* x, int32_t n, int32_t seq_len,
                        int32_t hidden, const float* w, const float* b,
                        int32_t n_classes, int32_t round_bf16,
                        float* log_probs, ddg_stream_t stream);

/* ------------------------------------------------------------------------
 * First-order D-CBG (diffusion.py:1532-1548): one classifier forward and the
 * backward of log p(y | x) w.r.t. the one-hot input.  Only activation
 * gradients are propagated.  The linears' input gradients are ddg_gemm calls on
 * transposed weights (16-bit STORE, or ACCUM_F32 for the final G @ E^T); the
 * entry points below are the other pieces of models/dit.py:221-325, 470-556
 * run backwards.
 * ---------------------------------------------------------------------- */
/* head: log_probs (n, classes) as ddg_classifier_head, and
 * g_x[(n*L), hidden] = d log_probs[:, target] / d x  (mean-pool + Linear).   */
DDG_API int ddg_classifier_head_backward(const float* x, int32_t n, int32_t seq_len, int32_t hidden,
                                 const float* w, const float* b, int32_t n_classes, int32_t target,
                                 int32_t round_bf16, float* log_probs, float* g_x, ddg_stream_t stream);
/* out = dtype(g * gate[row / rows_per_mod]) (gate == NULL: plain cast): the
 * gradient entering a gated branch (models/dit.py:312-316, 324-325).          */
DDG_API int ddg_gate_mul(const float* g, const void* gate, int64_t gate_stride, int64_t rows_per_mod,
                 int64_t rows, int32_t hidden, int32_t dtype, void* out, ddg_stream_t stream);
/* out = g_h * gelu_tanh'(pre)  (models/dit.py:243).                          */
DDG_API int ddg_gelu_backward(const void* g_h, const void* pre, void* out, int64_t n, int32_t dtype,
                      ddg_stream_t stream);
/* g_resid += (d [LN(x) * w * one_plus_scale + shift] / d x)^T g_a
 * (models/dit.py:131-139, 47-50); statistics are recomputed from x.          */
DDG_API int ddg_ln_modulate_backward(const float* x, const void* g_a, int64_t rows, int32_t hidden,
                             const float* ln_weight, const void* one_plus_scale, int64_t mod_stride,
                             int64_t rows_per_mod, int32_t dtype, float* g_resid, ddg_stream_t stream);
/* qkv: the forward's (rotated) QKV tensor; g_out: gradient of the attention
 * output; g_qkv: gradient of the QKV GEMM output (rotary undone)
 * (models/dit.py:284-307).                                                    */
DDG_API int ddg_attention_backward(const void* qkv, const void* g_out, const void* rope_cos, const void* rope_sin,
                           int32_t batch, int32_t seq_len, int32_t n_heads, int32_t head_dim,
                           int32_t causal, int32_t dtype, void* g_qkv, ddg_stream_t stream);
/* src = hi + lo with hi = dtype(src), lo = dtype(src - hi): three 16-bit
 * GEMMs (hi*hi + lo*hi + hi*lo) then reproduce an fp32 matmul to ~2^-16.      */
DDG_API int ddg_split16(const float* src, void* hi, void* lo, int64_t n, int32_t dtype, ddg_stream_t stream);
/* term[row, v] = grad[row, v] - grad[row, xt[row]] + log_probs[row / L, target]
 * (diffusion.py:1541-1548).                                                   */
DDG_API int ddg_cbg_first_order(const float* grad, int64_t ld_grad, const int64_t* xt, const float* log_probs,
                        int32_t n_classes, int32_t target, int32_t seq_len, int32_t vocab, int64_t rows,
                        float* term, int64_t ld_term, ddg_stream_t stream);

/* Exact D-CBG candidate batch (diffusion.py:1551-1585): for candidate c of
 * chunk [c0, c0+n): seq b = c / (L*V), l = (c / V) % L, v = c % V;
 * ids_out[c - c0, :] = xt[b, :] with position l replaced by v. */
DDG_API int ddg_cbg_candidates(const int64_t* xt, int32_t seq_len, int32_t vocab,
                       int64_t c0, int64_t n, int64_t* ids_out,
                       ddg_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DDG_B200_H_ */
