"""Host side of the DiT denoiser / classifier forward.

Mirrors the interface of the reference backbone (`models/dit.py`: `DIT`
:369-468, `DITClassifier` :470-556) -- `DIT(config, vocab_size)`,
`load_state_dict`, `forward(indices, sigma, cond=None)` -> logits (B, L, V) --
but every per-token operation is a launch of the sm_100a kernels behind the C
ABI (gather+LayerNorm+modulate, tcgen05 GEMMs with fused epilogues, attention).
PyTorch only owns the device buffers.

Precision follows the reference on a GPU (SURVEY.md 8a'): fp32 residual
stream / LayerNorm / conditioning, fp16 blocks for the denoiser
(`models/dit.py:454`), bf16 blocks for the classifier (:531), bf16 final layer
and logits (:463).
"""
from __future__ import annotations

import dataclasses
import typing

import torch

from . import _lib
from ._lib import check, lib, ptr, stream_ptr

_BIG = 1 << 40
_FREQ_DIM = 256  # TimestepEmbedder.frequency_embedding_size, models/dit.py:160


def _pad_to(n, m):
  return (n + m - 1) // m * m


@dataclasses.dataclass
class ModSpec:
  """Where the conditioning rows of one forward come from (one `c` row per
  `seqs_per_mod` consecutive sequences)."""
  n_mod: int
  seqs_per_mod: int
  sigma: typing.Optional[torch.Tensor] = None   # f32, device
  sigma_row_stride: int = 0
  sigma_step_stride: int = 0
  step_ptr: typing.Optional[torch.Tensor] = None  # int32[1], device
  labels: typing.Optional[torch.Tensor] = None  # int64 (n_mod,), device


class DIT:
  """B200-native counterpart of `models.dit.DIT` (and, with
  `classifier=True`, of `models.dit.DITClassifier`)."""

  def __init__(self, config, vocab_size: int, device=None,
               classifier: bool = False):
    self.config = config
    self.vocab_size = vocab_size
    self.classifier = classifier
    m = config.classifier_model if classifier else config.model
    self.hidden = m.hidden_size
    self.cond_dim = m.cond_dim
    self.n_blocks = m.n_blocks
    self.n_heads = m.n_heads
    self.head_dim = self.hidden // self.n_heads
    self.causal = config.parameterization == 'ar'
    self.device = torch.device(device if device is not None else 'cuda')
    self.block_dtype = torch.bfloat16 if classifier else torch.float16
    self.head_dtype = torch.bfloat16
    self.vocab_ld = _pad_to(vocab_size, 64)   # padded logits leading dim
    self.seg_seqs = None                      # micro-batch (sequences) or None
    self.fuse_ln = True                       # LayerNorm + modulate inside the residual GEMMs
    if classifier:
      pooling = m.get('pooling', 'mean')
      if pooling != 'mean':
        raise NotImplementedError(f'`{pooling}` method not implemented.')
      self.num_classes = m.num_classes
    if self.head_dim != 64 or self.hidden % 128 != 0:
      raise NotImplementedError(
        'ddg_b200 kernels need head_dim == 64 and hidden % 128 == 0 '
        f'(got hidden={self.hidden}, n_heads={self.n_heads})')
    # the attention kernels take whole sequences of at most 256 tokens (multiples of 16); the
    # reference's default `model.length=1024` (OpenWebText) is outside this path -- say so here
    # instead of at the first forward
    L = int(m.get('length', 0) or 0)
    if L > 256 or L % 16 != 0:
      raise NotImplementedError(
        f'ddg_b200 attention kernels support model.length in multiples of 16 up to 256 '
        f'(text8 256, LM1B / Amazon 128, QM9 32); got model.length={L}: override it, e.g. '
        f'model.length=128')
    self._loaded = False
    self._ws = None
    self._ws_rows = 0
    self._ws_mods = 0
    self.ws_generation = 0    # bumped when buffers move (captured graphs go stale)
    self._rope = {}

  # ------------------------------------------------------------------ weights
  def _cast(self, t: torch.Tensor, dtype, pad_to: int = 0):
    """fp32 master tensor -> 16-bit device copy via the library's cast."""
    src = t.to(self.device, torch.float32).contiguous()
    n = src.numel()
    out = torch.zeros(max(n, pad_to), dtype=dtype, device=self.device)
    check(lib.ddg_cast(ptr(src), ptr(out), n, _lib.DTYPE_CODE[dtype],
                       stream_ptr()))
    torch.cuda.current_stream().synchronize()
    return out

  def load_state_dict(self, sd, prefix: typing.Optional[str] = None):
    _lib.require_device()
    if prefix is None:
      prefix = 'classifier_model.' if self.classifier else 'backbone.'

    def get(k):
      return sd[prefix + k]

    def f32(k):
      return get(k).to(self.device, torch.float32).contiguous()

    h, V = self.hidden, self.vocab_size
    if tuple(get('vocab_embed.embedding').shape) != (V, h):
      raise ValueError('vocab_embed.embedding shape mismatch')
    self.embedding = f32('vocab_embed.embedding')
    self.has_sigma = (prefix + 'sigma_map.mlp.0.weight') in sd
    if self.has_sigma:
      self.sm_w0, self.sm_b0 = f32('sigma_map.mlp.0.weight'), f32('sigma_map.mlp.0.bias')
      self.sm_w2, self.sm_b2 = f32('sigma_map.mlp.2.weight'), f32('sigma_map.mlp.2.bias')
    self.label_table = None
    if (prefix + 'cond_map.embedding_table.weight') in sd:
      self.label_table = f32('cond_map.embedding_table.weight')
    self.use_adaln = (prefix + 'blocks.0.adaLN_modulation.weight') in sd
    bd = self.block_dtype
    self.ln1, self.ln2 = [], []
    self.w_qkv, self.w_out, self.w_fc1, self.w_fc2 = [], [], [], []
    self.b_fc1, self.b_fc2 = [], []
    ada_w, ada_b = [], []
    for i in range(self.n_blocks):
      p = f'blocks.{i}.'
      self.ln1.append(f32(p + 'norm1.weight'))
      self.ln2.append(f32(p + 'norm2.weight'))
      self.w_qkv.append(self._cast(get(p + 'attn_qkv.weight'), bd))
      self.w_out.append(self._cast(get(p + 'attn_out.weight'), bd))
      self.w_fc1.append(self._cast(get(p + 'mlp.0.weight'), bd))
      self.w_fc2.append(self._cast(get(p + 'mlp.2.weight'), bd))
      self.b_fc1.append(self._cast(get(p + 'mlp.0.bias'), bd, _pad_to(4 * h, 256)))
      self.b_fc2.append(self._cast(get(p + 'mlp.2.bias'), bd, _pad_to(h, 256)))
      if self.use_adaln:
        ada_w.append(get(p + 'adaLN_modulation.weight'))
        ada_b.append(get(p + 'adaLN_modulation.bias'))
    if self.use_adaln:
      self.ada_w = torch.cat(ada_w, 0).to(self.device, torch.float32).contiguous()
      self.ada_b = torch.cat(ada_b, 0).to(self.device, torch.float32).contiguous()
    if self.classifier:
      self.cls_w, self.cls_b = f32('output_layer.weight'), f32('output_layer.bias')
      # first-order D-CBG: input gradients are GEMMs on the transposed weights
      # (dX = dY @ W), and the last one, G @ E^T, runs as three 16-bit GEMMs on
      # (hi, lo) halves of the fp32 operands
      def tcast(k):
        return self._cast(get(k).t().contiguous(), bd)
      self.w_qkv_t = [tcast(f'blocks.{i}.attn_qkv.weight') for i in range(self.n_blocks)]
      self.w_out_t = [tcast(f'blocks.{i}.attn_out.weight') for i in range(self.n_blocks)]
      self.w_fc1_t = [tcast(f'blocks.{i}.mlp.0.weight') for i in range(self.n_blocks)]
      self.w_fc2_t = [tcast(f'blocks.{i}.mlp.2.weight') for i in range(self.n_blocks)]
      emb_hi = torch.empty(V, h, dtype=bd, device=self.device)
      emb_lo = torch.empty(V, h, dtype=bd, device=self.device)
      check(lib.ddg_split16(ptr(self.embedding), ptr(emb_hi), ptr(emb_lo),
                            V * h, _lib.DTYPE_CODE[bd], stream_ptr()))
      # G @ E^T ~ G_hi E_hi^T + G_lo E_hi^T + G_hi E_lo^T as ONE GEMM over the concatenated
      # reduction dimension: [G_hi | G_lo | G_hi] @ [E_hi | E_hi | E_lo]^T, K = 3 h
      self.emb_cat = torch.cat([emb_hi, emb_hi, emb_lo], dim=1).contiguous()
      torch.cuda.current_stream().synchronize()
    else:
      self.ln_f = f32('output_layer.norm_final.weight')
      self.w_vocab = self._cast(get('output_layer.linear.weight'), self.head_dtype)
      self.b_vocab = self._cast(get('output_layer.linear.bias'), self.head_dtype,
                                _pad_to(V, 256))
      self.final_adaln = (prefix + 'output_layer.adaLN_modulation.weight') in sd
      if self.final_adaln:
        self.ada_f_w = f32('output_layer.adaLN_modulation.weight')
        self.ada_f_b = f32('output_layer.adaLN_modulation.bias')
    self._loaded = True
    return self

  # ------------------------------------------------------------------ buffers
  def _rope_tables(self, L: int):
    if L not in self._rope:
      # models/dit.py:82-108: fp32 cos/sin, rounded to the qkv dtype (:290-293)
      inv_freq = 1.0 / (10000 ** (torch.arange(0, self.head_dim, 2).float()
                                  / self.head_dim))
      freqs = torch.einsum('i,j->ij', torch.arange(L).float(), inv_freq)
      self._rope[L] = (freqs.cos().to(self.block_dtype).to(self.device).contiguous(),
                       freqs.sin().to(self.block_dtype).to(self.device).contiguous())
    return self._rope[L]

  def _workspace(self, rows: int, n_mod: int):
    if self._ws is None or rows > self._ws_rows or n_mod > self._ws_mods:
      rows = max(rows, self._ws_rows)
      n_mod = max(n_mod, self._ws_mods)
      h, dev = self.hidden, self.device
      seg = rows if self.seg_seqs is None else rows   # sized for the full batch
      ws = {
        'x': torch.empty(rows, h, dtype=torch.float32, device=dev),
        'a16': torch.empty(seg, h, dtype=torch.int16, device=dev),
        'att': torch.empty(seg, h, dtype=torch.int16, device=dev),
        # scratch of the fused LayerNorm epilogue (arrival counters + per-row partials):
        # zero once, every launch leaves the counters zero again
        'ln_work': torch.zeros(int(lib.ddg_gemm_ln_work_bytes(rows, h)), dtype=torch.uint8,
                               device=dev),
        'qkv': torch.empty(seg, 3 * h, dtype=torch.int16, device=dev),
        'h4': torch.empty(seg, 4 * h, dtype=torch.int16, device=dev),
        'c': torch.empty(n_mod, self.cond_dim, dtype=torch.float32, device=dev),
        'mod_blk': torch.empty(n_mod, self.n_blocks * 6 * h, dtype=torch.int16, device=dev),
        'mod_fin': torch.empty(n_mod, 2 * h, dtype=torch.int16, device=dev),
      }
      self._ws, self._ws_rows, self._ws_mods = ws, rows, n_mod
      self.ws_generation += 1
    return self._ws

  def alloc_logits(self, rows: int) -> torch.Tensor:
    """(rows, vocab_ld) bf16 logits buffer; columns >= vocab are padding."""
    return torch.empty(rows, self.vocab_ld, dtype=self.head_dtype,
                       device=self.device)

  @property
  def stats_chunks(self) -> int:
    return (self.vocab_size + 127) // 128

  def alloc_stats(self, rows: int) -> torch.Tensor:
    """(chunks, rows, 2) fp32 log-softmax partials written by the vocab GEMM
    epilogue and consumed by the single-pass sampler."""
    return torch.empty(self.stats_chunks, rows, 2, dtype=torch.float32,
                       device=self.device)

  # ------------------------------------------------------------------ forward
  def run(self, ids: torch.Tensor, n_seqs: int, L: int, mod: ModSpec,
          logits_out: typing.Optional[torch.Tensor] = None,
          ids_period: typing.Optional[int] = None,
          stats_out: typing.Optional[torch.Tensor] = None,
          stats_mask_col: int = -1,
          save_for_backward: typing.Optional[list] = None,
          live: typing.Optional[tuple] = None):
    """Launches one forward over `n_seqs` sequences of length L.  `ids` holds
    `ids_period` (default n_seqs*L) int64 token ids, reused cyclically.
    Denoiser: writes logits (rows, vocab_ld) bf16 into `logits_out`.
    Classifier: returns the fp32 residual stream (rows, hidden).

    `live = (seq_map, n_live)` (device int32 tensors) runs the forward over a
    DYNAMIC batch: only the sequences `seq_map[0..n_live)` of `ids` go through the
    network (compact activations), and their logits / partials are written back at
    their own rows of `logits_out` / `stats_out`; every launch is sized for
    `n_seqs` and reads the live count on the device (graph-capturable).
    `seq_map = None`: the live sequences are the first `n_live` ones in place (the
    whole-batch cache: n_live is n_seqs or 0)."""
    assert self._loaded, 'load_state_dict first'
    M = n_seqs * L
    h, nb = self.hidden, self.n_blocks
    period = M if ids_period is None else ids_period
    ws = self._workspace(M, mod.n_mod)
    s = stream_ptr()
    bcode = _lib.DTYPE_CODE[self.block_dtype]
    hcode = _lib.DTYPE_CODE[self.head_dtype]
    rpm = mod.seqs_per_mod * L
    mod_stride = nb * 6 * h
    if self.use_adaln:
      if mod.labels is not None and self.label_table is None:
        raise ValueError('Conditioning variable provided, but Model was not '
                         'initialized with condition embedding layer.')
      use_sigma = self.has_sigma
      check(lib.ddg_cond_embed(
        ptr(mod.sigma) if use_sigma else 0, mod.sigma_row_stride,
        ptr(mod.step_ptr), mod.sigma_step_stride, ptr(mod.labels), mod.n_mod,
        self.cond_dim, _FREQ_DIM,
        ptr(self.sm_w0) if use_sigma else 0, ptr(self.sm_b0) if use_sigma else 0,
        ptr(self.sm_w2) if use_sigma else 0, ptr(self.sm_b2) if use_sigma else 0,
        ptr(self.label_table), ptr(ws['c']), s))
      check(lib.ddg_adaln(ptr(ws['c']), mod.n_mod, self.cond_dim, ptr(self.ada_w),
                          ptr(self.ada_b), nb * 6 * h, bcode, h, 0b010010, 6,
                          ptr(ws['mod_blk']), mod_stride, s))
      if not self.classifier and self.final_adaln:
        check(lib.ddg_adaln(ptr(ws['c']), mod.n_mod, self.cond_dim,
                            ptr(self.ada_f_w), ptr(self.ada_f_b), 2 * h, hcode, h,
                            0b10, 2, ptr(ws['mod_fin']), 2 * h, s))
    elif mod.labels is not None:
      raise ValueError('Conditioning variable provided, but Model was not '
                       'initialized with condition embedding layer.')
    cos, sin = self._rope_tables(L)
    if live is not None and (self.classifier or save_for_backward is not None
                             or (live[0] is not None and L % 32)):
      raise ValueError('dynamic batch: denoiser forward only; with a sequence map seq_len % 32 == 0')
    seg_rows = M if (self.seg_seqs is None or save_for_backward is not None or live is not None) \
        else max(1, self.seg_seqs) * L
    for r0 in range(0, M, seg_rows):
      n = min(seg_rows, M - r0)
      if r0 % rpm == 0:
        mod0, rpm_arg = r0 // rpm, rpm
      elif r0 // rpm == (r0 + n - 1) // rpm:
        mod0, rpm_arg = r0 // rpm, _BIG
      else:
        raise ValueError('micro-batch straddles a conditioning group')
      if (r0 % period) + n > period and not (r0 == 0 and n == M):
        raise ValueError('micro-batch straddles the id period')
      x_final = self._run_segment(ws, ids, r0, n, L, period, mod0, rpm_arg, mod_stride,
                                  cos, sin, logits_out, bcode, hcode, s, stats_out,
                                  stats_mask_col, save_for_backward, live)
    if self.classifier:    # with save_for_backward (one segment) the rows live in their own buffer
      return x_final if save_for_backward is not None else ws['x'][:M]
    return logits_out

  def _run_segment(self, ws, ids, r0, n, L, period, mod0, rpm, mod_stride, cos,
                   sin, logits_out, bcode, hcode, s, stats_out=None,
                   stats_mask_col=-1, saved=None, live=None):
    h, nb = self.hidden, self.n_blocks
    seq_map, n_live = (ptr(live[0]), ptr(live[1])) if live is not None else (0, 0)
    x = ptr(ws['x']) + r0 * h * 4
    a16, qkv, h4 = ptr(ws['a16']), ptr(ws['qkv']), ptr(ws['h4'])
    # The LayerNorm + modulate that follows a residual add runs inside that GEMM's epilogue
    # (rows normalised while still on chip); only the first norm (embedding gather) is a launch
    # of its own.  The attention output then needs its own buffer: a16 is the norm's output.
    fuse = self.fuse_ln and h % 128 == 0 and h <= 1024 and rpm % 32 == 0
    # save_for_backward: every residual add writes a NEW buffer (out = residual + delta), so
    # each block's input rows (xs[2 i]) and mid rows (xs[2 i + 1]) stay available to the
    # backward without copying the residual stream
    xs = None
    if saved is not None:
      xs = torch.empty(2 * nb + 1, n, h, dtype=torch.float32, device=self.device)

    def xp(k):
      return xs[k].data_ptr() if xs is not None else x
    att = ptr(ws['att']) if fuse else a16
    ln_work = ptr(ws['ln_work'])
    modb = ptr(ws['mod_blk']) + mod0 * mod_stride * 2
    ids_p = ptr(ids) + (r0 % period) * 8
    ada = self.use_adaln

    def gemm(a, lda, w, ldw, out, ld_out, N, K, epi, dtype, bias=0, gate=0,
             gate_stride=0, stats=0, stats_ld=0, out_map=0, ln=None, residual=0, aux=0):
      g = _lib.GemmArgs(
        a=a, w=w, out=out, bias=bias, gate=gate, rope_cos=ptr(cos),
        rope_sin=ptr(sin), lda=lda, ldw=ldw, ld_out=ld_out,
        gate_stride=gate_stride, rows_per_mod=rpm, M=n, N=N, K=K, dtype=dtype,
        epilogue=epi, seq_len=L, head_dim=self.head_dim, rope_cols=2 * h,
        stats=stats, stats_ld=stats_ld, stats_mask_col=stats_mask_col,
        live_seqs=n_live, out_seq_map=out_map, residual=residual, aux=aux,
        ld_aux=N if aux else 0)
      if ln is not None:   # (weight, shift, 1 + scale, stride between conditioning rows, dtype)
        g.ln_weight, g.ln_shift, g.ln_scale, g.ln_mod_stride, g.ln_dtype = ln
        g.ln_out, g.ln_work = a16, ln_work
      check(lib.ddg_gemm(g, s))

    def ln(xp, ids_arg, emb, w, shift, scale, stride, code, out):
      if live is None:
        check(lib.ddg_ln_modulate(xp, n, h, ids_arg, period, emb, w, shift, scale,
                                  stride, rpm, code, out, s))
      else:
        check(lib.ddg_ln_modulate_live(xp, n, h, ids_arg, period, emb, w, shift, scale,
                                       stride, rpm, code, out, n_live, L,
                                       seq_map if ids_arg else 0, s))

    fin = self.use_adaln and not self.classifier and self.final_adaln
    mf = ptr(ws['mod_fin']) + mod0 * 2 * h * 2
    for i in range(nb):
      mb = modb + i * 6 * h * 2
      if i == 0 or not fuse:
        ln(xp(2 * i), ids_p if i == 0 else 0, ptr(self.embedding), ptr(self.ln1[i]),
           mb if ada else 0, mb + h * 2 if ada else 0, mod_stride, bcode, a16)
      qkv_i = qkv
      if saved is not None:   # the rotated QKV of this block is written straight into its own
        # buffer (kept for the backward)
        # buffer, and so is the MLP's pre-activation (the fc1 GEMM's side output): the backward
        # needs GELU'(pre) and would otherwise redo the norm and the fc1 GEMM
        keep = torch.empty(n, 3 * h, dtype=self.block_dtype, device=self.device)
        pre = torch.empty(n, 4 * h, dtype=self.block_dtype, device=self.device)
        saved.append([xs[2 * i], xs[2 * i + 1], keep, pre])
        qkv_i = ptr(keep)
      gemm(a16, h, ptr(self.w_qkv[i]), h, qkv_i, 3 * h, 3 * h, h,
           _lib.EPI_QKV_ROPE, bcode)
      if live is None:
        check(lib.ddg_attention(qkv_i, att, n // L, L, self.n_heads, self.head_dim,
                                1 if self.causal else 0, bcode, s))
      else:
        check(lib.ddg_attention_live(qkv_i, att, n // L, L, self.n_heads, self.head_dim,
                                     1 if self.causal else 0, bcode, n_live, s))
      ln2 = (ptr(self.ln2[i]), mb + 3 * h * 2 if ada else 0, mb + 4 * h * 2 if ada else 0,
             mod_stride, bcode)
      gemm(att, h, ptr(self.w_out[i]), h, xp(2 * i + 1), h, h, h, _lib.EPI_GATE_RESIDUAL,
           bcode, gate=mb + 2 * h * 2 if ada else 0, gate_stride=mod_stride,
           ln=ln2 if fuse else None, residual=xp(2 * i) if xs is not None else 0)
      if not fuse:
        ln(xp(2 * i + 1), 0, 0, *ln2[:4], bcode, a16)
      gemm(a16, h, ptr(self.w_fc1[i]), h, h4, 4 * h, 4 * h, h,
           _lib.EPI_BIAS_GELU, bcode, bias=ptr(self.b_fc1[i]),
           aux=ptr(saved[i][3]) if saved is not None else 0)
      # the norm after this block's last residual add: norm1 of the next block, or the
      # final layer's norm (bf16, models/dit.py:463); the classifier pools the raw rows
      nxt = None
      if fuse and i + 1 < nb:
        mn = mb + 6 * h * 2
        nxt = (ptr(self.ln1[i + 1]), mn if ada else 0, mn + h * 2 if ada else 0, mod_stride,
               bcode)
      elif fuse and not self.classifier:
        nxt = (ptr(self.ln_f), mf if fin else 0, mf + h * 2 if fin else 0, 2 * h, hcode)
      gemm(h4, 4 * h, ptr(self.w_fc2[i]), 4 * h, xp(2 * i + 2), h, h, 4 * h,
           _lib.EPI_GATE_RESIDUAL, bcode, bias=ptr(self.b_fc2[i]),
           gate=mb + 5 * h * 2 if ada else 0, gate_stride=mod_stride, ln=nxt,
           residual=xp(2 * i + 1) if xs is not None else 0)
    if self.classifier:
      return xs[2 * nb] if xs is not None else None
    if not fuse:
      ln(xp(2 * nb), 0, 0, ptr(self.ln_f), mf if fin else 0, mf + h * 2 if fin else 0, 2 * h,
         hcode, a16)
    st, st_ld = 0, 0
    if stats_out is not None:    # (chunks, rows_total, 2): this segment starts at row r0
      st, st_ld = ptr(stats_out) + r0 * 8, stats_out.shape[1]
    gemm(a16, h, ptr(self.w_vocab), h, ptr(logits_out) + r0 * self.vocab_ld * 2,
         self.vocab_ld, self.vocab_size, h, _lib.EPI_STORE, hcode,
         bias=ptr(self.b_vocab), stats=st, stats_ld=st_ld, out_map=seq_map)

  # ------------------------------------------------------------------ AR decode
  def alloc_kv_cache(self, n_seqs: int, L: int):
    """Per block, the (n_seqs * L, 3 * hidden) tensor the QKV GEMM fills one token
    row at a time (row b * L + t = token t of sequence b, rotary already applied)."""
    return [torch.zeros(n_seqs * L, 3 * self.hidden, dtype=self.block_dtype,
                        device=self.device) for _ in range(self.n_blocks)]

  def run_decode(self, ids_t: torch.Tensor, n_seqs: int, L: int, pos: int,
                 mod: ModSpec, kv_cache, logits_out: torch.Tensor,
                 ids_period: typing.Optional[int] = None):
    """One incremental step of the causal forward (AR sampling with a KV cache
    instead of the reference's prefix re-encode, diffusion.py:1132): the token
    at position `pos` of every sequence (`ids_t`, reused cyclically with period
    `ids_period`) -> logits (n_seqs, vocab_ld) of the next position.  Keys and
    values of positions 0..pos-1 come from `kv_cache`."""
    assert self._loaded and self.causal and not self.classifier
    n, h, nb = n_seqs, self.hidden, self.n_blocks
    period = n if ids_period is None else ids_period
    ws = self._workspace(max(n, 1), mod.n_mod)
    s = stream_ptr()
    bcode = _lib.DTYPE_CODE[self.block_dtype]
    hcode = _lib.DTYPE_CODE[self.head_dtype]
    mod_stride = nb * 6 * h
    rpm = mod.seqs_per_mod            # one token row per sequence
    ada = self.use_adaln
    if ada:
      check(lib.ddg_cond_embed(
        0, 0, 0, 0, ptr(mod.labels), mod.n_mod, self.cond_dim, _FREQ_DIM, 0, 0, 0, 0,
        ptr(self.label_table), ptr(ws['c']), s))
      check(lib.ddg_adaln(ptr(ws['c']), mod.n_mod, self.cond_dim, ptr(self.ada_w),
                          ptr(self.ada_b), nb * 6 * h, bcode, h, 0b010010, 6,
                          ptr(ws['mod_blk']), mod_stride, s))
      if self.final_adaln:
        check(lib.ddg_adaln(ptr(ws['c']), mod.n_mod, self.cond_dim,
                            ptr(self.ada_f_w), ptr(self.ada_f_b), 2 * h, hcode, h,
                            0b10, 2, ptr(ws['mod_fin']), 2 * h, s))
    cos, sin = self._rope_tables(L)
    x, a16, h4 = ptr(ws['x']), ptr(ws['a16']), ptr(ws['h4'])
    modb = ptr(ws['mod_blk'])

    def gemm(a, lda, w, ldw, out, ld_out, N, K, epi, dtype, bias=0, gate=0,
             gate_stride=0, rope_pos0=0):
      g = _lib.GemmArgs(
        a=a, w=w, out=out, bias=bias, gate=gate, rope_cos=ptr(cos),
        rope_sin=ptr(sin), lda=lda, ldw=ldw, ld_out=ld_out,
        gate_stride=gate_stride, rows_per_mod=rpm, M=n, N=N, K=K, dtype=dtype,
        epilogue=epi, seq_len=1, head_dim=self.head_dim, rope_cols=2 * h,
        stats=0, stats_ld=0, stats_mask_col=-1, rope_pos0=rope_pos0)
      check(lib.ddg_gemm(g, s))

    for i in range(nb):
      mb = modb + i * 6 * h * 2
      check(lib.ddg_ln_modulate(
        x, n, h, ptr(ids_t) if i == 0 else 0, period, ptr(self.embedding),
        ptr(self.ln1[i]), mb if ada else 0, mb + h * 2 if ada else 0,
        mod_stride, rpm, bcode, a16, s))
      # this token's q, k, v go straight into row b * L + pos of the cache
      gemm(a16, h, ptr(self.w_qkv[i]), h, ptr(kv_cache[i]) + pos * 3 * h * 2,
           L * 3 * h, 3 * h, h, _lib.EPI_QKV_ROPE, bcode, rope_pos0=pos)
      check(lib.ddg_attention_decode(ptr(kv_cache[i]), a16, n, L, pos, self.n_heads,
                                     self.head_dim, bcode, s))
      gemm(a16, h, ptr(self.w_out[i]), h, x, h, h, h, _lib.EPI_GATE_RESIDUAL,
           bcode, gate=mb + 2 * h * 2 if ada else 0, gate_stride=mod_stride)
      check(lib.ddg_ln_modulate(
        x, n, h, 0, period, 0, ptr(self.ln2[i]), mb + 3 * h * 2 if ada else 0,
        mb + 4 * h * 2 if ada else 0, mod_stride, rpm, bcode, a16, s))
      gemm(a16, h, ptr(self.w_fc1[i]), h, h4, 4 * h, 4 * h, h,
           _lib.EPI_BIAS_GELU, bcode, bias=ptr(self.b_fc1[i]))
      gemm(h4, 4 * h, ptr(self.w_fc2[i]), 4 * h, x, h, h, 4 * h,
           _lib.EPI_GATE_RESIDUAL, bcode, bias=ptr(self.b_fc2[i]),
           gate=mb + 5 * h * 2 if ada else 0, gate_stride=mod_stride)
    fin = ada and self.final_adaln
    mf = ptr(ws['mod_fin'])
    check(lib.ddg_ln_modulate(
      x, n, h, 0, period, 0, ptr(self.ln_f), mf if fin else 0,
      mf + h * 2 if fin else 0, 2 * h, rpm, hcode, a16, s))
    gemm(a16, h, ptr(self.w_vocab), h, ptr(logits_out), self.vocab_ld,
         self.vocab_size, h, _lib.EPI_STORE, hcode, bias=ptr(self.b_vocab))
    return logits_out

  # ------------------------------------------------------------------ API
  def _mod_spec(self, sigma, n: int, labels=None, seqs_per_mod: int = 1) -> ModSpec:
    """Conditioning rows of a forward over n sequences: one per `seqs_per_mod`
    consecutive sequences.  Every sequence of a sampling step shares t
    (diffusion.py:1304-1310 builds `t * ones(B)`), so an expanded scalar sigma
    (stride 0) without labels collapses to ONE row -- the conditioning MLP and the
    adaLN linears then run once instead of once per sequence (same values)."""
    mod = ModSpec(n_mod=n // seqs_per_mod, seqs_per_mod=seqs_per_mod)
    if not self.causal:
      sig = sigma.to(self.device, torch.float32).reshape(-1)
      if labels is None and (sig.numel() == 1 or sig.stride(0) == 0):
        mod.n_mod, mod.seqs_per_mod = 1, n
        sig = sig[:1]
      assert sig.numel() == mod.n_mod, (sig.numel(), mod.n_mod)
      mod.sigma = sig.contiguous()
      mod.sigma_row_stride = 1 if mod.n_mod > 1 else 0
    if labels is not None:
      mod.labels = labels.to(self.device, torch.int64).contiguous()
    return mod

  def forward(self, indices, sigma, cond=None, x_emb=None,
              return_hidden_states=False, stats_out=None, stats_mask_col=-1):
    """`DIT.forward(indices, sigma, cond)` (models/dit.py:426-468): logits
    (B, L, V) bf16 (a view of the padded buffer)."""
    if x_emb is not None or return_hidden_states:
      raise NotImplementedError('x_emb / return_hidden_states (NOS/PPLM paths) '
                                'are outside the sampling hot path')
    B, L = indices.shape
    ids = indices.to(self.device, torch.int64).contiguous()
    mod = self._mod_spec(sigma, B, labels=cond)
    logits = self.alloc_logits(B * L)
    self.run(ids, B, L, mod, logits, stats_out=stats_out,
             stats_mask_col=stats_mask_col)
    return logits.view(B, L, self.vocab_ld)[:, :, :self.vocab_size]

  __call__ = forward

  def classify(self, indices, sigma, round_bf16=True, seqs_per_mod: int = 1):
    """`Classifier.get_log_probs` for token ids (classifier.py:316-327 over
    models/dit.py:516-556): (n, L) ids -> (n, num_classes) fp32 log-probs.
    `sigma` holds one value per `seqs_per_mod` consecutive sequences (exact D-CBG:
    the L * V single-token edits of one sequence share its sigma)."""
    assert self.classifier
    n, L = indices.shape
    ids = indices.to(self.device, torch.int64).contiguous()
    mod = self._mod_spec(sigma, n, seqs_per_mod=seqs_per_mod)
    x = self.run(ids, n, L, mod)
    out = torch.empty(n, self.num_classes, dtype=torch.float32,
                      device=self.device)
    check(lib.ddg_classifier_head(ptr(x), n, L, self.hidden, ptr(self.cls_w),
                                  ptr(self.cls_b), self.num_classes,
                                  1 if round_bf16 else 0, ptr(out),
                                  stream_ptr()))
    return out

  def first_order_term(self, indices, sigma, target, round_bf16=True, finalize=True,
                       out=None, live=None):
    """First-order D-CBG classifier term (diffusion.py:1532-1548): one forward
    and the backward of log p(target | x) w.r.t. the one-hot input, then
    `grad - grad[xt] + log p(target | xt)` -> (n, L, V) fp32.  The forward is
    the one of `classify` (a one-hot input times the embedding matrix is the
    embedding gather, models/dit.py:518-522) with each block's residual-stream
    inputs and rotated QKV kept for the backward.

    `finalize=False` returns the raw gradient `grad` instead: the other two terms
    are the same for every vocabulary entry of a row, and a row-constant shift of
    the classifier term cancels in the softmax of `log q + gamma * term`
    (diffusion.py:1612-1617), i.e. it changes neither q_xs nor the ids -- the
    sampling loop skips that extra pass over (n, L, V).

    `out`: preallocated (n * L, pad4(V)) fp32 buffer for the term (the captured step loop);
    `live`: device int32 scalar -- when it reads 0 the term buffer keeps its previous contents
    (the whole-batch `sampling.use_cache` rule of the captured loop)."""
    assert self.classifier
    n, L = indices.shape
    h, nb, V, M = self.hidden, self.n_blocks, self.vocab_size, n * L
    dev, s = self.device, stream_ptr()
    bcode = _lib.DTYPE_CODE[self.block_dtype]
    ids = indices.to(dev, torch.int64).contiguous()
    mod = self._mod_spec(sigma, n)
    saved = []
    x = self.run(ids, n, L, mod, save_for_backward=saved)
    ws = self._ws
    rpm = mod.seqs_per_mod * L
    mod_stride = nb * 6 * h
    ada = self.use_adaln
    cos, sin = self._rope_tables(L)
    lp = torch.empty(n, self.num_classes, dtype=torch.float32, device=dev)
    g = torch.empty(M, h, dtype=torch.float32, device=dev)
    check(lib.ddg_classifier_head_backward(
      ptr(x), n, L, h, ptr(self.cls_w), ptr(self.cls_b), self.num_classes,
      int(target), 1 if round_bf16 else 0, ptr(lp), ptr(g), s))
    g16 = torch.empty(M, h, dtype=self.block_dtype, device=dev)
    g_a = torch.empty(M, h, dtype=self.block_dtype, device=dev)
    g_h = torch.empty(M, 4 * h, dtype=self.block_dtype, device=dev)
    g_qkv = torch.empty(M, 3 * h, dtype=self.block_dtype, device=dev)

    def gemm(a, lda, w, ldw, out, ld_out, N, K, epi=_lib.EPI_STORE, bias=0,
             accumulate=0, live_seqs=0, aux=0):
      ga = _lib.GemmArgs(
        a=a, w=w, out=out, bias=bias, gate=0, rope_cos=ptr(cos), rope_sin=ptr(sin),
        lda=lda, ldw=ldw, ld_out=ld_out, gate_stride=0, rows_per_mod=rpm, M=M, N=N,
        K=K, dtype=bcode, epilogue=epi, seq_len=L, head_dim=self.head_dim,
        rope_cols=0, stats=0, stats_ld=0, stats_mask_col=-1, accumulate=accumulate,
        live_seqs=live_seqs, aux=aux, ld_aux=N if aux else 0)
      check(lib.ddg_gemm(ga, s))

    for i in reversed(range(nb)):
      mb = ptr(ws['mod_blk']) + i * 6 * h * 2
      x_in, x_mid, qkv_i, pre_i = saved[i]
      # ---- MLP branch: x_out = x_mid + gate_mlp * fc2(gelu(fc1(mod(LN(x_mid))))).  The fc2
      # input-gradient GEMM multiplies by GELU'(pre) in its epilogue (pre = the forward's fc1
      # side output): no recomputed norm / fc1 GEMM and no elementwise pass over (M, 4h)
      check(lib.ddg_gate_mul(ptr(g), mb + 5 * h * 2 if ada else 0, mod_stride, rpm,
                             M, h, bcode, ptr(g16), s))
      gemm(ptr(g16), h, ptr(self.w_fc2_t[i]), h, ptr(g_h), 4 * h, 4 * h, h,
           epi=_lib.EPI_GELU_BWD, aux=ptr(pre_i))
      gemm(ptr(g_h), 4 * h, ptr(self.w_fc1_t[i]), 4 * h, ptr(g_a), h, h, 4 * h)
      check(lib.ddg_ln_modulate_backward(
        ptr(x_mid), ptr(g_a), M, h, ptr(self.ln2[i]), mb + 4 * h * 2 if ada else 0,
        mod_stride, rpm, bcode, ptr(g), s))
      # ---- attention branch: x_mid = x_in + gate_msa * out(attn(rope(qkv(mod(LN(x_in))))))
      check(lib.ddg_gate_mul(ptr(g), mb + 2 * h * 2 if ada else 0, mod_stride, rpm,
                             M, h, bcode, ptr(g16), s))
      gemm(ptr(g16), h, ptr(self.w_out_t[i]), h, ptr(g_a), h, h, h)
      check(lib.ddg_attention_backward(
        ptr(qkv_i), ptr(g_a), ptr(cos), ptr(sin), n, L, self.n_heads,
        self.head_dim, 1 if self.causal else 0, bcode, ptr(g_qkv), s))
      gemm(ptr(g_qkv), 3 * h, ptr(self.w_qkv_t[i]), 3 * h, ptr(g_a), h, h, 3 * h)
      check(lib.ddg_ln_modulate_backward(
        ptr(x_in), ptr(g_a), M, h, ptr(self.ln1[i]), mb + 1 * h * 2 if ada else 0,
        mod_stride, rpm, bcode, ptr(g), s))
    # ---- d / d one-hot = G @ E^T in (almost) fp32: hi*hi + lo*hi + hi*lo, accumulated in ONE
    # pass over the (M, V) fp32 output by concatenating the three products along K
    g_lo = g16
    g_hi = g_a
    check(lib.ddg_split16(ptr(g), ptr(g_hi), ptr(g_lo), M * h, bcode, s))
    g_cat = torch.cat([g_hi, g_lo, g_hi], dim=1)
    ldv = _pad_to(V, 8)      # whole 8-column vectors per row: the single-pass sampler reads the term 2 x 16 B at a time
    term = out if out is not None else torch.empty(M, ldv, dtype=torch.float32, device=dev)
    assert term.shape == (M, ldv) and term.is_contiguous()
    gemm(ptr(g_cat), 3 * h, ptr(self.emb_cat), 3 * h, ptr(term), ldv, V, 3 * h,
         epi=_lib.EPI_ACCUM_F32, live_seqs=ptr(live))
    if finalize:
      check(lib.ddg_cbg_first_order(
        ptr(term), ldv, ptr(ids), ptr(lp), self.num_classes, int(target), L, V, M,
        ptr(term), ldv, s))
    return term.view(n, L, ldv)[:, :, :V], lp
