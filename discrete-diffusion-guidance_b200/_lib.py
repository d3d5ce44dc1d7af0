"""ctypes binding of libddg_b200.so (the C ABI declared in include/ddg_b200.h).

There is no fallback: if the library is missing this module raises at import
time, and on a machine without an sm_100 GPU every compute call raises."""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# DDG_B200_LIB: developer knob for same-box A/B runs of two builds of the library
LIB_PATH = os.environ.get('DDG_B200_LIB') or os.path.join(_HERE, 'libddg_b200.so')

ABI_VERSION = 6   # DDG_ABI_VERSION of include/ddg_b200.h these bindings were written for
OK, ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED = 0, -1, -2, -3
F32, F16, BF16 = 0, 1, 2
ABSORBING, UNIFORM, AR = 0, 1, 2
GUIDE_NONE, GUIDE_CFG, GUIDE_CBG = 0, 1, 2
SELECT_RACE, SELECT_ARGMAX, SELECT_GUMBEL = 0, 1, 2
EPI_STORE, EPI_QKV_ROPE, EPI_BIAS_GELU, EPI_GATE_RESIDUAL, EPI_ACCUM_F32, EPI_GELU_BWD = 0, 1, 2, 3, 4, 5

DTYPE_CODE = {torch.float32: F32, torch.float16: F16, torch.bfloat16: BF16}

EXPORTS = (
  'ddg_abi_version', 'ddg_last_error', 'ddg_device_check', 'ddg_sm_count',
  'ddg_sample_step', 'ddg_low_confidence_select', 'ddg_log_probs', 'ddg_sample_prior', 'ddg_step_advance',
  'ddg_cond_embed', 'ddg_adaln', 'ddg_ln_modulate', 'ddg_ln_modulate_live', 'ddg_seq_compact', 'ddg_gemm', 'ddg_gemm_ln_work_bytes',
  'ddg_attention', 'ddg_attention_live', 'ddg_attention_decode', 'ddg_cast', 'ddg_classifier_head', 'ddg_cbg_candidates',
  'ddg_classifier_head_backward', 'ddg_gate_mul', 'ddg_gelu_backward',
  'ddg_ln_modulate_backward', 'ddg_attention_backward', 'ddg_split16',
  'ddg_cbg_first_order')


class SampleStepArgs(C.Structure):
  _fields_ = [
    ('logits_a', C.c_void_p), ('logits_b', C.c_void_p), ('clf', C.c_void_p),
    ('noise', C.c_void_p), ('xt', C.c_void_p), ('mc_t', C.c_void_p),
    ('mc_s', C.c_void_p), ('step_ptr', C.c_void_p), ('seed_ptr', C.c_void_p),
    ('xs', C.c_void_p),
    ('q_out', C.c_void_p), ('rows', C.c_int64), ('ld_a', C.c_int64),
    ('ld_b', C.c_int64), ('ld_clf', C.c_int64), ('ld_noise', C.c_int64),
    ('ld_q', C.c_int64), ('seq_len', C.c_int32), ('vocab', C.c_int32),
    ('mc_seq_stride', C.c_int32), ('mc_step_stride', C.c_int32),
    ('logits_dtype', C.c_int32), ('diffusion', C.c_int32),
    ('guidance', C.c_int32), ('subs', C.c_int32), ('mask_logit', C.c_int32),
    ('mask_index', C.c_int32), ('round_bf16', C.c_int32),
    ('select', C.c_int32), ('fast_math', C.c_int32), ('gamma', C.c_float),
    ('seed', C.c_uint64), ('step', C.c_int32), ('stats_chunks', C.c_int32),
    ('stats_a', C.c_void_p), ('stats_b', C.c_void_p), ('stats_ld', C.c_int64),
    ('race_group', C.c_int32), ('compute_f64', C.c_int32),
    ('clf_round_bf16', C.c_int32), ('reserved0', C.c_int32),   # ABI v3
    ('seq_changed', C.c_void_p)]                               # ABI v4


class GemmArgs(C.Structure):
  _fields_ = [
    ('a', C.c_void_p), ('w', C.c_void_p), ('out', C.c_void_p),
    ('bias', C.c_void_p), ('gate', C.c_void_p), ('rope_cos', C.c_void_p),
    ('rope_sin', C.c_void_p), ('lda', C.c_int64), ('ldw', C.c_int64),
    ('ld_out', C.c_int64), ('gate_stride', C.c_int64),
    ('rows_per_mod', C.c_int64), ('M', C.c_int32), ('N', C.c_int32),
    ('K', C.c_int32), ('dtype', C.c_int32), ('epilogue', C.c_int32),
    ('seq_len', C.c_int32), ('head_dim', C.c_int32), ('rope_cols', C.c_int32),
    ('stats', C.c_void_p), ('stats_ld', C.c_int64),
    ('stats_mask_col', C.c_int32), ('accumulate', C.c_int32),
    ('rope_pos0', C.c_int32), ('reserved', C.c_int32),
    ('live_seqs', C.c_void_p), ('out_seq_map', C.c_void_p),      # ABI v4
    ('ln_weight', C.c_void_p), ('ln_shift', C.c_void_p), ('ln_scale', C.c_void_p),   # ABI v5
    ('ln_out', C.c_void_p), ('ln_work', C.c_void_p), ('ln_mod_stride', C.c_int64),
    ('ln_dtype', C.c_int32), ('reserved1', C.c_int32), ('residual', C.c_void_p),
    ('aux', C.c_void_p), ('ld_aux', C.c_int64)]                                      # ABI v6


def _load():
  if not os.path.exists(LIB_PATH):
    raise ImportError(
      f'{LIB_PATH} not found: build it with `python -c "import '
      '__graft_entry__ as g; g.build()"` (nvcc, sm_100a). ddg_b200 has no '
      'CPU or PyTorch fallback.')
  lib = C.CDLL(LIB_PATH)
  for name in EXPORTS:
    if not hasattr(lib, name):
      raise ImportError(f'{LIB_PATH} does not export {name}')
  lib.ddg_last_error.restype = C.c_char_p
  if lib.ddg_abi_version() != ABI_VERSION:
    raise ImportError(f'{LIB_PATH} has ABI version {lib.ddg_abi_version()}, these bindings need '
                      f'{ABI_VERSION}: rebuild it (python -c "import __graft_entry__ as g; g.build()")')
  i32, i64, vp, u32, u64 = C.c_int32, C.c_int64, C.c_void_p, C.c_uint32, C.c_uint64
  lib.ddg_sample_step.argtypes = [C.POINTER(SampleStepArgs), vp]
  lib.ddg_low_confidence_select.argtypes = [vp, i64, i64, i32, C.c_float, vp, i64, i32, u64, vp,
                                            i32, vp, vp, i32, vp, vp]
  lib.ddg_log_probs.argtypes = [vp, i64, i32, vp, i64, i32, i32, i32, i32,
                                i32, vp, i64, vp]
  lib.ddg_sample_prior.argtypes = [vp, i64, i32, i32, i32, u64, vp]
  lib.ddg_step_advance.argtypes = [vp, vp]
  lib.ddg_cond_embed.argtypes = [vp, i32, vp, i32, vp, i32, i32, i32, vp, vp,
                                 vp, vp, vp, vp, vp]
  lib.ddg_adaln.argtypes = [vp, i32, i32, vp, vp, i32, i32, i32, u32, i32, vp,
                            i64, vp]
  lib.ddg_ln_modulate.argtypes = [vp, i64, i32, vp, i64, vp, vp, vp, vp, i64,
                                  i64, i32, vp, vp]
  lib.ddg_ln_modulate_live.argtypes = [vp, i64, i32, vp, i64, vp, vp, vp, vp, i64,
                                       i64, i32, vp, vp, i32, vp, vp]
  lib.ddg_seq_compact.argtypes = [vp, i32, vp, vp, i32, vp]
  lib.ddg_gemm.argtypes = [C.POINTER(GemmArgs), vp]
  lib.ddg_gemm_ln_work_bytes.argtypes = [i32, i32]
  lib.ddg_gemm_ln_work_bytes.restype = i64
  lib.ddg_attention.argtypes = [vp, vp, i32, i32, i32, i32, i32, i32, vp]
  lib.ddg_attention_live.argtypes = [vp, vp, i32, i32, i32, i32, i32, i32, vp, vp]
  lib.ddg_attention_decode.argtypes = [vp, vp, i32, i32, i32, i32, i32, i32, vp]
  lib.ddg_cast.argtypes = [vp, vp, i64, i32, vp]
  lib.ddg_classifier_head.argtypes = [vp, i32, i32, i32, vp, vp, i32, i32, vp,
                                      vp]
  lib.ddg_cbg_candidates.argtypes = [vp, i32, i32, i64, i64, vp, vp]
  lib.ddg_classifier_head_backward.argtypes = [vp, i32, i32, i32, vp, vp, i32, i32, i32, vp, vp, vp]
  lib.ddg_gate_mul.argtypes = [vp, vp, i64, i64, i64, i32, i32, vp, vp]
  lib.ddg_gelu_backward.argtypes = [vp, vp, vp, i64, i32, vp]
  lib.ddg_ln_modulate_backward.argtypes = [vp, vp, i64, i32, vp, vp, i64, i64, i32, vp, vp]
  lib.ddg_attention_backward.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp]
  lib.ddg_split16.argtypes = [vp, vp, vp, i64, i32, vp]
  lib.ddg_cbg_first_order.argtypes = [vp, i64, vp, vp, i32, i32, i32, i32, i64, vp, i64, vp]
  return lib


lib = _load()


def last_error() -> str:
  return (lib.ddg_last_error() or b'').decode()


def check(rc: int):
  """Maps a DDG status to the exception the reference raises for the same
  mistake (NotImplementedError for unknown diffusion/guidance, ValueError for a
  condition without cond_map)."""
  if rc == OK:
    return
  msg = last_error()
  if 'not implemented' in msg:
    raise NotImplementedError(msg)
  if rc == ERR_INVALID:
    raise ValueError(msg)
  if rc == ERR_UNSUPPORTED:
    raise RuntimeError(f'ddg_b200 unsupported: {msg}')
  raise RuntimeError(f'ddg_b200 CUDA failure: {msg}')


def stream_ptr() -> int:
  return torch.cuda.current_stream().cuda_stream


def ptr(t) -> int:
  return 0 if t is None else t.data_ptr()


def require_device():
  check(lib.ddg_device_check())
