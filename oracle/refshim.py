"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference sampler.

Imports /root/reference (read-only, exists only in the build container) or, on
the GPU box, its staged byte-for-byte copy `baseline/_ref/` (git-ignored, made by
oracle/stage_reference.py at build() time) with import-time stubs for the packages this image lacks (lightning, hydra,
omegaconf, torchmetrics, mamba_ssm, causal_conv1d, timm) and a pure-torch
stand-in for flash_attn's Triton-only rotary (reference models/dit.py:116-121).
Nothing of the reference is copied: the modules are imported where they lie.

Used by tests/golden/make_golden.py (golden-vector generator), by
tests/test_oracle_pins.py / tests/test_gpu_reference.py (skipped when no copy of
the reference is reachable) and by `bench.py --impl reference` (the CPU arm:
the reference's own step on the host cores).  Never imported by the product
package, bench.py's GPU arm or smoke().
"""
from __future__ import annotations

import importlib.machinery
import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGED_ROOT = os.path.join(_REPO, 'baseline', '_ref')   # oracle/stage_reference.py


def _find_root() -> str:
  """/root/reference in the build container; on the GPU box the byte-for-byte
  copy `__graft_entry__.build()` staged under baseline/_ref (git-ignored)."""
  env = os.environ.get('DDG_REFERENCE_ROOT')
  if env:
    return env
  for cand in ('/root/reference', STAGED_ROOT):
    if os.path.isfile(os.path.join(cand, 'diffusion.py')):
      return cand
  return '/root/reference'


REFERENCE_ROOT = _find_root()


def reference_available() -> bool:
  return os.path.isfile(os.path.join(REFERENCE_ROOT, 'diffusion.py'))


class AttrDict(dict):
  """Attribute-style dict standing in for omegaconf.DictConfig."""

  def __getattr__(self, k):
    try:
      v = self[k]
    except KeyError as e:
      raise AttributeError(k) from e
    return v

  def __setattr__(self, k, v):
    self[k] = v


def to_attrdict(d):
  if isinstance(d, dict):
    return AttrDict({k: to_attrdict(v) for k, v in d.items()})
  if isinstance(d, (list, tuple)):
    return type(d)(to_attrdict(v) for v in d)
  return d


def _mod(name, **attrs):
  m = types.ModuleType(name)
  m.__spec__ = importlib.machinery.ModuleSpec(name, None)
  m.__path__ = []
  for k, v in attrs.items():
    setattr(m, k, v)
  sys.modules[name] = m
  return m


def _install_stubs():
  import torch
  import torch.nn as nn

  if 'lightning' not in sys.modules:
    class LightningModule(nn.Module):
      def __init__(self, *a, **k):
        super().__init__()

      def save_hyperparameters(self, *a, **k):
        pass

      def log(self, *a, **k):
        pass

      def log_dict(self, *a, **k):
        pass

      @property
      def dtype(self):
        return torch.float32

      @property
      def device(self):
        for p in self.parameters():
          return p.device
        return torch.device('cpu')

      @classmethod
      def load_from_checkpoint(cls, *a, **k):
        raise RuntimeError('stub: patch load_from_checkpoint')

    rz = _mod('lightning.pytorch.utilities',
              rank_zero_only=lambda f: f)
    lp = _mod('lightning.pytorch', utilities=rz)
    _mod('lightning', LightningModule=LightningModule, pytorch=lp,
         seed_everything=lambda s: torch.manual_seed(s))
  if 'hydra' not in sys.modules:
    hu = _mod('hydra.utils', instantiate=lambda *a, **k: None)
    _mod('hydra', utils=hu)
  if 'omegaconf' not in sys.modules:
    class OmegaConf:
      @staticmethod
      def create(d):
        return to_attrdict(d)

      @staticmethod
      def register_new_resolver(*a, **k):
        pass
    _mod('omegaconf', OmegaConf=OmegaConf, DictConfig=AttrDict)
  if 'torchmetrics' not in sys.modules:
    class Metric(nn.Module):
      def __init__(self, *a, **k):
        super().__init__()

      def add_state(self, name, default, dist_reduce_fx=None):
        setattr(self, name, default)

    class MeanMetric(Metric):
      """torchmetrics.aggregation.MeanMetric as the reference uses it (diffusion.py:46-67: NLL /
      BPD / Perplexity read `mean_value` and `weight`): weighted running sum and weight sum."""
      def __init__(self, *a, **k):
        super().__init__()
        self.mean_value = torch.zeros(())
        self.weight = torch.zeros(())

      def update(self, value, weight=1.0):
        value = torch.as_tensor(value, dtype=torch.float32)
        weight = torch.broadcast_to(torch.as_tensor(weight, dtype=torch.float32, device=value.device),
                                    value.shape)
        self.mean_value = self.mean_value.to(value.device) + (value * weight).sum()
        self.weight = self.weight.to(value.device) + weight.sum()

      def compute(self):
        return self.mean_value / self.weight

      def reset(self):
        self.mean_value = torch.zeros(())
        self.weight = torch.zeros(())

    class MetricCollection(nn.Module):
      def __init__(self, *a, **k):
        super().__init__()

      def set_dtype(self, *_):
        pass

      def clone(self, prefix=''):
        return MetricCollection()

      def add_metrics(self, *_):
        pass

      def reset(self):
        pass
    agg = _mod('torchmetrics.aggregation', MeanMetric=MeanMetric)
    _mod('torchmetrics', Metric=Metric, MetricCollection=MetricCollection,
         aggregation=agg)
  if 'mamba_ssm' not in sys.modules:
    class InferenceParams:
      def __init__(self, *a, **k):
        pass
    _mod('mamba_ssm')
    _mod('mamba_ssm.ops')
    _mod('mamba_ssm.ops.selective_scan_interface',
         mamba_inner_fn=None, selective_scan_fn=None)
    _mod('mamba_ssm.ops.triton')
    _mod('mamba_ssm.ops.triton.layernorm', RMSNorm=None,
         layer_norm_fn=None, rms_norm_fn=None)
    _mod('mamba_ssm.ops.triton.selective_state_update',
         selective_state_update=None)
    _mod('mamba_ssm.utils')
    _mod('mamba_ssm.utils.generation', InferenceParams=InferenceParams)
  if 'causal_conv1d' not in sys.modules:
    _mod('causal_conv1d', causal_conv1d_fn=None, causal_conv1d_update=None)
  if 'timm' not in sys.modules:
    class CosineLRScheduler:
      def __init__(self, *a, **k):
        pass
    sch = _mod('timm.scheduler', CosineLRScheduler=CosineLRScheduler)
    _mod('timm', scheduler=sch)


_REF = None


def load_reference():
  """Returns a namespace with the reference's `diffusion`, `classifier`,
  `dataloader`, `noise_schedule`, `models` modules (imported unmodified)."""
  global _REF
  if _REF is not None:
    return _REF
  if not reference_available():
    raise RuntimeError(f'reference not found at {REFERENCE_ROOT}')
  import torch  # noqa: F401
  import transformers  # noqa: F401
  import datasets  # noqa: F401
  _install_stubs()
  if REFERENCE_ROOT not in sys.path:
    sys.path.insert(0, REFERENCE_ROOT)
  import flash_attn.layers.rotary as far

  if not getattr(far, '_ddg_cpu_shim', False):
    _orig = far.apply_rotary_emb_qkv_

    def _rotary_qkv_(qkv, cos, sin, *a, **k):
      # reference models/dit.py:116-121 calls the Triton kernel; on CPU use
      # the same package's pure-torch form on q and k (v untouched).  Like the
      # Triton kernel (flash_attn/ops/triton/rotary.py:73-88) the math runs in
      # fp32 on the 16-bit q, k, cos, sin values and is rounded once.
      if qkv.is_cuda:
        return _orig(qkv, cos, sin, *a, **k)
      for i in (0, 1):
        qkv[:, :, i] = far.apply_rotary_emb_torch(
          qkv[:, :, i].float(), cos.float(), sin.float()).to(qkv.dtype)
      return qkv
    far.apply_rotary_emb_qkv_ = _rotary_qkv_
    far._ddg_cpu_shim = True
  import classifier
  import dataloader
  import diffusion
  import models
  import noise_schedule
  _REF = types.SimpleNamespace(
    diffusion=diffusion, classifier=classifier, dataloader=dataloader,
    noise_schedule=noise_schedule, models=models)
  return _REF


class _CpuAutocast:
  """Stand-in for `torch.cuda.amp.autocast` that applies the SAME dtype policy
  with CPU autocast, so that the unmodified reference rounds on the CPU where it
  rounds on a GPU: fp16 / bf16 requests become `torch.autocast('cpu', dtype)`.
  A float32 request (diffusion.py:353, classifier.py:312,325) is NOT a no-op on
  CUDA: torch keeps that autocast enabled (torch/amp/autocast_mode.py only rejects
  unsupported dtypes for non-cuda devices), so ops on autocast's fp32 list --
  on this path `log_softmax` in `Classifier.get_log_probs` -- get their 16-bit
  inputs cast up to float32.  Confirmed by the reference's own run on the B200
  (classifier_log_prob is float32 there).  CPU autocast has no float32 mode, so it
  is emulated: nested lower-precision autocast is switched off and
  `F.log_softmax` upcasts while the block is active."""
  _fp32_depth = 0

  def __init__(self, enabled=True, dtype=None, cache_enabled=True):
    import torch
    dtype = torch.float16 if dtype is None else dtype
    self._fp32 = enabled and dtype not in (torch.float16, torch.bfloat16)
    if self._fp32:
      enabled, dtype = False, torch.bfloat16
    self._ctx = torch.autocast('cpu', dtype=dtype, enabled=enabled,
                               cache_enabled=cache_enabled)

  def __enter__(self):
    import torch.nn.functional as F
    if self._fp32:
      cls = _CpuAutocast
      if cls._fp32_depth == 0:
        cls._orig_log_softmax = F.log_softmax

        def log_softmax(x, *a, **k):
          # inner fp16 / bf16 blocks have exited by the time the reference calls this
          return cls._orig_log_softmax(x.float(), *a, **k)
        F.log_softmax = log_softmax
      cls._fp32_depth += 1
    return self._ctx.__enter__()

  def __exit__(self, *exc):
    import torch.nn.functional as F
    if self._fp32:
      cls = _CpuAutocast
      cls._fp32_depth -= 1
      if cls._fp32_depth == 0:
        F.log_softmax = cls._orig_log_softmax
    return self._ctx.__exit__(*exc)


class cuda_autocast_on_cpu:
  """Context manager: while active, the reference's `torch.cuda.amp.autocast`
  blocks take effect on CPU tensors (see _CpuAutocast).  This is how the GPU
  rounding points of the path (fp16 blocks, bf16 final layer, bf16 log-softmax /
  exp; SURVEY.md 8a') are pinned without a GPU run of the reference."""

  def __enter__(self):
    import torch
    self._saved = torch.cuda.amp.autocast
    torch.cuda.amp.autocast = _CpuAutocast
    return self

  def __exit__(self, *exc):
    import torch
    torch.cuda.amp.autocast = self._saved
    return False


class SurrogateTokenizer:
  """Offline stand-in for the HF tokenizers the reference downloads
  (bert-base-uncased: vocab 30522, [MASK]=103; reference diffusion.py:80-95
  only reads these attributes on the sampling path)."""

  def __init__(self, vocab_size, mask_token_id, bos_token_id=101,
               eos_token_id=102, pad_token_id=0):
    self.vocab_size = vocab_size
    self.mask_token = '[MASK]'
    self.mask_token_id = mask_token_id
    self.bos_token_id = bos_token_id
    self.eos_token_id = eos_token_id
    self.pad_token_id = pad_token_id
