"""Config surface of the sampling path.

The reference is driven by Hydra/OmegaConf (`configs/config.yaml` plus groups,
`main.py:399-435`); neither package is in this image, so this module parses
the same override syntax itself: group selection (`guidance=cfg`), dotted
assignments (`guidance.gamma=2`), additions (`+data.num_classes=2`) and `null`.
Only keys the sampling path reads are carried; the key
names and default values are the reference's.  A directory of the reference's
own YAML files can be given instead of the built-in group tables.
"""
from __future__ import annotations

import copy
import os
import typing


class Config(dict):
  """Attribute-style nested dict (stand-in for omegaconf.DictConfig)."""

  def __getattr__(self, k):
    try:
      return self[k]
    except KeyError as e:
      raise AttributeError(k) from e

  def __setattr__(self, k, v):
    self[k] = v

  def __deepcopy__(self, memo):
    return Config({k: copy.deepcopy(v, memo) for k, v in self.items()})


def to_config(d):
  if isinstance(d, dict):
    return Config({k: to_config(v) for k, v in d.items()})
  if isinstance(d, (list, tuple)):
    return [to_config(v) for v in d]
  return d


_MODEL = {
  'tiny': dict(name='tiny', type='ddit', hidden_size=512, cond_dim=128,
               length=1024, n_blocks=8, n_heads=8, scale_by_sigma=True,
               dropout=0.1, tie_word_embeddings=False),
  'small': dict(name='small', type='ddit', hidden_size=768, cond_dim=128,
                length=1024, n_blocks=12, n_heads=12, scale_by_sigma=True,
                dropout=0.1, tie_word_embeddings=False),
  'medium': dict(name='medium', type='ddit', hidden_size=1024, cond_dim=128,
                 length=1024, n_blocks=24, n_heads=16, scale_by_sigma=True,
                 dropout=0.1, tie_word_embeddings=False),
}
_CLASSIFIER_MODEL = {
  'tiny-classifier': dict(name='tiny', type='ddit', hidden_size=512,
                          cond_dim=128, n_blocks=8, n_heads=8,
                          scale_by_sigma=True, dropout=0.1, pooling='mean'),
  'small-classifier': dict(name='small', type='ddit', hidden_size=768,
                           cond_dim=128, n_blocks=12, n_heads=12,
                           scale_by_sigma=True, dropout=0.1, pooling='mean'),
}
_GUIDANCE = {
  'cfg': dict(method='cfg', condition=0, gamma=1.0),
  'cbg': dict(method='cbg', condition=0, classifier_checkpoint_path='',
              gamma=1.0, use_approx=False),
}
_NOISE = {
  'loglinear': dict(type='loglinear', sigma_min=1e-4, sigma_max=20),
  'cosine': dict(type='cosine'),
  'cosinesqr': dict(type='cosinesqr'),
  'ar': dict(type='ar', scale=6.0),
}
_DATA = {
  'lm1b': dict(train='lm1b', valid='lm1b',
               tokenizer_name_or_path='bert-base-uncased'),
  'text8': dict(train='text8', valid='text8', tokenizer_name_or_path='text8'),
  'qm9': dict(train='qm9', valid='qm9',
              tokenizer_name_or_path='yairschiff/qm9-tokenizer',
              num_classes=2),
  'amazon_polarity': dict(train='amazon_polarity', valid='amazon_polarity',
                          tokenizer_name_or_path='bert-base-uncased',
                          num_classes=2),
}
_GROUPS = {'model': _MODEL, 'classifier_model': _CLASSIFIER_MODEL,
           'guidance': _GUIDANCE, 'noise': _NOISE, 'data': _DATA}

# configs/config.yaml defaults, sampling-path keys only
_ROOT = dict(
  mode='sample', diffusion='absorbing_state', backbone='dit',
  classifier_backbone=None, parameterization='subs', time_conditioning=False,
  subs_masking=False, T=0, is_vision=False, seed=1,
  sampling=dict(use_cache=True, steps=128, batch_size=1,
                num_sample_batches=2, use_float64=False),
  training=dict(guidance=dict(cond_dropout=0.0), sampling_eps=1e-3),
  eval=dict(checkpoint_path='', disable_ema=False,
            low_confidence_sampling=False, low_confidence_threshold=0.3,
            argmax_sampling=False),
)
_DEFAULT_GROUPS = dict(data='lm1b', model='small', noise='loglinear',
                       classifier_model=None, guidance=None)


def _parse_value(s: str):
  low = s.strip()
  if low in ('null', 'None', '~'):
    return None
  if low in ('true', 'True'):
    return True
  if low in ('false', 'False'):
    return False
  for cast in (int, float):
    try:
      return cast(low)
    except ValueError:
      pass
  if len(low) >= 2 and low[0] == low[-1] and low[0] in '"\'':
    return low[1:-1]
  return low


def _load_group(group: str, name: str, config_dir: typing.Optional[str]):
  if config_dir is not None:
    import yaml
    path = os.path.join(config_dir, group, name + '.yaml')
    with open(path) as f:
      return {k: v for k, v in yaml.safe_load(f).items()
              if not (isinstance(v, str) and '${' in v)}
  try:
    return copy.deepcopy(_GROUPS[group][name])
  except KeyError:
    raise KeyError(f"Could not find '{group}/{name}'") from None


def compose(overrides: typing.Sequence[str] = (),
            config_dir: typing.Optional[str] = None) -> Config:
  """Builds the config the way `python main.py <overrides>` would."""
  cfg = copy.deepcopy(_ROOT)
  groups = dict(_DEFAULT_GROUPS)
  dotted = []
  for ov in overrides:
    key, _, val = ov.partition('=')
    key = key.lstrip('+')
    if '.' not in key and key in _GROUPS:
      groups[key] = _parse_value(val)
    else:
      dotted.append((key, _parse_value(val)))
  for g, name in groups.items():
    cfg[g] = None if name is None else _load_group(g, str(name), config_dir)
  for key, val in dotted:
    node = cfg
    parts = key.split('.')
    for p in parts[:-1]:
      if node.get(p) is None:
        node[p] = {}
      node = node[p]
    node[parts[-1]] = val
  if cfg.get('classifier_model') is not None:
    cm = cfg['classifier_model']
    cm.setdefault('length', cfg['model']['length'])
    if cm.get('num_classes') is None and cfg.get('data'):
      cm['num_classes'] = cfg['data'].get('num_classes')
  return to_config(cfg)


def validate(cfg: Config):
  """Diffusion._validate_configuration (reference diffusion.py:184-192)."""
  if cfg.diffusion != 'absorbing_state':
    assert cfg.parameterization not in {'ar', 'subs'}
  if cfg.T > 0:
    assert cfg.parameterization in {'d3pm', 'subs'}
  if cfg.subs_masking:
    assert cfg.parameterization == 'd3pm'
