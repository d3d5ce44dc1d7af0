"""Callers' post-processing of sampled ids: what the reference's
`_gen_ppl_eval` does with the output of `Diffusion.sample()` (main.py:253-305)
-- `tokenizer.batch_decode`, stripping of [PAD] / [MASK], BOS prefixing, and the
unigram entropy of the re-encoded samples.  Host-side code: the ids are (B, L)
int64, three orders of magnitude less data than one step's logits.

Generative perplexity (`eval_utils.compute_generative_ppl`, eval_utils.py:28-90)
scores the decoded text with a pretrained GPT-2; its weights need a network, so
`generate_and_score` takes the scorer as an optional callable instead.
"""
from __future__ import annotations

import typing

import torch


def decode_samples(tokenizer, ids: torch.Tensor) -> typing.List[str]:
  """main.py:263-277: decode, drop '[PAD]' / '[MASK]', strip, and put the BOS
  (or CLS) token in front of every sample that does not start with it."""
  samples = tokenizer.batch_decode(ids.cpu() if isinstance(ids, torch.Tensor) else ids)
  bos = tokenizer.bos_token if tokenizer.bos_token is not None else tokenizer.cls_token
  samples = [s.replace('[PAD]', '').replace('[MASK]', '').strip() for s in samples]
  return [s if s.startswith(bos) else f'{bos} {s}' for s in samples]


def _encode(tokenizer, samples, max_length):
  enc = getattr(tokenizer, 'batch_encode_plus', None) or tokenizer
  return enc(samples, return_tensors='pt', add_special_tokens=False, max_length=max_length,
             padding='max_length', truncation=True)['input_ids']


def unigram_entropy(tokenizer, samples: typing.Sequence[str], max_length: int) -> float:
  """main.py:285-296: entropy (nats) of the unigram distribution of the samples
  re-encoded to `max_length` tokens (padding included, like the reference)."""
  tokens = torch.as_tensor(_encode(tokenizer, list(samples), max_length))
  return token_entropy(tokens)


def token_entropy(ids: torch.Tensor) -> float:
  """`torch.special.entr(counts / counts.sum()).sum()` over the unique ids
  (main.py:293-296); works on the device the ids live on."""
  _, counts = torch.unique(ids, return_counts=True, sorted=False)
  return torch.special.entr(counts.float() / counts.sum()).sum().item()


def generate_and_score(model, tokenizer=None, num_sample_batches=None,
                       generative_ppl_fn: typing.Optional[typing.Callable] = None) -> dict:
  """The loop of `_gen_ppl_eval` (main.py:259-305) around `model.sample()`:
  returns {'generated_seqs', 'entropy', 'generative_ppl' (None without a scorer)}."""
  tokenizer = tokenizer or model.tokenizer
  n = (num_sample_batches if num_sample_batches is not None
       else model.config.sampling.num_sample_batches)
  ids = [model.sample() for _ in range(n)]
  samples = decode_samples(tokenizer, torch.cat(ids, dim=0))
  length = model.config.model.length
  ppl = None
  if generative_ppl_fn is not None:
    ppl = generative_ppl_fn(samples, max_length=length)
  return {'generative_ppl': ppl,
          'entropy': unigram_entropy(tokenizer, samples, length),
          'generated_seqs': samples}
