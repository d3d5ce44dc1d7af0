"""Callers' post-processing of sampled ids: what the reference's
`_gen_ppl_eval` does with the output of `Diffusion.sample()` (main.py:253-305)
-- `tokenizer.batch_decode`, stripping of [PAD] / [MASK], BOS prefixing, and the
unigram entropy of the re-encoded samples.  Host-side code: the ids are (B, L)
int64, three orders of magnitude less data than one step's logits.

Generative perplexity (`eval_utils.compute_generative_ppl`, eval_utils.py:28-90)
scores the decoded text with a pretrained causal LM (GPT-2 large in the reference's
scripts).  `generative_ppl` is that computation for an eval model / tokenizer the caller
already holds, `compute_generative_ppl` the reference's signature (name or LOCAL path ->
`transformers.*.from_pretrained`; there is no network here, so the tests score with a
random-init GPT-2 built from a config).  `generate_and_score` takes the scorer as an
optional callable.
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


@torch.no_grad()
def generative_ppl(sentences: typing.Sequence[str], eval_model, eval_tokenizer,
                   gen_ppl_eval_batch_size: int = 8, max_length: typing.Optional[int] = 128,
                   eval_context_size: int = 1024, device=None) -> float:
  """eval_utils.py:28-90 with the eval model and tokenizer passed in: the sentences are
  re-tokenised with the eval model's tokenizer (truncation + padding to the longest, pad =
  eos when the tokenizer has no pad token), scored in batches of `gen_ppl_eval_batch_size`
  (a trailing partial batch is dropped, as in the reference) and in chunks of
  `eval_context_size` tokens; perplexity = exp(sum of the next-token NLLs over the
  non-padding target positions / their count) -- `diffusion.Perplexity` (diffusion.py:46-67:
  a weighted mean with the attention mask as weight)."""
  if getattr(eval_tokenizer, 'pad_token', None) is None:
    eval_tokenizer.pad_token = eval_tokenizer.eos_token
    eval_tokenizer.pad_token_id = eval_tokenizer.eos_token_id
  if device is None:
    device = 'cuda' if torch.cuda.is_available() else 'cpu'
  eval_model = eval_model.eval().to(device)
  enc = eval_tokenizer(list(sentences), return_tensors='pt', return_token_type_ids=False,
                       return_attention_mask=True, truncation=True, padding=True,
                       max_length=max_length)
  samples = enc['input_ids'].to(device)
  attn_mask = enc['attention_mask'].to(device)
  nll_sum = torch.zeros((), dtype=torch.float32, device=device)
  weight = torch.zeros((), dtype=torch.float32, device=device)
  bs = gen_ppl_eval_batch_size
  for i in range(samples.shape[0] // bs):
    chunks = zip(torch.split(samples[i * bs:(i + 1) * bs], eval_context_size, dim=-1),
                 torch.split(attn_mask[i * bs:(i + 1) * bs], eval_context_size, dim=-1))
    for sample_chunk, mask_chunk in chunks:
      logits = eval_model(sample_chunk, attention_mask=mask_chunk)[0].transpose(-1, -2)
      nlls = torch.nn.functional.cross_entropy(logits[..., :-1], sample_chunk[..., 1:],
                                               reduction='none')
      w = mask_chunk[..., 1:].to(nlls.dtype)
      nll_sum += (nlls * w).sum().float()
      weight += w.sum().float()
  return torch.exp(nll_sum / weight).item()


def compute_generative_ppl(sentences, eval_model_name_or_path, gen_ppl_eval_batch_size=8,
                           max_length=128) -> float:
  """The reference's entry point (eval_utils.py:28-33): loads tokenizer and causal LM with
  `transformers` (a local directory works offline) and scores with `generative_ppl`."""
  import os
  import transformers
  os.environ['TOKENIZERS_PARALLELISM'] = 'false'
  tok = transformers.AutoTokenizer.from_pretrained(eval_model_name_or_path)
  lm = transformers.AutoModelForCausalLM.from_pretrained(eval_model_name_or_path)
  return generative_ppl(sentences, lm, tok, gen_ppl_eval_batch_size, max_length)


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
