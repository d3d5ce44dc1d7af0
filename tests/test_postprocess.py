"""Callers' post-processing (reference main.py:263-296) against the reference's
own lines run on its own offline tokenizer (`dataloader.Text8Tokenizer`)."""
import pytest
import torch

import ddg_b200
from oracle import refshim

needs_reference = pytest.mark.skipif(
  not refshim.reference_available(),
  reason='no copy of the reference reachable (baseline/_ref is staged by build())')


def _reference_lines(tokenizer, sample, length):
  """main.py:263-296 verbatim in behaviour (the reference has them inline in
  `_gen_ppl_eval`, not as a function): decode, strip, BOS-prefix, re-encode,
  unigram entropy."""
  samples = list(tokenizer.batch_decode(sample))
  tok_bos_token = tokenizer.bos_token if tokenizer.bos_token is not None else tokenizer.cls_token
  samples = [s.replace('[PAD]', '').replace('[MASK]', '').strip() for s in samples]
  samples = [s if s.startswith(tok_bos_token) else f'{tok_bos_token} {s}' for s in samples]
  enc = getattr(tokenizer, 'batch_encode_plus', None) or tokenizer   # transformers >= 5
  tokens = enc(samples, return_tensors='pt', add_special_tokens=False, max_length=length,
               padding='max_length', truncation=True)['input_ids']
  _, counts = torch.unique(torch.as_tensor(tokens), return_counts=True, sorted=False)
  entropy = torch.special.entr(counts.float() / counts.sum()).sum().item()
  return samples, entropy


@needs_reference
def test_decode_and_entropy_match_reference_lines():
  ref = refshim.load_reference()
  tok = ref.dataloader.Text8Tokenizer()
  g = torch.Generator().manual_seed(5)
  L = 48
  ids = torch.randint(0, tok.vocab_size, (16, L), generator=g)
  ids[0, 0] = tok.bos_token_id          # one sample already starts with BOS
  ids[1, 5:9] = tok.pad_token_id        # padding and masks get stripped
  ids[2, 10:20] = tok.mask_token_id
  want_samples, want_entropy = _reference_lines(tok, ids, L)
  got = ddg_b200.postprocess.decode_samples(tok, ids)
  assert got == want_samples
  assert got[0].startswith('[BOS]') and not got[0].startswith('[BOS] [BOS]')
  assert '[PAD]' not in got[1] and '[MASK]' not in got[2]
  assert ddg_b200.postprocess.unigram_entropy(tok, got, L) == pytest.approx(want_entropy, abs=1e-6)


def test_token_entropy_known_values():
  assert ddg_b200.postprocess.token_entropy(torch.zeros(10, dtype=torch.long)) == 0.0
  ids = torch.arange(8).repeat(4)
  assert ddg_b200.postprocess.token_entropy(ids) == pytest.approx(torch.log(torch.tensor(8.0)).item())


@needs_reference
def test_generate_and_score_loop():
  ref = refshim.load_reference()
  tok = ref.dataloader.Text8Tokenizer()

  class FakeModel:
    tokenizer = tok
    config = ddg_b200.config.to_config(
      {'sampling': {'num_sample_batches': 3}, 'model': {'length': 16}})
    calls = 0

    def sample(self):
      self.calls += 1
      return torch.randint(6, 30, (2, 16), generator=torch.Generator().manual_seed(self.calls))
  m = FakeModel()
  out = ddg_b200.postprocess.generate_and_score(
    m, generative_ppl_fn=lambda samples, max_length: float(len(samples)))
  assert m.calls == 3 and len(out['generated_seqs']) == 6 and out['generative_ppl'] == 6.0
  assert out['entropy'] > 0
