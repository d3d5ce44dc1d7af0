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


# ---------------------------------------------------------------- generative perplexity
class _WordTok:
  """Offline stand-in for the eval model's tokenizer (the reference loads GPT-2's with
  `AutoTokenizer.from_pretrained`; no network here): whitespace words -> ids, the call
  signature `compute_generative_ppl` uses (eval_utils.py:47-58)."""
  eos_token, eos_token_id = '<eos>', 0
  pad_token, pad_token_id = None, None
  vocab = 61

  def __call__(self, sentences, return_tensors='pt', return_token_type_ids=False,
               return_attention_mask=True, truncation=True, padding=True, max_length=None):
    rows = [[1 + (sum(map(ord, w)) * 7 + len(w)) % (self.vocab - 1) for w in s.split()][:max_length]
            for s in sentences]
    width = max(len(r) for r in rows)
    ids = torch.full((len(rows), width), self.pad_token_id, dtype=torch.long)
    mask = torch.zeros(len(rows), width, dtype=torch.long)
    for i, r in enumerate(rows):
      ids[i, :len(r)] = torch.tensor(r)
      mask[i, :len(r)] = 1
    return {'input_ids': ids, 'attention_mask': mask}


def _tiny_lm(seed=0):
  import transformers
  torch.manual_seed(seed)
  cfg = transformers.GPT2Config(vocab_size=_WordTok.vocab, n_positions=64, n_embd=32, n_layer=2,
                                n_head=2, resid_pdrop=0.0, embd_pdrop=0.0, attn_pdrop=0.0)
  return transformers.GPT2LMHeadModel(cfg).eval()


def _sentences(n, seed=3):
  g = torch.Generator().manual_seed(seed)
  words = ['alpha', 'beta', 'gamma', 'delta', 'mol', 'ring', 'the', 'of', 'and', 'C1=CC=CC=C1']
  out = []
  for _ in range(n):
    k = int(torch.randint(3, 20, (1,), generator=g))
    out.append(' '.join(words[int(j)] for j in torch.randint(0, len(words), (k,), generator=g)))
  return out


def test_generative_ppl_matches_restatement_and_reference_rules():
  """eval_utils.py:28-90: exp of the attention-mask-weighted mean next-token NLL under the
  eval model, batches of `gen_ppl_eval_batch_size` with the trailing partial batch DROPPED,
  pad = eos for a tokenizer without a pad token, truncation to `max_length`."""
  lm, tok = _tiny_lm(), _WordTok()
  sents = _sentences(19)
  got = ddg_b200.postprocess.generative_ppl(sents, lm, tok, gen_ppl_eval_batch_size=8, max_length=12,
                                            device='cpu')
  assert tok.pad_token_id == tok.eos_token_id == 0          # the fallback of eval_utils.py:38-42
  # independent restatement in float64, sentence by sentence, over the first 16 sentences only
  num, den = 0.0, 0.0
  with torch.no_grad():
    for s in sents[:16]:
      enc = tok([s], max_length=12)
      ids = enc['input_ids']
      lp = torch.log_softmax(lm(ids)[0].double(), -1)
      nll = -lp[0, :-1].gather(-1, ids[0, 1:, None])[:, 0]
      num += float(nll.sum())
      den += nll.numel()
  assert got == pytest.approx(float(torch.exp(torch.tensor(num / den))), rel=2e-5)
  # the 3 trailing sentences (19 = 2 * 8 + 3) do not count; padding does not count either
  assert ddg_b200.postprocess.generative_ppl(sents[:16], lm, tok, 8, 12, device='cpu') == pytest.approx(got, rel=1e-6)
  # the batch size only groups the work (8 sentences: one batch of 8 == two of 4)
  base = ddg_b200.postprocess.generative_ppl(sents[:8], lm, tok, 8, 12, device='cpu')
  assert ddg_b200.postprocess.generative_ppl(sents[:8], lm, tok, 4, 12, device='cpu') == pytest.approx(base, rel=1e-5)
  # a uniform model has perplexity = vocabulary size
  for p in lm.parameters():
    p.data.zero_()
  assert ddg_b200.postprocess.generative_ppl(sents, lm, tok, 8, 12, device='cpu') == pytest.approx(_WordTok.vocab, rel=1e-5)


@pytest.mark.gpu
@needs_reference
def test_generative_ppl_matches_reference_function_on_gpu(monkeypatch):
  """The reference's own `eval_utils.compute_generative_ppl` (it moves everything to 'cuda')
  with `from_pretrained` answered by a local random-init GPT-2 and the offline tokenizer."""
  import transformers
  ref = refshim.load_reference()
  import importlib
  eval_utils = importlib.import_module('eval_utils')
  lm, tok = _tiny_lm(1), _WordTok()
  monkeypatch.setattr(transformers.AutoTokenizer, 'from_pretrained', staticmethod(lambda *_a, **_k: tok))
  monkeypatch.setattr(transformers.AutoModelForCausalLM, 'from_pretrained', staticmethod(lambda *_a, **_k: lm))
  sents = _sentences(21, seed=8)
  want = eval_utils.compute_generative_ppl(sents, 'local-gpt2', gen_ppl_eval_batch_size=4, max_length=16)
  got = ddg_b200.postprocess.generative_ppl(sents, _tiny_lm(1), _WordTok(), 4, 16, device='cuda')
  assert got == pytest.approx(want, rel=1e-4)
  assert ddg_b200.postprocess.compute_generative_ppl(sents, 'local-gpt2', 4, 16) == pytest.approx(want, rel=1e-4)
