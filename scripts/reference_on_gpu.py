"""Context measurement (not a bench line): the UNMODIFIED reference's own
`_cfg_denoise` step (PyTorch eager: cuBLAS, SDPA, flash-attn rotary, ATen pointwise)
running on the B200, config 3's network and shapes, at growing batch sizes until it
runs out of memory.  This is "the reference on the same GPU" -- the bar SURVEY.md 2.2
names -- next to `bench.py`'s CPU arm.  `python scripts/reference_on_gpu.py [B ...]`
writes one JSON line per batch size."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle import ref_model


def main():
  batches = [int(a) for a in sys.argv[1:]] or [32, 64, 128, 256, 512]
  dev = 'cuda'
  for B in batches:
    try:
      cfg_dict = ref_model.config3_dict(B, steps=1000, gamma=2.0)
      model, _, cfg, _, _ = ref_model.build_reference(cfg_dict, ref_model.BertTok(), seed=1,
                                                      device=dev)
      stepper = ref_model.ReferenceStepper(model, cfg, B, device=dev)
      xt = torch.randint(0, 30522, (B, 128), device=dev)
      for i in range(3):
        xt = stepper.step(xt, i)
      torch.cuda.synchronize()
      n = 10
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      e0.record()
      for i in range(n):
        xt = stepper.step(xt, 3 + i)
      e1.record()
      torch.cuda.synchronize()
      ms = e0.elapsed_time(e1) / n
      print(json.dumps({'impl': 'reference on GPU (PyTorch eager, unmodified)', 'batch': B,
                        'ms_per_step': round(ms, 3),
                        'sequences_per_s_at_1000_steps': round(B / ms, 3),
                        'peak_mem_gb': round(torch.cuda.max_memory_allocated() / 2**30, 1)}),
            flush=True)
      del model, stepper, xt
    except torch.OutOfMemoryError:
      print(json.dumps({'impl': 'reference on GPU', 'batch': B, 'error': 'out of memory'}),
            flush=True)
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats()


if __name__ == '__main__':
  main()
