"""Benchmark of the reverse-diffusion sampling hot path (BASELINE.json metric:
sampled sequences/sec, LM1B UDLM D-CFG, L=128, V=30522, sampling.steps=1000).

    python bench.py --gpus N --steps K --warmup W          # this repo
    python bench.py --impl reference --gpus N ...          # CPU reference arm
    python bench.py --config c1|c2|c3|c4|c4_exact|c5|c5_cbg  # another BASELINE.json shape

A bench "step" is ONE reverse-diffusion step over the whole per-GPU batch:
the DiT-small forward on the batched cond+uncond pair (2*B*L token rows) plus
the fused sampler.  Every step of config 3 has identical shapes and no
data-dependent work, so   sequences/sec = B_total / (sampling.steps * mean step time);
with --steps 1000 the timed region is literally one full sample() of config 3.

Scaling.  BASELINE.json configs[2] is "batch 512 on 8xB200": at N > 1 the
headline is STRONG scaling -- the 512 sequences of the config are sharded
512/N per GPU (no data-path collective; one NCCL all_gather of the ids at the
end) -- and the weak-scaling number (512 sequences on every GPU) is reported
next to it under "weak".

The default line also carries `configs`: a short measurement of every other
BASELINE.json shape (C1, C2, C4, C5) sharded the same way.
"""
from __future__ import annotations

import argparse
import dataclasses
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

V, L, GAMMA = 30522, 128, 2.0
METRIC = 'sampled sequences/sec, LM1B UDLM D-CFG L=128 T=1000'


class BertTok:
  vocab_size, mask_token, mask_token_id = V, '[MASK]', 103
  bos_token_id, eos_token_id, pad_token_id = 101, 102, 0


class Text8Tok:       # reference dataloader.py:49-101: 35 ids, [MASK]=4
  vocab_size, mask_token, mask_token_id = 35, '[MASK]', 4
  bos_token_id, eos_token_id, pad_token_id = 0, 1, 3


class Qm9Tok:         # notebooks/qm9_vocab.json: 30 tokens + specials; scripts use 40
  vocab_size, mask_token, mask_token_id = 40, '[MASK]', 4
  bos_token_id, eos_token_id, pad_token_id = 2, 3, 5


@dataclasses.dataclass(frozen=True)
class Workload:
  key: str
  desc: str                    # BASELINE.json configs[i], spelled out
  overrides: tuple
  tok: type
  batch: int                   # total sequences of the config
  sampling_steps: int          # sampling.steps of the config (seq/s denominator)
  cbg: bool = False
  classifier: str = ''         # classifier_model group for cbg
  metric: str = ''


# sampling.use_cache=False: with time conditioning the reference's whole-batch cache rule
# (diffusion.py:1371-1374, honoured by the step loop) skips the forward of a step when no token of
# the batch changed in the previous one -- late in a 1000-step run, and the earlier the smaller the
# per-GPU batch.  The timed loop must not contain skipped work, so every bench step is a full step.
_UDLM = ('parameterization=d3pm', 'diffusion=uniform', 'time_conditioning=True', 'T=0',
         'sampling.use_cache=False')
_CFG = ('guidance=cfg', f'guidance.gamma={GAMMA}', 'guidance.condition=1',
        '+data.num_classes=2')
_CBG = ('classifier_backbone=dit', 'guidance=cbg', 'guidance.gamma=2.0',
        'guidance.condition=1', '+data.num_classes=2', 'training.guidance=null')
WORKLOADS = {w.key: w for w in [
  Workload('c1', 'BASELINE.json configs[0]: text8 UDLM DiT-tiny random-init, L=256, vocab 35 '
           '(27 letters + specials), sampling.steps=128, batch 4, guidance=null',
           _UDLM + ('model=tiny', 'model.length=256', 'training.guidance=null'),
           Text8Tok, 4, 128, metric='sampled sequences/sec, text8 UDLM DiT-tiny L=256 T=128'),
  Workload('c2', 'BASELINE.json configs[1]: LM1B MDLM (subs, absorbing) DiT-small, L=128, vocab '
           '30522, sampling.steps=1000, guidance=null, batch 512',
           ('parameterization=subs', 'diffusion=absorbing_state', 'time_conditioning=False',
            'model=small', f'model.length={L}', 'training.guidance=null'),
           BertTok, 512, 1000, metric='sampled sequences/sec, LM1B MDLM L=128 T=1000'),
  Workload('c3', 'BASELINE.json configs[2]: LM1B UDLM (d3pm, uniform) DiT-small, L=128, '
           'vocab 30522, sampling.steps=1000, guidance=cfg gamma=2, batch 512',
           _UDLM + ('model=small', f'model.length={L}') + _CFG, BertTok, 512, 1000,
           metric=METRIC),
  Workload('c4', 'BASELINE.json configs[3]: QM9 SMILES UDLM DiT-small, L=32, vocab 40, '
           'sampling.steps=32, guidance=cbg (first-order, guidance.use_approx=True) with the '
           'DiT-tiny classifier, batch 4096',
           _UDLM + ('model=small', 'model.length=32', 'guidance.use_approx=True') + _CBG,
           Qm9Tok, 4096, 32, cbg=True, classifier='tiny-classifier',
           metric='sampled sequences/sec, QM9 UDLM D-CBG (first-order) L=32 T=32'),
  Workload('c4_exact', 'BASELINE.json configs[3] with the EXACT classifier term (all L*V '
           'single-token edits per sequence per step, guidance.use_approx=False), batch 64',
           _UDLM + ('model=small', 'model.length=32', 'guidance.use_approx=False') + _CBG,
           Qm9Tok, 64, 32, cbg=True, classifier='tiny-classifier',
           metric='sampled sequences/sec, QM9 UDLM D-CBG (exact) L=32 T=32'),
  Workload('c5', 'BASELINE.json configs[4], D-CFG half: Amazon-polarity UDLM DiT-small, L=128, '
           'vocab 30522, sampling.steps=128, guidance=cfg gamma=2, batch 1024',
           _UDLM + ('model=small', f'model.length={L}') + _CFG, BertTok, 1024, 128,
           metric='sampled sequences/sec, Amazon UDLM D-CFG L=128 T=128'),
  Workload('c5_cbg', 'BASELINE.json configs[4], D-CBG half: Amazon-polarity UDLM DiT-small + '
           'DiT-small classifier, L=128, vocab 30522, sampling.steps=128, guidance=cbg '
           '(first-order), batch 1024',
           _UDLM + ('model=small', f'model.length={L}', 'guidance.use_approx=True') + _CBG,
           BertTok, 1024, 128, cbg=True, classifier='small-classifier',
           metric='sampled sequences/sec, Amazon UDLM D-CBG (first-order) L=128 T=128'),
]}
SIDE_CONFIGS = ('c1', 'c2', 'c4', 'c4_exact', 'c5', 'c5_cbg')


def peaks():
  try:
    with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
      p = json.load(f)
    return dict(hbm=p['hbm_gbs'], tf_burst=p['bf16_tflops'],
                tf_sustained=p['bf16_tflops_sustained'], src='measured')
  except Exception:
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src='fallback')


class ClockSampler:
  """SM clock, board power and throttle reasons sampled DURING the timed region: NVML polled every
  20 ms from a thread (a timed region at 64 sequences per GPU is ~0.1 s: `nvidia-smi -lms` would
  return at most one sample), `nvidia-smi` when NVML is unavailable."""
  Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
       'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
       'clocks_event_reasons.sw_power_cap')
  REASON_BITS = (('sw_power_cap', 0x4), ('hw_slowdown', 0x8), ('sw_thermal_slowdown', 0x20),
                 ('hw_thermal_slowdown', 0x40))

  def __init__(self, index):
    self.index, self.rows, self.proc, self.nv, self.h = index, [], None, None, None
    self.stop_flag = threading.Event()
    try:
      import pynvml
      pynvml.nvmlInit()
      self.h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(index))
      self.nv = pynvml
      self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
    except Exception:
      self.nv = None

  @staticmethod
  def _physical_index(index):
    vis = os.environ.get('CUDA_VISIBLE_DEVICES', '')
    ids = [v for v in vis.split(',') if v.strip() != '']
    if ids and all(v.strip().isdigit() for v in ids) and index < len(ids):
      return int(ids[index])
    return index

  def _poll(self):
    nv = self.nv
    while True:
      try:
        sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
        pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
        try:
          bits = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
        except Exception:
          bits = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
        self.rows.append((sm, pw, bits))
      except Exception:
        pass
      if self.stop_flag.wait(0.02):
        break

  def start(self):
    if self.nv is not None:
      self.t = threading.Thread(target=self._poll, daemon=True)
      self.t.start()
      return
    try:
      self.proc = subprocess.Popen(
        ['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
         '-i', str(self.index), '-lms', '100'], stdout=subprocess.PIPE, text=True)
      self.t = threading.Thread(target=self._read, daemon=True)
      self.t.start()
    except Exception:
      self.proc = None

  def _read(self):
    for line in self.proc.stdout:
      self.rows.append([x.strip() for x in line.split(',')])

  def stop(self):
    if self.nv is not None:
      self.stop_flag.set()
      self.t.join(timeout=2)
      sm = [r[0] for r in self.rows]
      pw = [r[1] for r in self.rows]
      bits = 0
      for r in self.rows:
        bits |= r[2]
      return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': self.max_mhz,
              'reasons': sorted(n for n, b in self.REASON_BITS if bits & b),
              'power_w': round(statistics.median(pw), 2) if pw else None, 'samples': len(sm),
              'source': 'nvml, 20 ms'}
    if self.proc is None:
      return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unavailable']}
    self.proc.terminate()
    try:
      self.proc.wait(timeout=5)
    except Exception:
      self.proc.kill()
    sm = [float(r[0]) for r in self.rows if r and r[0].replace('.', '').isdigit()]
    mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace('.', '').isdigit()]
    pw = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace('.', '').isdigit()]
    names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
    reasons = sorted({n for r in self.rows if len(r) >= 7
                      for n, v in zip(names, r[3:7]) if v.lower().startswith('active')})
    return {'sm_mhz': statistics.median(sm) if sm else None,
            'sm_max_mhz': max(mx) if mx else None, 'reasons': reasons,
            'power_w': statistics.median(pw) if pw else None, 'samples': len(sm),
            'source': 'nvidia-smi -lms 100'}


# ---------------------------------------------------------------- CPU arm
def reference_kind():
  from oracle import refshim
  return 'reference' if refshim.reference_available() else 'port'


def cpu_reference_steps(batch, steps, warmup, budget_s=None):
  """The reference's OWN `Diffusion._cfg_denoise` (unmodified code, imported by
  oracle/refshim.py from /root/reference or its staged copy baseline/_ref) on
  config 3's network and shapes at a small batch, fp32 on all host cores, Philox
  replaced by nothing: the reference draws its own `torch.rand_like` noise.
  Returns (s/step, cores, steps actually timed, kind).  Falls back to the oracle
  port only when no copy of the reference is reachable (kind = "port").  Never
  touches the product package or its CUDA library."""
  cores = os.cpu_count() or 1
  torch.set_num_threads(cores)
  kind = reference_kind()
  g = torch.Generator().manual_seed(0)
  xt = torch.randint(0, V, (batch, L), generator=g)
  if kind == 'reference':
    from oracle import ref_model
    cfg_dict = ref_model.config3_dict(batch, steps=1000, gamma=GAMMA)
    model, _, cfg, _, _ = ref_model.build_reference(cfg_dict, ref_model.BertTok(), seed=1)
    stepper = ref_model.ReferenceStepper(model, cfg, batch)

    def one(i, x):
      return stepper.step(x, i)
  else:
    from oracle import ref_model, ref_sampler
    cfg_dict = ref_model.config3_dict(batch, steps=1000, gamma=GAMMA)
    sd = ref_model.random_weights(cfg_dict, V, seed=1)
    oracle = ref_sampler.OracleSampler(
      sd, vocab_size=V, mask_index=103, parameterization='d3pm', diffusion='uniform',
      n_heads=12, time_conditioning=True, num_classes=2)
    sig, mct, mcs = ref_sampler.step_scalars(1000, 1e-5, 0)
    cond = torch.full((batch,), 1, dtype=torch.long)

    def one(i, x):
      U = torch.rand(batch, L, V, generator=g)
      with torch.no_grad():
        xs, _, _ = oracle.step(x, sig[i].expand(batch), mct[i].expand(batch)[:, None, None],
                               mcs[i].expand(batch)[:, None, None], U, guidance='cfg',
                               cond=cond, gamma=GAMMA)
      return xs
  times = []
  t_start = time.perf_counter()
  for i in range(warmup + steps):
    t0 = time.perf_counter()
    xt = one(i, xt)
    dt = time.perf_counter() - t0
    if i >= warmup:
      times.append(dt)
      if budget_s is not None and time.perf_counter() - t_start > budget_s:
        break
  return sum(times) / len(times), cores, len(times), kind


def run_reference_arm(args, rank):
  if rank != 0:
    return
  wl = WORKLOADS['c3']
  batch = args.ref_batch
  steps, warmup = max(1, args.steps), max(1, min(args.warmup, 3))
  s_per_step, cores, timed, kind = cpu_reference_steps(batch, steps, warmup,
                                                       budget_s=args.ref_budget_s)
  value = batch / (wl.sampling_steps * s_per_step)
  what = ("the reference's own Diffusion._cfg_denoise (unmodified code, fp32 torch on CPU)"
          if kind == 'reference' else 'oracle/ CPU restatement of the reference step (fp32 torch)')
  sample = (f'{what}, {cores} threads, config-3 network and shapes on a bounded sample of '
            f'{batch} of the 512 sequences, {timed} reverse steps timed after {warmup} warm-up'
            + ('' if timed == steps else f' (stopped at the {args.ref_budget_s:.0f} s budget; '
               f'{steps} requested)') + '; sequences/s extrapolated linearly in batch and steps')
  line = {
    'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'sequences/s',
    'n_gpus': args.gpus, 'steps': timed, 'warmup': warmup,
    'ms_per_step': s_per_step * 1e3, 'higher_is_better': True,
    'scaling': 'strong' if args.gpus > 1 else 'weak',
    'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
    'config': dict(workload_config(wl, max(1, wl.batch // args.gpus), args.gpus),
                   sample_batch=batch),
    'cpu_baseline': {'value': value, 'unit': 'sequences/s', 'cores': cores,
                     'kind': kind, 'sample': sample},
    'e2e': {'value': value, 'unit': 'sequences/s', 'h2d_bytes_per_step': 0,
            'd2h_bytes_per_step': 0},
    'gpu_launches': 0,
    # evidence that this arm is the reference alone: no product module, no product .so
    'product_modules_loaded': sorted(m for m in sys.modules if m.startswith('ddg_b200')),
    'native_so_loaded': _mapped_product_libs(),
  }
  print(json.dumps(line), flush=True)


def _mapped_product_libs():
  try:
    with open('/proc/self/maps') as f:
      return sorted({l.split()[-1] for l in f if 'libddg' in l})
  except OSError:
    return None


def workload_config(wl, batch, world):
  cfg = {
    'workload': wl.desc,
    'step': ('one reverse-diffusion step = DiT forward (cond+uncond pair batched for D-CFG; '
             '+ classifier forward/backward for D-CBG) + fused sampler'),
    'batch_per_gpu': batch, 'global_batch': batch * world, 'vocab': wl.tok.vocab_size,
    'sampling_steps': wl.sampling_steps, 'parallelism': f'batch-sharded x{world}',
    'cache': 'per-step working set (>= 1 GB of logits per 64 sequences) exceeds the 126 MB L2; '
             'no flush needed',
  }
  if 'sampling.use_cache=False' in wl.overrides:
    cfg['use_cache'] = ('off in the timed loop: every step runs the full forward (the reference\'s '
                        'whole-batch cache rule would skip steps in which no token changed)')
  return cfg


# ---------------------------------------------------------------- GPU arm
def build_model(wl, batch, device, seed, steps=None):
  import ddg_b200
  over = list(wl.overrides) + [f'sampling.batch_size={batch}',
                               f'sampling.steps={steps or wl.sampling_steps}']
  if wl.classifier:
    over.append(f'classifier_model={wl.classifier}')
  cfg = ddg_b200.compose(over)
  tok = wl.tok()
  sd = ddg_b200.weights.random_state_dict(cfg, tok.vocab_size, seed=1)
  model = ddg_b200.Diffusion(cfg, tok, device=device).load_state_dict(sd)
  model.seed = seed
  if wl.cbg:
    csd = ddg_b200.weights.random_state_dict(cfg, tok.vocab_size, seed=2, classifier=True)
    model.set_classifier(ddg_b200.Classifier(cfg, tok, device=device).load_state_dict(csd))
  if os.environ.get('DDG_RACE_GROUP'):             # sampler A/B knob (0 | 8 | 64)
    model.race_group = int(os.environ['DDG_RACE_GROUP'])
  if os.environ.get('DDG_FUSE_LN'):                # LayerNorm A/B knob (0 = separate launches)
    model.backbone.fuse_ln = os.environ['DDG_FUSE_LN'] != '0'
    if model.classifier_model is not None:
      model.classifier_model.classifier_model.fuse_ln = model.backbone.fuse_ln
  return model, cfg


def _max_over_ranks(x, device, world):
  if world > 1:
    t = torch.tensor([x], device=device, dtype=torch.float64)
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    return float(t.item())
  return x


def _barrier(world):
  if world > 1:
    torch.distributed.barrier()
  torch.cuda.synchronize()


class GraphRun:
  """Steady-state timing of the captured step loop (`_GraphLoop`) of a workload
  (null / cfg / cbg guidance)."""

  def __init__(self, wl, B, device, seed, table_steps=None):
    self.wl, self.B = wl, B
    self.model, self.cfg = build_model(wl, B, device, seed)
    m, cfg = self.model, self.cfg
    self.Lw = cfg.model.length
    guidance = cfg.get('guidance')
    method = None if guidance is None else guidance.method
    self.cond = (torch.full((B,), guidance.condition, dtype=torch.long, device=device)
                 if method == 'cfg' else None)
    self.table_steps = table_steps or wl.sampling_steps
    self.xt0 = m._sample_prior(B, self.Lw)
    self.tables = m._step_scalars(self.table_steps, 1e-5)
    self.loop = m._get_loop(B, self.Lw, method, guidance)
    self.loop.prepare(self.xt0, *self.tables, self.cond)
    self.pos = 0

  def replay(self, n):
    """n steps, wrapping the device step counter at the end of the schedule tables."""
    done = 0
    while done < n:
      if self.pos == self.table_steps:
        self.loop.reset(self.xt0)
        self.pos = 0
      k = min(n - done, self.table_steps - self.pos)
      self.loop.replay(k)
      done += k
      self.pos += k

  def restart(self):
    self.loop.reset(self.xt0)
    self.pos = 0

  def timed(self, K, world, device, gather=True, clock_index=None):
    from ddg_b200 import dist as ddist
    _barrier(world)
    clocks = None
    if clock_index is not None:
      clocks = ClockSampler(clock_index)
      clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    self.replay(K)
    ids = ddist.gather_samples(self.loop.xt) if gather else self.loop.xt  # the one collective
    e1.record()
    _barrier(world)
    clk = clocks.stop() if clocks is not None else None
    ms = _max_over_ranks(e0.elapsed_time(e1), device, world)
    assert ids.shape[0] == (world if gather else 1) * self.B
    return ms / K, clk


def time_cbg(wl, B, device, seed, K, W, world):
  """D-CBG workloads: the captured step loop of `sample()` (denoiser forward, classifier
  forward + backward or exact single-token edits, guided sampler in one CUDA graph)."""
  run = GraphRun(wl, B, device, seed)
  run.replay(W)
  run.restart()
  ms, _ = run.timed(K, world, device, gather=True)
  xt = run.loop.xt
  assert int(xt.min()) >= 0 and int(xt.max()) < wl.tok.vocab_size
  del run
  return ms


def shard(total, world):
  return max(1, total // world)


def side_config(key, world, rank, device, K, W):
  """Short measurement of another BASELINE.json shape, sharded like the headline."""
  from ddg_b200 import dist as ddist
  wl = WORKLOADS[key]
  B = shard(wl.batch, world)
  try:
    if wl.cbg:
      Kc = max(2, min(K, 6 if key != 'c4_exact' else 3))
      ms = time_cbg(wl, B, device, ddist.rank_seed(1, rank), Kc, 2, world)
      steps_timed = Kc
    else:
      absorbing = 'diffusion=absorbing_state' in wl.overrides
      # MDLM: the work per step depends on where the run is (masked fraction 1 -> 0, and
      # with the per-sequence logits cache on how many sequences a step changes, which is
      # set by sampling.steps), so time one WHOLE sample() of the config's own 1000 steps;
      # uniform: steps are shape-identical
      Kc = wl.sampling_steps if absorbing else max(8, min(K, 24))
      run = GraphRun(wl, B, device, ddist.rank_seed(1, rank))
      run.replay(min(W, 8))
      run.restart()
      ms, _ = run.timed(Kc, world, device, gather=True)
      steps_timed = Kc
      del run
  except torch.OutOfMemoryError as e:      # never take the box down for a side line
    torch.cuda.empty_cache()
    return {'error': f'out of memory at batch_per_gpu={B}: {str(e)[:80]}'}
  torch.cuda.empty_cache()
  return {'workload': wl.desc, 'batch_per_gpu': B, 'global_batch': B * world,
          'ms_per_step': round(ms, 4), 'steps_timed': steps_timed,
          'value': round(B * world / (wl.sampling_steps * ms * 1e-3), 3),
          'unit': 'sequences/s'}


def flops_per_step(cfg, batch, n_fwd, Lw, vocab):
  h, nb = cfg.model.hidden_size, cfg.model.n_blocks
  rows = n_fwd * batch * Lw
  per_token = nb * (24 * h * h + 4 * Lw * h) + 2 * h * vocab
  gemm = {
    'qkv': 2 * rows * h * 3 * h * nb, 'attn_out': 2 * rows * h * h * nb,
    'fc1': 2 * rows * h * 4 * h * nb, 'fc2': 2 * rows * 4 * h * h * nb,
    'vocab': 2 * rows * h * vocab}
  return rows * per_token, gemm


def ncu_traffic(batch):
  """DRAM bytes per launch of each kernel family from the newest committed ncu
  capture (profiles/r0N_traffic.json); None when it was taken at another batch."""
  for name in ('r02_traffic.json', 'r01_traffic.json'):
    try:
      with open(os.path.join(ROOT, 'profiles', name)) as f:
        t = json.load(f)
      if t.get('batch_per_gpu') == batch:
        return t['bytes_per_launch'], name
    except Exception:
      continue
  return None, None


def kernel_breakdown(model, cfg, loop):
  """One instrumented eager step: CUDA events around every launch group on
  the launching stream -> per-kernel-family device time."""
  from ddg_b200 import _lib
  lib = _lib.lib
  names = {'ddg_gemm': None, 'ddg_attention': 'attention', 'ddg_attention_live': 'attention',
           'ddg_ln_modulate': 'ln_modulate', 'ddg_ln_modulate_live': 'ln_modulate',
           'ddg_sample_step': 'sample_step', 'ddg_cond_embed': 'cond', 'ddg_adaln': 'cond',
           'ddg_step_advance': 'cond', 'ddg_seq_compact': 'cond'}
  records = []
  fused_ln = set()     # GEMM families whose launches also do the following LayerNorm + modulate
  orig = {n: getattr(lib, n) for n in names}
  h = cfg.model.hidden_size

  def wrap(name):
    fn = orig[name]

    def inner(*a):
      label = names[name]
      if name == 'ddg_gemm':
        g = a[0]
        g = g._obj if hasattr(g, '_obj') else g
        label = {(3 * h, h): 'gemm_qkv', (h, h): 'gemm_attn_out', (4 * h, h): 'gemm_fc1',
                 (h, 4 * h): 'gemm_fc2'}.get((g.N, g.K), 'gemm_vocab')
        if g.ln_weight:
          fused_ln.add(label)
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      e0.record()
      rc = fn(*a)
      e1.record()
      records.append((label, e0, e1))
      return rc
    return inner
  try:
    for n in names:
      setattr(lib, n, wrap(n))
    loop._body()
    torch.cuda.synchronize()
  finally:
    for n in names:
      setattr(lib, n, orig[n])
  out = {}
  for label, e0, e1 in records:
    d = out.setdefault(label, {'ms': 0.0, 'launches': 0})
    d['ms'] += e0.elapsed_time(e1)
    d['launches'] += 1
  for label in fused_ln:
    out[label]['fused_ln'] = True
  return out


def sampler_alone_ms(model, loop, reps=10, race_group=None):
  """The step's fused-sampler launch (same buffers, same arguments) repeated on its own:
  what the kernel does when the SM clock is not held down by the GEMMs of the step.
  `race_group=0`: the per-column race kernel whose ids are bit-comparable with the
  reference under injected noise (here with its own Philox stream)."""
  from ddg_b200 import _lib
  B, Lw = loop.B, loop.L
  _, mct, mcs = loop.tables
  ld = model.backbone.vocab_ld
  xs = torch.empty_like(loop.xt)
  saved = model.race_group
  if race_group is not None:
    model.race_group = race_group

  def launch():
    model._launch_sampler(
      logits_a=loop.logits.data_ptr(), ld_a=ld,
      logits_b=(loop.logits.data_ptr() + B * Lw * ld * 2) if loop.two else None,
      ld_b=ld if loop.two else 0, xt=loop.xt, xs=xs, mc_t=mct, mc_s=mcs, rows=B * Lw,
      seq_len=Lw, mc_seq_stride=0, mc_step_stride=1, step_ptr=loop.step,
      seed_ptr=loop.seed_dev, stats=(loop.stats, 0, B * Lw if loop.two else 0),
      guidance=_lib.GUIDE_CFG if loop.two else _lib.GUIDE_NONE,
      gamma=loop.gamma if loop.two else 1.0)
  try:
    loop.step.zero_()
    torch.cuda.synchronize()
    time.sleep(0.5)
    for _ in range(3):
      launch()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
      launch()
    e1.record()
    torch.cuda.synchronize()
  finally:
    model.race_group = saved
  return e0.elapsed_time(e1) / reps


def e2e_cfg(model, cfg, B, xt0, tables, world, device, Ke):
  """Same metric through the per-step public API with HOST buffers: pinned host xt
  -> device, `_cfg_denoise`, xs -> pinned host, every step."""
  sig, mct, mcs = tables
  cond = torch.full((B,), cfg.guidance.condition, dtype=torch.long, device=device)
  host_xt = torch.empty(B, L, dtype=torch.int64).pin_memory()
  host_xt.copy_(xt0.cpu())
  host_xs = torch.empty(B, L, dtype=torch.int64).pin_memory()
  # schedule scalars are constants of the run, resident on the device; the per-step INPUT is xt
  sig_d, mct_d, mcs_d = (t.to(device, torch.float32) for t in (sig, mct, mcs))

  def e2e_step(i):
    x = host_xt.to(device, non_blocking=True)
    xs, _, _ = model._cfg_denoise(cond, GAMMA, x, sig_d[i].expand(B),
                                  mct_d[i].expand(B)[:, None, None],
                                  mcs_d[i].expand(B)[:, None, None])
    host_xs.copy_(xs, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    host_xt.copy_(host_xs)
  for i in range(3):
    e2e_step(i)
  _barrier(world)
  t0 = time.perf_counter()
  for i in range(Ke):
    e2e_step(3 + i)
  torch.cuda.synchronize()
  return _max_over_ranks((time.perf_counter() - t0) / Ke, device, world)


def run_gpu_arm(args, rank, world, local_rank):
  import ddg_b200  # noqa: F401
  from ddg_b200 import dist as ddist
  torch.cuda.set_device(local_rank)
  device = torch.device('cuda', local_rank)
  wl = WORKLOADS[args.config]
  pk = peaks()
  K, W = args.steps, max(args.warmup, 3)
  B = args.batch or shard(wl.batch, world)     # STRONG scaling: the config's batch / N
  seed = ddist.rank_seed(1, rank)

  if wl.cbg:
    ms_per_step = time_cbg(wl, B, device, seed, max(2, min(K, 20)), min(W, 3), world)
    if rank == 0:
      print(json.dumps({
        'metric': wl.metric, 'value': B * world / (wl.sampling_steps * ms_per_step * 1e-3),
        'unit': 'sequences/s', 'n_gpus': world, 'steps': max(2, min(K, 20)),
        'warmup': min(W, 3), 'ms_per_step': ms_per_step, 'higher_is_better': True,
        'scaling': 'strong' if world > 1 else 'weak', 'vs_baseline': None,
        'dtype': 'fp16 denoiser / bf16 classifier blocks, bf16 head, fp32 accumulate',
        'data': 'synthetic (random-init networks, Philox prior and noise)',
        'config': workload_config(wl, B, world)}), flush=True)
    return

  # ---- graph loop: tables for a full run of the config, one captured step
  run = GraphRun(wl, B, device, seed)
  model, cfg, loop = run.model, run.cfg, run.loop
  if 'diffusion=absorbing_state' in wl.overrides:
    # data-dependent work per step: the timed region is one whole sample() of the config
    run.replay(W)
    run.restart()
    K = wl.sampling_steps
    args.no_settle = True
  # warm-up: W steps, and at least ~2 s of replay so the board reaches its power-capped
  # steady state before the timed region (clocks settle over the first second)
  if args.no_settle:
    settle, W_eff = 0, W
  else:
    run.replay(W)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run.replay(8)
    torch.cuda.synchronize()
    per = (time.perf_counter() - t0) / 8
    settle = int(min(400, max(0, 2.0 / max(per, 1e-4) - W - 8)))
    run.replay(settle)
    W_eff = W + 8 + settle
  ms_per_step, clk = run.timed(K, world, device, gather=True, clock_index=local_rank)
  value = world * B / (wl.sampling_steps * ms_per_step * 1e-3)

  line = {
    'metric': wl.metric, 'value': value, 'unit': 'sequences/s', 'n_gpus': world,
    'steps': K, 'warmup': W_eff, 'ms_per_step': ms_per_step, 'higher_is_better': True,
    'scaling': 'strong' if world > 1 else 'weak', 'vs_baseline': None,
    'dtype': 'fp16 blocks / bf16 head, fp32 accumulate',
    'data': 'synthetic (random-init DiT, Philox prior and noise)',
    'config': workload_config(wl, B, world)}
  if world > 1:
    line['config']['scaling_note'] = (
      f'strong scaling: the {wl.batch} sequences of the config sharded {B} per GPU; '
      'the weak-scaling figure (the full batch on every GPU) is under "weak"')

  if args.config != 'c3':
    if rank == 0:
      line['clocks'] = clk
      print(json.dumps(line), flush=True)
    return

  # ---- e2e: the per-step public API with HOST buffers, copies inside the timing
  Ke = max(3, min(K, 40))
  e2e_s = e2e_cfg(model, cfg, B, run.xt0, run.tables, world, device, Ke)
  line['e2e'] = {'value': world * B / (wl.sampling_steps * e2e_s), 'unit': 'sequences/s',
                 'h2d_bytes_per_step': B * L * 8 + 3 * 4, 'd2h_bytes_per_step': B * L * 8,
                 'api': 'Diffusion._cfg_denoise per step, pinned host xt in / xs out',
                 'steps': Ke}

  # ---- per-kernel roofline (every rank runs it so ranks stay in step; rank 0 reports)
  n_fwd = 2
  total_flops, gemm_flops = flops_per_step(cfg, B, n_fwd, L, V)
  kb = kernel_breakdown(model, cfg, loop)
  step_kernel_ms = sum(d['ms'] for d in kb.values())
  gemm_keys = [k for k in kb if k.startswith('gemm_')]
  gemm_ms = sum(kb[k]['ms'] for k in gemm_keys)
  gemm_launches = sum(kb[k]['launches'] for k in gemm_keys)
  gemm_total_flops = sum(gemm_flops.values())
  kernels = {}
  h_ = cfg.model.hidden_size
  rows = n_fwd * B * L
  g_k = {'gemm_attn_out': h_, 'gemm_fc2': 4 * h_}
  pure = [k for k in gemm_keys if not kb[k].get('fused_ln')]
  pure_tf = (sum(gemm_flops[k[5:]] for k in pure) / (sum(kb[k]['ms'] for k in pure) * 1e-3) / 1e12
             if pure else None)
  for k, d in sorted(kb.items(), key=lambda kv: -kv[1]['ms']):
    ent = {'ms_per_step': round(d['ms'], 4), 'launches': d['launches'],
           'share': round(d['ms'] / step_kernel_ms, 4)}
    if k.startswith('gemm_') and k[5:] in gemm_flops:
      fl = gemm_flops[k[5:]]
      ent['tflops'] = round(fl / (d['ms'] * 1e-3) / 1e12, 1)
      ent['frac_of_sustained_peak'] = round(ent['tflops'] / pk['tf_sustained'], 4)
      if d.get('fused_ln'):
        # the launch also normalises the rows it has just updated (the next LayerNorm + modulate):
        # against HBM it moves A + x in/out + the 16-bit norm output
        byts = rows * (g_k[k] * 2 + h_ * 4 * 2 + h_ * 2) * d['launches']
        ent['fused'] = 'LayerNorm + adaLN modulate of the updated residual rows (was a separate launch)'
        ent['hbm_gbs'] = round(byts / (d['ms'] * 1e-3) / 1e9, 1)
        ent['frac_of_hbm_peak'] = round(ent['hbm_gbs'] / pk['hbm'], 4)
    kernels[k] = ent
  samp_bytes = B * L * (2 * V * 2 + 16)
  samp_ms = kb['sample_step']['ms']
  samp_gbs = samp_bytes / (samp_ms * 1e-3) / 1e9
  samp_alone = sampler_alone_ms(model, loop)
  samp_alone_gbs = samp_bytes / (samp_alone * 1e-3) / 1e9
  samp_col = sampler_alone_ms(model, loop, race_group=0)
  gemm_tf = gemm_total_flops / (gemm_ms * 1e-3) / 1e12
  n_launch = sum(d['launches'] for d in kb.values())
  traffic, traffic_src = ncu_traffic(B)
  gemm_traffic = samp_traffic = None
  if traffic:
    try:   # per launch, averaged over the GEMM launches of a step like `achieved`
      gemm_traffic = sum(traffic[k] * kb[k]['launches'] for k in gemm_keys) / gemm_launches
      samp_traffic = traffic['sample_step']
    except KeyError:
      pass
  line.update({
    'gpu_launches': n_launch * K,
    'roofline': {
      'kernel': 'gemm_kernel (tcgen05; all DiT linears of one step)', 'bound': 'tensor',
      'achieved': round(gemm_tf, 1), 'peak': pk['tf_sustained'], 'unit': 'TFLOP/s',
      'frac': round(gemm_tf / pk['tf_sustained'], 4), 'traffic': gemm_traffic,
      'traffic_note': f'ncu dram bytes per launch, mean over the GEMM launches of a step '
                      f'(profiles/{traffic_src})',
      'flops_per_launch': gemm_total_flops / gemm_launches,
      'ms_per_launch': gemm_ms / gemm_launches,
      'peak_source': f"{pk['src']} sustained bf16 (kernel timed inside a long step)",
      'share_of_step': round(gemm_ms / step_kernel_ms, 4),
      'note': ('GEMM FLOPs over the time of ALL GEMM launches of a step; the 24 residual GEMMs '
               '(attn-out, fc2) also do the following LayerNorm + modulate in their epilogue '
               '(24 of the 25 norms of a step), so their time is not GEMM time alone'
               if any(kb[k].get('fused_ln') for k in gemm_keys) else None),
      'gemms_without_fused_layernorm': (
        {'kernels': pure, 'achieved': round(pure_tf, 1),
         'frac': round(pure_tf / pk['tf_sustained'], 4)} if pure_tf else None)},
    'roofline_sampler': {
      'kernel': 'sample_group_kernel (fused sampler, hierarchical race)', 'bound': 'hbm',
      'achieved': round(samp_gbs, 1), 'peak': pk['hbm'], 'unit': 'GB/s',
      'frac': round(samp_gbs / pk['hbm'], 4), 'traffic': samp_traffic,
      'algorithmic_bytes': samp_bytes, 'peak_source': pk['src'],
      'share_of_step': round(samp_ms / step_kernel_ms, 4),
      'note': 'timed inside the power-capped step (SM clocks held down by the GEMMs)',
      'alone': {'ms_per_launch': round(samp_alone, 4), 'achieved': round(samp_alone_gbs, 1),
                'frac': round(samp_alone_gbs / pk['hbm'], 4),
                'note': 'the same launch on the same buffers repeated 10x on its own after '
                        'the step timing'},
      'per_column_race': {
        'kernel': 'sample_fused_kernel (race_group=0: one exponential per vocab entry; the '
                  'variant whose ids are bit-comparable with the reference under injected noise)',
        'ms_per_launch': round(samp_col, 4),
        'achieved': round(samp_bytes / (samp_col * 1e-3) / 1e9, 1),
        'frac': round(samp_bytes / (samp_col * 1e-3) / 1e9 / pk['hbm'], 4)}},
    'step_roofline': {
      'flops_per_step': total_flops,
      'tensor_bound_ms': round(total_flops / (pk['tf_sustained'] * 1e12) * 1e3, 3),
      'hbm_bound_ms_sampler': round(samp_bytes / (pk['hbm'] * 1e9) * 1e3, 3),
      'frac': round((total_flops / (pk['tf_sustained'] * 1e12) * 1e3
                     + samp_bytes / (pk['hbm'] * 1e9) * 1e3) / ms_per_step, 4)},
    'kernels': kernels,
    'clocks': clk,
  })
  del run, loop, model
  torch.cuda.empty_cache()

  # ---- weak-scaling companion at N > 1: the whole config batch on every GPU
  if world > 1 and not args.batch and not args.no_weak:
    wrun = GraphRun(wl, wl.batch, device, seed)
    wrun.replay(W + 8 + settle // 4)
    wms, _ = wrun.timed(max(8, K // 2), world, device, gather=True)
    line['weak'] = {'batch_per_gpu': wl.batch, 'global_batch': wl.batch * world,
                    'ms_per_step': wms, 'steps': max(8, K // 2),
                    'value': world * wl.batch / (wl.sampling_steps * wms * 1e-3),
                    'unit': 'sequences/s'}
    del wrun
    torch.cuda.empty_cache()

  # ---- the other BASELINE.json shapes, sharded the same way
  if not args.no_configs:
    line['configs'] = {k: side_config(k, world, rank, device, K, W) for k in SIDE_CONFIGS}

  if rank != 0:
    return
  if world == 1 and not args.no_cpu_baseline:
    # ~10 s of CPU work: 16 reverse steps at batch 8 (0.6 - 0.7 s each on 16 threads)
    s_per_step, cores, timed, kind = cpu_reference_steps(8, 16, 1, budget_s=30.0)
    what = ("the reference's own Diffusion._cfg_denoise (unmodified, baseline/_ref)"
            if kind == 'reference' else 'oracle/ CPU restatement of the reference step')
    line['cpu_baseline'] = {
      'value': 8 / (wl.sampling_steps * s_per_step), 'unit': 'sequences/s', 'cores': cores,
      'kind': kind, 'ms_per_step': s_per_step * 1e3,
      'sample': (f'{what}, fp32 torch on {cores} host threads, config-3 network and shapes at '
                 f'batch 8, {timed} reverse steps timed after 1 warm-up, extrapolated linearly '
                 'in batch and steps')}
  print(json.dumps(line), flush=True)


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=200)
  ap.add_argument('--warmup', type=int, default=5)
  ap.add_argument('--impl', default='ddg_b200', choices=['ddg_b200', 'reference'])
  ap.add_argument('--config', default='c3', choices=sorted(WORKLOADS))
  ap.add_argument('--batch', type=int, default=0,
                  help='sequences per GPU (default: the config batch / N)')
  ap.add_argument('--ref-batch', type=int, default=8,
                  help='reference arm: sequences of the bounded CPU sample')
  ap.add_argument('--ref-budget-s', type=float, default=150.0)
  ap.add_argument('--no-cpu-baseline', action='store_true')
  ap.add_argument('--no-configs', action='store_true')
  ap.add_argument('--no-weak', action='store_true')
  ap.add_argument('--no-settle', action='store_true')
  args = ap.parse_args()
  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  local_rank = int(os.environ.get('LOCAL_RANK', '0'))
  if args.impl == 'reference':
    run_reference_arm(args, rank)
    return
  if world > 1:
    import ddg_b200  # noqa: F401
    from ddg_b200 import dist as ddist
    ddist.init_from_env(backend='nccl')
  run_gpu_arm(args, rank, world, local_rank)
  if world > 1:
    torch.distributed.destroy_process_group()


if __name__ == '__main__':
  main()
