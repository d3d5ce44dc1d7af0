"""Benchmark of the reverse-diffusion sampling hot path (BASELINE.json metric:
sampled sequences/sec, LM1B UDLM D-CFG, L=128, V=30522, sampling.steps=1000).

    python bench.py --gpus N --steps K --warmup W          # this repo
    python bench.py --impl reference --gpus N ...          # CPU reference arm

A bench "step" is ONE reverse-diffusion step over the whole per-GPU batch:
the DiT-small forward on the batched cond+uncond pair (2*B*L token rows) plus
the fused sampler.  Every step has identical shapes and no data-dependent work,
so   sequences/sec = B_total / (sampling.steps * mean step time);
with --steps 1000 the timed region is literally one full sample() of config 3.
Weak scaling: every rank samples its own B sequences (no data-path collective;
one NCCL all_gather of the ids at the end).
"""
from __future__ import annotations

import argparse
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

SAMPLING_STEPS = 1000          # config 3: T=1000 reverse steps per sample()
V, L, GAMMA = 30522, 128, 2.0
OVERRIDES = ['parameterization=d3pm', 'diffusion=uniform', 'time_conditioning=True',
             'model=small', f'model.length={L}', 'guidance=cfg',
             f'guidance.gamma={GAMMA}', 'guidance.condition=1',
             '+data.num_classes=2', f'sampling.steps={SAMPLING_STEPS}']
METRIC = 'sampled sequences/sec, LM1B UDLM D-CFG L=128 T=1000'


class BertTok:
  vocab_size, mask_token, mask_token_id = V, '[MASK]', 103
  bos_token_id, eos_token_id, pad_token_id = 101, 102, 0


def peaks():
  try:
    with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
      p = json.load(f)
    return dict(hbm=p['hbm_gbs'], tf_burst=p['bf16_tflops'],
                tf_sustained=p['bf16_tflops_sustained'], src='measured')
  except Exception:
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src='fallback')


class ClockSampler:
  """nvidia-smi clocks / throttle reasons sampled during the timed region."""
  Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
       'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
       'clocks_event_reasons.sw_power_cap')

  def __init__(self, index):
    self.index, self.rows, self.proc = index, [], None

  def start(self):
    try:
      self.proc = subprocess.Popen(
        ['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
         '-i', str(self.index), '-lms', '200'], stdout=subprocess.PIPE, text=True)
      self.t = threading.Thread(target=self._read, daemon=True)
      self.t.start()
    except Exception:
      self.proc = None

  def _read(self):
    for line in self.proc.stdout:
      self.rows.append([x.strip() for x in line.split(',')])

  def stop(self):
    if self.proc is None:
      return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unavailable']}
    self.proc.terminate()
    try:
      self.proc.wait(timeout=5)
    except Exception:
      self.proc.kill()
    sm = [float(r[0]) for r in self.rows if r and r[0].replace('.', '').isdigit()]
    mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace('.', '').isdigit()]
    names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
    reasons = sorted({n for r in self.rows if len(r) >= 7
                      for n, v in zip(names, r[3:7]) if v.lower().startswith('active')})
    return {'sm_mhz': statistics.median(sm) if sm else None,
            'sm_max_mhz': max(mx) if mx else None, 'reasons': reasons,
            'samples': len(sm)}


# ---------------------------------------------------------------- CPU arm
def cpu_reference_steps(batch, steps, warmup):
  """Times the CPU restatement of the reference's step (oracle/, fp32 torch on
  all host cores) on config 3's shapes with a small batch; returns (s/step,
  cores)."""
  from oracle import ref_sampler
  import ddg_b200
  cores = os.cpu_count() or 1
  torch.set_num_threads(cores)
  cfg = ddg_b200.compose(OVERRIDES)
  sd = ddg_b200.weights.random_state_dict(cfg, V, seed=1)
  oracle = ref_sampler.OracleSampler(
    sd, vocab_size=V, mask_index=103, parameterization='d3pm', diffusion='uniform',
    n_heads=cfg.model.n_heads, time_conditioning=True, num_classes=2)
  g = torch.Generator().manual_seed(0)
  xt = torch.randint(0, V, (batch, L), generator=g)
  sig, mct, mcs = ref_sampler.step_scalars(SAMPLING_STEPS, 1e-5, 0)
  cond = torch.full((batch,), 1, dtype=torch.long)
  times = []
  with torch.no_grad():
    for i in range(warmup + steps):
      U = torch.rand(batch, L, V, generator=g)
      t0 = time.perf_counter()
      xt, _, _ = oracle.step(
        xt, sig[i].expand(batch), mct[i].expand(batch)[:, None, None],
        mcs[i].expand(batch)[:, None, None], U, guidance='cfg', cond=cond, gamma=GAMMA)
      if i >= warmup:
        times.append(time.perf_counter() - t0)
  return sum(times) / len(times), cores


def run_reference_arm(args, rank):
  if rank != 0:
    return
  batch = 4
  steps, warmup = max(1, min(args.steps, 12)), max(1, min(args.warmup, 2))
  s_per_step, cores = cpu_reference_steps(batch, steps, warmup)
  value = batch / (SAMPLING_STEPS * s_per_step)
  sample = (f'oracle/ CPU restatement of the reference step (fp32 torch, {cores} threads), '
            f'config-3 shapes at batch {batch}, {steps} reverse steps timed after {warmup} warm-up')
  line = {
    'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'sequences/s',
    'n_gpus': args.gpus, 'steps': steps, 'warmup': warmup,
    'ms_per_step': s_per_step * 1e3, 'higher_is_better': True, 'scaling': 'weak',
    'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
    'config': workload_config(batch, 1),
    'cpu_baseline': {'value': value, 'unit': 'sequences/s', 'cores': cores,
                     'kind': 'port', 'sample': sample},
    'e2e': {'value': value, 'unit': 'sequences/s', 'h2d_bytes_per_step': 0,
            'd2h_bytes_per_step': 0},
    'gpu_launches': 0,
  }
  print(json.dumps(line), flush=True)


def workload_config(batch, world):
  return {
    'workload': ('BASELINE.json configs[2]: LM1B UDLM (d3pm, uniform) DiT-small, L=128, '
                 'vocab 30522, sampling.steps=1000, guidance=cfg gamma=2'),
    'step': 'one reverse-diffusion step = DiT forward on the batched cond+uncond pair + fused sampler',
    'batch_per_gpu': batch, 'global_batch': batch * world, 'seq_len': L, 'vocab': V,
    'sampling_steps': SAMPLING_STEPS, 'parallelism': f'batch-sharded x{world}',
    'cache': 'per-step working set (>= 8 GB of logits) exceeds the 126 MB L2; no flush needed',
  }


# ---------------------------------------------------------------- GPU arm
def build_model(batch, device, seed):
  import ddg_b200
  cfg = ddg_b200.compose(OVERRIDES + [f'sampling.batch_size={batch}'])
  sd = ddg_b200.weights.random_state_dict(cfg, V, seed=1)
  model = ddg_b200.Diffusion(cfg, BertTok(), device=device).load_state_dict(sd)
  model.seed = seed
  if os.environ.get('DDG_RACE_GROUP'):             # sampler A/B knob (0 | 8 | 64)
    model.race_group = int(os.environ['DDG_RACE_GROUP'])
  seg = int(os.environ.get('DDG_SEG_SEQS', '0'))   # micro-batch experiment knob
  if seg > 0:
    model.backbone.seg_seqs = seg
  return model, cfg


def ncu_traffic(batch):
  """DRAM bytes per launch of each kernel family from the committed ncu
  capture (profiles/r01_traffic.json); None when the capture was taken at
  another batch size."""
  try:
    with open(os.path.join(ROOT, 'profiles', 'r01_traffic.json')) as f:
      t = json.load(f)
    return t['bytes_per_launch'] if t.get('batch_per_gpu') == batch else None
  except Exception:
    return None


def flops_per_step(cfg, batch):
  h, nb = cfg.model.hidden_size, cfg.model.n_blocks
  rows = 2 * batch * L
  per_token = nb * (24 * h * h + 4 * L * h) + 2 * h * V
  gemm = {
    'qkv': 2 * rows * h * 3 * h * nb, 'attn_out': 2 * rows * h * h * nb,
    'fc1': 2 * rows * h * 4 * h * nb, 'fc2': 2 * rows * 4 * h * h * nb,
    'vocab': 2 * rows * h * V}
  return rows * per_token, gemm


def kernel_breakdown(model, cfg, batch):
  """One instrumented eager step: CUDA events around every launch group on
  the launching stream -> per-kernel-family device time."""
  from ddg_b200 import _lib
  lib = _lib.lib
  names = {'ddg_gemm': None, 'ddg_attention': 'attention', 'ddg_ln_modulate': 'ln_modulate',
           'ddg_sample_step': 'sample_step', 'ddg_cond_embed': 'cond', 'ddg_adaln': 'cond',
           'ddg_step_advance': 'cond'}
  records = []
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
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      e0.record()
      rc = fn(*a)
      e1.record()
      records.append((label, e0, e1))
      return rc
    return inner
  loop = model._loop
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
  return out


def sampler_alone_ms(model, loop, reps=10):
  """The step's fused-sampler launch (same buffers, same arguments) repeated on its own:
  what the kernel does when the SM clock is not held down by the GEMMs of the step."""
  from ddg_b200 import _lib
  B, L = loop.B, loop.L
  _, mct, mcs = loop.tables
  ld = model.backbone.vocab_ld
  xs = torch.empty_like(loop.xt)

  def launch():
    model._launch_sampler(
      logits_a=loop.logits.data_ptr(), ld_a=ld,
      logits_b=(loop.logits.data_ptr() + B * L * ld * 2) if loop.two else None,
      ld_b=ld if loop.two else 0, xt=loop.xt, xs=xs, mc_t=mct, mc_s=mcs, rows=B * L, seq_len=L,
      mc_seq_stride=0, mc_step_stride=1, step_ptr=loop.step, seed_ptr=loop.seed_dev,
      stats=(loop.stats, 0, B * L if loop.two else 0),
      guidance=_lib.GUIDE_CFG if loop.two else _lib.GUIDE_NONE,
      gamma=loop.gamma if loop.two else 1.0)
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
  return e0.elapsed_time(e1) / reps


def run_gpu_arm(args, rank, world, local_rank):
  import ddg_b200
  from ddg_b200 import dist as ddist
  torch.cuda.set_device(local_rank)
  device = torch.device('cuda', local_rank)
  B = args.batch
  model, cfg = build_model(B, device, ddist.rank_seed(1, rank))
  pk = peaks()

  # ---- graph loop: prepare tables for a full 1000-step run, capture one step
  cond = torch.full((B,), cfg.guidance.condition, dtype=torch.long, device=device)
  xt0 = model._sample_prior(B, L)
  sig, mct, mcs = model._step_scalars(SAMPLING_STEPS, 1e-5)
  loop = model._get_loop(B, L, 'cfg', cfg.guidance)
  loop.prepare(xt0, sig, mct, mcs, cond)
  K, W = args.steps, max(args.warmup, 3)
  loop.replay(W)                       # untimed warm-up steps
  if world > 1:
    torch.distributed.barrier()
  torch.cuda.synchronize()
  clocks = ClockSampler(local_rank)
  clocks.start()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  done = 0
  while done < K:                      # tables hold 1000 steps: wrap the counter
    room = SAMPLING_STEPS - (W + done) % SAMPLING_STEPS
    n = min(K - done, room)
    if (W + done) % SAMPLING_STEPS == 0 and done > 0:
      loop.step.zero_()
    loop.replay(n)
    done += n
  ids = ddist.gather_samples(loop.xt)  # the one collective of sample()
  e1.record()
  torch.cuda.synchronize()
  if world > 1:
    torch.distributed.barrier()
  clk = clocks.stop()
  elapsed_ms = e0.elapsed_time(e1)
  if world > 1:
    t = torch.tensor([elapsed_ms], device=device)
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    elapsed_ms = float(t.item())
  ms_per_step = elapsed_ms / K
  value = world * B / (SAMPLING_STEPS * ms_per_step * 1e-3)
  assert ids.shape[0] == world * B

  # ---- e2e: the per-step public API with HOST buffers, copies inside the timing
  Ke = max(3, min(K, 40))
  host_xt = torch.empty(B, L, dtype=torch.int64).pin_memory()
  host_xt.copy_(xt0.cpu())
  host_xs = torch.empty(B, L, dtype=torch.int64).pin_memory()
  dev = device

  # schedule scalars are constants of the run, resident on the device; the per-step INPUT is xt
  sig_d, mct_d, mcs_d = (t.to(dev, torch.float32) for t in (sig, mct, mcs))

  def e2e_step(i):
    x = host_xt.to(dev, non_blocking=True)
    xs, _, _ = model._cfg_denoise(cond, GAMMA, x, sig_d[i].expand(B), mct_d[i].expand(B)[:, None, None],
                                  mcs_d[i].expand(B)[:, None, None])
    host_xs.copy_(xs, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    host_xt.copy_(host_xs)
  for i in range(3):
    e2e_step(i)
  if world > 1:
    torch.distributed.barrier()
  torch.cuda.synchronize()
  t0 = time.perf_counter()
  for i in range(Ke):
    e2e_step(3 + i)
  torch.cuda.synchronize()
  e2e_s = (time.perf_counter() - t0) / Ke
  if world > 1:
    t = torch.tensor([e2e_s], device=device)
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    e2e_s = float(t.item())
  e2e_value = world * B / (SAMPLING_STEPS * e2e_s)

  if rank != 0:
    return
  # ---- per-kernel roofline (rank 0, after the timed regions)
  total_flops, gemm_flops = flops_per_step(cfg, B)
  kb = kernel_breakdown(model, cfg, B)
  step_kernel_ms = sum(d['ms'] for d in kb.values())
  gemm_ms = sum(d['ms'] for k, d in kb.items() if k.startswith('gemm_'))
  gemm_total_flops = sum(gemm_flops.values())
  kernels = {}
  for k, d in sorted(kb.items(), key=lambda kv: -kv[1]['ms']):
    ent = {'ms_per_step': round(d['ms'], 4), 'launches': d['launches'],
           'share': round(d['ms'] / step_kernel_ms, 4)}
    if k.startswith('gemm_'):
      fl = gemm_flops[k[5:]]
      ent['tflops'] = round(fl / (d['ms'] * 1e-3) / 1e12, 1)
      ent['frac_of_sustained_peak'] = round(ent['tflops'] / pk['tf_sustained'], 4)
    kernels[k] = ent
  rows = B * L
  samp_bytes = rows * (2 * V * 2 + 16)
  samp_ms = kb['sample_step']['ms']
  samp_gbs = samp_bytes / (samp_ms * 1e-3) / 1e9
  samp_alone_ms = sampler_alone_ms(model, loop)
  samp_alone_gbs = samp_bytes / (samp_alone_ms * 1e-3) / 1e9
  gemm_tf = gemm_total_flops / (gemm_ms * 1e-3) / 1e12
  n_launch = sum(d['launches'] for d in kb.values())
  traffic = ncu_traffic(B)
  gemm_traffic = samp_traffic = None
  if traffic:
    # per launch, averaged over the 49 GEMM launches of a step like `achieved`
    gemm_traffic = sum(traffic[k] * kb[k]['launches'] for k in kb if k.startswith('gemm_')) / sum(
      kb[k]['launches'] for k in kb if k.startswith('gemm_'))
    samp_traffic = traffic['sample_step']
  line = {
    'metric': METRIC, 'value': value, 'unit': 'sequences/s', 'n_gpus': world,
    'steps': K, 'warmup': W, 'ms_per_step': ms_per_step, 'higher_is_better': True,
    'scaling': 'weak', 'vs_baseline': None, 'dtype': 'fp16 blocks / bf16 head, fp32 accumulate',
    'data': 'synthetic (random-init DiT-small, Philox prior and noise)',
    'config': workload_config(B, world),
    'e2e': {'value': e2e_value, 'unit': 'sequences/s',
            'h2d_bytes_per_step': B * L * 8 + 3 * 4, 'd2h_bytes_per_step': B * L * 8,
            'api': 'Diffusion._cfg_denoise per step, pinned host xt in / xs out', 'steps': Ke},
    'gpu_launches': n_launch * K,
    'roofline': {
      'kernel': 'gemm_kernel (tcgen05; all DiT linears of one step)', 'bound': 'tensor',
      'achieved': round(gemm_tf, 1), 'peak': pk['tf_sustained'], 'unit': 'TFLOP/s',
      'frac': round(gemm_tf / pk['tf_sustained'], 4), 'traffic': gemm_traffic,
      'traffic_note': 'ncu dram bytes per launch, mean over the 49 GEMM launches of a step (profiles/r01_traffic.json)',
      'flops_per_launch': gemm_total_flops / 49.0, 'ms_per_launch': gemm_ms / 49.0,
      'peak_source': f"{pk['src']} sustained bf16 (kernel timed inside a long step)",
      'share_of_step': round(gemm_ms / step_kernel_ms, 4)},
    'roofline_sampler': {
      'kernel': 'sample_group_kernel (fused sampler, hierarchical race)', 'bound': 'hbm',
      'achieved': round(samp_gbs, 1), 'peak': pk['hbm'], 'unit': 'GB/s',
      'frac': round(samp_gbs / pk['hbm'], 4), 'traffic': samp_traffic,
      'algorithmic_bytes': samp_bytes, 'peak_source': pk['src'],
      'share_of_step': round(samp_ms / step_kernel_ms, 4),
      'note': 'timed inside the power-capped step (SM clocks held at ~1.2-1.3 GHz by the GEMMs)',
      'alone': {'ms_per_launch': round(samp_alone_ms, 4), 'achieved': round(samp_alone_gbs, 1),
                'frac': round(samp_alone_gbs / pk['hbm'], 4),
                'note': 'the same launch on the same buffers repeated 10x on its own after the step timing'}},
    'step_roofline': {
      'flops_per_step': total_flops,
      'tensor_bound_ms': round(total_flops / (pk['tf_sustained'] * 1e12) * 1e3, 3),
      'hbm_bound_ms_sampler': round(samp_bytes / (pk['hbm'] * 1e9) * 1e3, 3),
      'frac': round((total_flops / (pk['tf_sustained'] * 1e12) * 1e3
                     + samp_bytes / (pk['hbm'] * 1e9) * 1e3) / ms_per_step, 4)},
    'kernels': kernels,
    'clocks': clk,
  }
  if world == 1 and not args.no_cpu_baseline:
    s_per_step, cores = cpu_reference_steps(4, 3, 1)
    line['cpu_baseline'] = {
      'value': 4 / (SAMPLING_STEPS * s_per_step), 'unit': 'sequences/s', 'cores': cores,
      'kind': 'port',
      'sample': ('oracle/ CPU restatement of the reference step, fp32 torch, config-3 shapes at '
                 'batch 4, 3 reverse steps timed after 1 warm-up, extrapolated linearly in steps')}
  print(json.dumps(line), flush=True)


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=200)
  ap.add_argument('--warmup', type=int, default=5)
  ap.add_argument('--impl', default='ddg_b200', choices=['ddg_b200', 'reference'])
  ap.add_argument('--batch', type=int, default=512, help='sequences per GPU')
  ap.add_argument('--no-cpu-baseline', action='store_true')
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
