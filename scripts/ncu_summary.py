"""Key metrics of every kernel in an .ncu-rep: python scripts/ncu_summary.py rep [csv_out]"""
import csv
import io
import subprocess
import sys

M = ['gpu__time_duration.sum', 'sm__cycles_active.avg', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
     'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
     'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct',
     'sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
     'smsp__inst_executed.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
     'launch__grid_size', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum']
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv', '--metrics', ','.join(M)],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h, units = rows[0], rows[1]
cols = [i for i, k in enumerate(h) if k in M]
table = []
for r in rows[2:]:
  name = r[h.index('Kernel Name')]
  name = name.replace('void ddg::<unnamed>::', '').split('(')[0]
  d = {h[i]: (r[i], units[i]) for i in cols}
  dur = float(d['gpu__time_duration.sum'][0].replace(',', ''))
  u = d['gpu__time_duration.sum'][1]
  dur_ms = dur / 1e6 if u in ('ns', 'nsecond') else dur / 1e3 if u in ('us', 'usecond') else dur
  cyc = float(d['sm__cycles_active.avg'][0].replace(',', ''))
  print(f"{name[:44]:44s} {dur_ms:8.4f} ms  clk {cyc / dur_ms / 1e6:5.2f} GHz  tensor {d[M[2]][0][:5]:>5s}%  "
        f"sm {d[M[3]][0][:5]:>5s}%  dram {d[M[4]][0][:5]:>5s}%  rd {d[M[5]][0][:7]} {d[M[5]][1]}  wr {d[M[6]][0][:7]} {d[M[6]][1]}  "
        f"xu {d[M[8]][0][:5]}%  issue {d[M[9]][0][:5]}%  regs {d[M[12]][0]}")
  table.append([sys.argv[1].split('/')[-1], name] + [d[k][0] + ' ' + d[k][1] for k in M])
if len(sys.argv) > 2:
  with open(sys.argv[2], 'a') as f:
    w = csv.writer(f)
    w.writerow(['report', 'kernel'] + M)
    w.writerows(table)
