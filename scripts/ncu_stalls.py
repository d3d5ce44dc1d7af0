"""Top stall sites of one kernel of an .ncu-rep (SASS page, warp-state sampling).
python scripts/ncu_stalls.py rep [launch_index] [top_n]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
idx = sys.argv[2] if len(sys.argv) > 2 else '0'
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--launch-skip', idx, '--launch-count', '1'],
                     capture_output=True, text=True).stdout
lines = out.splitlines()
print(lines[0][:160])
rows = list(csv.reader(io.StringIO('\n'.join(lines[1:]))))
h = rows[0]
si = h.index('# Samples')
stalls = [(i, k) for i, k in enumerate(h) if k.startswith('stall_') and 'Not Issued' not in k]
data = []
for n, r in enumerate(rows[1:]):
  try:
    s = int(r[si])
  except (ValueError, IndexError):
    continue
  data.append((s, n, r))
total = sum(d[0] for d in data)
agg = {}
for s, n, r in data:
  for i, k in stalls:
    agg[k] = agg.get(k, 0) + int(r[i] or 0)
print('total samples', total)
print('by reason:', ', '.join(f'{k[6:]} {v / total:.1%}' for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for s, n, r in sorted(data, key=lambda d: -d[0])[:top]:
  reasons = sorted(((int(r[i] or 0), k[6:]) for i, k in stalls), reverse=True)[:2]
  print(f'{s / total:6.2%} #{n:5d} {r[1].strip()[:90]:90s} {reasons[0][1]}:{reasons[0][0]} {reasons[1][1]}:{reasons[1][0]}')
