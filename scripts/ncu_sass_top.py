"""Top stall sites of one kernel from `ncu --page source --print-source sass --csv`.
usage: ncu -i rep --page source --csv --print-source sass --launch-skip K --launch-count 1 > f.csv
       python scripts/ncu_sass_top.py f.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
h = rows[hi]
rows = [r for r in rows[hi + 1:] if len(r) == len(h) and r[0] != 'Address']
ia, isrc, ismp, iex = (h.index(k) for k in ('Address', 'Source', '# Samples', 'Instructions Executed'))
stalls = [i for i, k in enumerate(h) if k.startswith('stall_') and 'Not Issued' not in k]
tot = sum(int(r[ismp] or 0) for r in rows)
print('total samples', tot, '| instructions', len(rows), '| warp-insts executed',
      sum(int(r[iex] or 0) for r in rows))
agg = {h[i]: sum(int(r[i] or 0) for r in rows) for i in stalls}
for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]:
  print(f'  {k:28s} {v:8d} {v / max(tot, 1) * 100:5.1f}%')
for r in sorted(rows, key=lambda r: -int(r[ismp] or 0))[:n]:
  st = sorted(((int(r[i] or 0), h[i][6:]) for i in stalls), reverse=True)[:2]
  print(f'{r[ia][-5:]} {int(r[ismp] or 0):6d} {int(r[iex] or 0):9d}  {r[isrc][:64]:64s} {st}')
