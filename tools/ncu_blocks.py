"""Summarise an ncu report here (no GPU): headline metrics + hottest SASS basic blocks. usage: ncu_blocks.py rep [topN]"""
import csv, subprocess, sys, io
from collections import Counter
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 16
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr, units, vals = rows[0], rows[1], rows[2]
for w in ['gpu__time_duration.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
          'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
          'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__grid_size', 'launch__block_size']:
    if w in hdr:
        i = hdr.index(w); print(f'{w:72s} {vals[i]:>18s} {units[i]}')
stall = [(h, float(vals[i])) for i, h in enumerate(hdr) if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio') and vals[i]]
if not stall:
    stall = [(h, float(vals[i])) for i, h in enumerate(hdr) if 'issue_stalled' in h and h.endswith('.pct') and vals[i].replace('.', '').isdigit()]
for h, v in sorted(stall, key=lambda x: -x[1])[:6]: print('  stall', h.replace('smsp__average_warps_issue_stalled_', '').replace('smsp__warp_issue_stalled_', ''), round(v, 2))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src))); hdr = rows[1]
isrc, iex, ith = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('Avg. Threads Executed')
data = [r for r in rows[2:] if len(r) > ith and r[iex].replace('.', '').isdigit()]
tot = sum(float(r[iex]) for r in data); nwarps = float(data[0][iex])
blocks = []; cur = None
for k, r in enumerate(data):
    ex = float(r[iex]); th = float(r[ith]) if r[ith] else 0
    if cur and abs(ex - cur['ex']) < 1e-9: cur['n'] += 1; cur['end'] = k; cur['ops'].append(r[isrc].split()[0])
    else: cur = {'start': k, 'end': k, 'ex': ex, 'th': th, 'n': 1, 'ops': [r[isrc].split()[0]]}; blocks.append(cur)
print('total warp inst %.4g, first-instruction count %d, sass lines %d' % (tot, nwarps, len(data)))
for b in sorted(blocks, key=lambda b: -b['ex'] * b['n'])[:top]:
    print(f"sass {b['start']:4d}-{b['end']:4d} n={b['n']:3d} exec/first={b['ex']/nwarps:8.2f} share={100*b['ex']*b['n']/tot:5.1f}% thr={b['th']:5.1f} {Counter(b['ops']).most_common(6)}")
