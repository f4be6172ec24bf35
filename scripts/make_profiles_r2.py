"""Turns the outputs of scripts/gpu_final_r2.sh (gpurun_out/final_*) into the committed profiles/r2_* files."""
import collections
import csv
import json
import os
import re
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, 'gpurun_out'), os.path.join(ROOT, 'profiles')


def last_json_line(path):
  return [l for l in open(path) if l.startswith('{')][-1]


for cfg in ('d0', 'd4', 'd7x', 'v2s'):
  src = os.path.join(G, 'final_bench_%s.log' % cfg)
  if os.path.exists(src):
    open(os.path.join(P, 'r2_bench_line_%s.json' % cfg), 'w').write(last_json_line(src))
  src = os.path.join(G, 'final_ops_%s.json' % cfg)
  if os.path.exists(src):
    shutil.copy(src, os.path.join(P, 'r2_ops_cuda_events_%s.json' % cfg))
if os.path.exists(os.path.join(G, 'final_ref.log')):
  open(os.path.join(P, 'r2_bench_reference_arm.json'), 'w').write(last_json_line(os.path.join(G, 'final_ref.log')))
for name in ('final_segments_d0.json', 'parity_bench_shapes.json'):
  if os.path.exists(os.path.join(G, name)):
    shutil.copy(os.path.join(G, name), os.path.join(P, 'r2_' + name.replace('final_', '')))
shutil.copy(os.path.join(G, 'final_launches.csv'), os.path.join(P, 'r2_launches_ncu.csv'))
shutil.copy(os.path.join(G, 'final_full_raw.csv'), os.path.join(P, 'r2_ncu_full_raw.csv'))

# ---- launch list -> one forward (stem launch to the next stem launch) per kernel family ----
rows = list(csv.reader(open(os.path.join(G, 'final_launches.csv'))))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
h = rows[hi]
ci = {n: i for i, n in enumerate(h)}
launches = collections.OrderedDict()
for r in rows[hi + 1:]:
  if len(r) < len(h):
    continue
  d = launches.setdefault(int(r[ci['ID']]), {'name': r[ci['Kernel Name']]})
  d[r[ci['Metric Name']]] = (float(r[ci['Metric Value']].replace(',', '')), r[ci['Metric Unit']])
L = list(launches.values())
stems = [i for i, l in enumerate(L) if 'stem_tc_kernel' in l['name'] or 'stem_kernel' in l['name']]
fwd = L[stems[0]:stems[1]]


def fam(n):
  n = re.sub(r'\b(void|edet|pwtc|sepc|mbf|convtc|pcn|topk|dwt|stemtc)::|void ', '', n)
  m = re.search(r'(\w+)(<|\()', n)
  return m.group(1) if m else n[:30]


agg = collections.OrderedDict()
for l in fwd:
  d = agg.setdefault(fam(l['name']), {'launches': 0, 'time_us': 0.0, 'dram_read_MB': 0.0, 'dram_write_MB': 0.0})
  d['launches'] += 1
  t, u = l['gpu__time_duration.sum']
  d['time_us'] += t / 1000 if u in ('ns', 'nsecond') else t
  for key, out in (('dram__bytes_read.sum', 'dram_read_MB'), ('dram__bytes_write.sum', 'dram_write_MB')):
    v, u = l[key]
    d[out] += v * {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3}[u]
tot = sum(d['time_us'] for d in agg.values())
json.dump({'source': 'ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum '
                     '--clock-control none on `python bench.py --steps 2 --warmup 1 --no-cpu-baseline` '
                     '(one forward = stem launch to the next stem launch; per-launch times are cold-cache '
                     'and serialised)', 'per_forward': agg, 'sum_us': tot, 'launches': len(fwd)},
          open(os.path.join(P, 'r2_traffic_d0.json'), 'w'), indent=1)
for f, d in sorted(agg.items(), key=lambda x: -x[1]['time_us']):
  print('%-28s n=%3d %8.1f us (%4.1f%%)  rd %7.1f MB wr %7.1f MB' % (
      f, d['launches'], d['time_us'], 100 * d['time_us'] / tot, d['dram_read_MB'], d['dram_write_MB']))
print('total us', tot, 'launches', len(fwd))

# ---- ncu --set full summary ----
rows = list(csv.reader(open(os.path.join(G, 'final_full_raw.csv'))))
h, units = rows[0], rows[1]
ci = {n: i for i, n in enumerate(h)}
want = [('gpu__time_duration.sum', 'time'), ('dram__bytes_read.sum', 'dram rd'), ('dram__bytes_write.sum', 'dram wr'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram %peak'),
        ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm %peak'),
        ('sm__inst_issued.avg.pct_of_peak_sustained_active', 'issue active %'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps active %'),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor pipe %'),
        ('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'xu pipe %'),
        ('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'fma pipe %'),
        ('sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'lsu pipe %'),
        ('launch__registers_per_thread', 'regs'), ('launch__grid_size', 'grid'), ('launch__block_size', 'block'),
        ('launch__occupancy_limit_registers', 'occ limit regs'), ('launch__occupancy_limit_shared_mem', 'occ limit smem'),
        ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'stall long_scoreboard'),
        ('smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'stall wait'),
        ('smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'stall barrier'),
        ('smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'stall math_pipe'),
        ('smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio', 'stall mio'),
        ('smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'stall short_scoreboard'),
        ('smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'stall not_selected')]
labels = ['stem 640->320x32', 'blocks_1 fused expand+dw (mbconv_front)', 'head tower layer L3 (sepconv_direct)',
          'BiFPN node L3 (fuse_dw)', 'blocks_0 dw k3s1 320x32', 'blocks_1 dw k3s2 320x96', 'blocks_4 dw k5s1 80x240',
          'blocks_9 dw k5s1 40x672', 'blocks_1 expand 16->96', 'blocks_2 project 144->24 (+res, SE weights)',
          'blocks_6 expand 80->480', 'blocks_12 project 1152->192', 'class-predict L3 64->810 (logits stored: forward())',
          'class head + arg-max L3 64->9x96 (detect path)']
out = ['# ncu --set full summaries (round 2, final build)', '',
       '`ncu --set full --clock-control none -k regex:"stem_tc_kernel|fuse_dw_kernel|pointwise_tc|dw_tile_kernel|'
       'depthwise_kernel|sepconv|mbconv_front" python scripts/profile_kernels.py 1`',
       '(representative D0 @ 640x640 batch-32 launches, one launch each, in the order of '
       'scripts/profile_kernels.py; scripts/gpu_final_r2.sh).  Raw page: `r2_ncu_full_raw.csv`.', '']
for k, r in enumerate(rows[2:]):
  if len(r) < len(h):
    continue
  out.append('## %d. %s' % (k + 1, labels[k] if k < len(labels) else ''))
  out.append('`%s`' % r[ci['Kernel Name']][:110])
  out += ['', '| metric | value |', '|---|---|']
  for m, lab in want:
    if m in ci:
      out.append('| %s | %s %s |' % (lab, r[ci[m]], units[ci[m]]))
  out.append('')
open(os.path.join(P, 'r2_ncu_full_summary.md'), 'w').write('\n'.join(out))
