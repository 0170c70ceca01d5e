"""Turns an `ncu --set full` report into the committed evidence under profiles/:
  python tools/ncu_summarize.py gpurun_out/step_r02.ncu-rep profiles/ncu_full_r02_bf16_step
writes <out>.csv (one row per launch: duration, DRAM bytes, DRAM / L2 / SM / tensor-pipe utilisation, grid, registers) and
<out>_summary.json (per-kernel-name totals, the K6 network group's DRAM bytes per step = bench.py's roofline.traffic)."""
import csv
import io
import json
import subprocess
import sys

METRICS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
           'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
           'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
           'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__grid_size', 'launch__registers_per_thread']
UNIT = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1, 'ms': 1e3, 'usecond': 1, 'nsecond': 1e-3, 'msecond': 1e3}
NETWORK = ('hgemm', 'hconv_dgrad', 'splitk_finish', 'colsum', 'dqn_head_td', 'duelling_head', 'tma_gemm', 'tma_conv', 'gemm_kernel',
           'tc_gemm', 'u8_rows')


def main():
  rep, out = sys.argv[1], sys.argv[2]
  steps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
  raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv', '--metrics', ','.join(METRICS)], capture_output=True, text=True).stdout
  rows = list(csv.reader(io.StringIO(raw)))
  hdr, units = rows[0], rows[1]
  ix = {h: i for i, h in enumerate(hdr)}
  table, by_name = [], {}
  for r in rows[2:]:
    if len(r) < len(hdr):
      continue
    rec = {'id': int(r[ix['ID']]), 'kernel': r[ix['Kernel Name']]}
    for m in METRICS:
      if m not in ix:
        continue
      v = float(r[ix[m]].replace(',', '') or 0)
      u = units[ix[m]]
      rec[m] = v * UNIT.get(u, 1) if ('bytes' in m or 'time' in m) else v
    table.append(rec)
    short = rec['kernel'].split('(')[0]
    d = by_name.setdefault(short, {'launches': 0, 'us': 0.0, 'dram_bytes': 0.0})
    d['launches'] += 1
    d['us'] += rec['gpu__time_duration.sum']
    d['dram_bytes'] += rec['dram__bytes_read.sum'] + rec['dram__bytes_write.sum']
  with open(out + '.csv', 'w', newline='') as f:
    w = csv.writer(f)
    w.writerow(['id', 'kernel', 'us', 'dram_read_MB', 'dram_write_MB', 'dram_pct', 'l2_pct', 'sm_pct', 'tensor_pct', 'warps_active_pct',
                'grid', 'regs'])
    for r in table:
      w.writerow([r['id'], r['kernel'][:110], round(r['gpu__time_duration.sum'], 2), round(r['dram__bytes_read.sum'] / 1e6, 3),
                  round(r['dram__bytes_write.sum'] / 1e6, 3)] + [round(r.get(m, 0), 2) for m in METRICS[3:8]] +
                 [int(r.get('launch__grid_size', 0)), int(r.get('launch__registers_per_thread', 0))])
  total_us = sum(d['us'] for d in by_name.values())
  net = {k: v for k, v in by_name.items() if any(t in k for t in NETWORK)}
  summary = {
      'what': f'ncu --set full --clock-control none over {steps} captured learner step(s); cold-cache, serialised launches',
      'launches': len(table), 'sum_gpu_time_us': total_us,
      'by_kernel': {k: dict(v, share=v['us'] / total_us) for k, v in sorted(by_name.items(), key=lambda kv: -kv[1]['us'])},
      'network_group': {'kernels': sorted(net), 'us_per_step': sum(v['us'] for v in net.values()) / steps,
                        'share_of_step': sum(v['us'] for v in net.values()) / total_us},
      'step_group_dram_bytes': sum(v['dram_bytes'] for v in net.values()) / steps,
      'step_total_dram_bytes': sum(v['dram_bytes'] for v in by_name.values()) / steps,
  }
  json.dump(summary, open(out + '_summary.json', 'w'), indent=1)
  print(json.dumps({k: summary[k] for k in ('launches', 'sum_gpu_time_us', 'step_group_dram_bytes', 'step_total_dram_bytes')}))


main()
