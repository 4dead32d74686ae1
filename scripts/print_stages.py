"""Reads one bench.py JSON line on stdin and prints the headline + per-stage medians (us)."""
import json
import sys

d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(d['value']), 'frames/s', round(d['ms_per_step'] * 1e3, 1), 'us/step',
      {k: round(v['median'] * 1e3, 1) for k, v in d['stages_ms'].items()})
for k in ('channels_last_step', 'sustained', 'e2e', 'dropin_op', 'depth_labels', 'lidar', 'ref_cuda', 'cpu_baseline'):
    v = d.get(k)
    if isinstance(v, dict):
        print(k, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in v.items() if isinstance(b, (int, float, str)) and a not in ('what', 'api', 'parity', 'sample')})
hp = d.get('half_precision')
if isinstance(hp, dict):
    print('half_precision', json.dumps(hp)[:600])
for r in d.get('sweep') or []:
    if isinstance(r, dict):
        print('sweep', r.get('workload'), r.get('frames_per_step'), round(r.get('ms_per_step', 0) * 1e3, 1), 'us', round(r.get('frames_per_s', 0)), 'f/s', round(r.get('frac_of_hbm_peak', 0), 3),
              'ref_cuda', round(r['ref_cuda_frames_per_s']) if 'ref_cuda_frames_per_s' in r else '-', 'cpu', round(r['cpu_port_frames_per_s'], 1) if 'cpu_port_frames_per_s' in r else '-')
