#!/usr/bin/env python
"""profiles/ncu_traffic.json from an `ncu --set full` capture of bench.py's eager step (no hand editing):
DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per launch, averaged over the captured launches of each
kernel, summed over the kernels bench.py times together as `fused_forward_kernel` (stage A + stage B + fix-up) and
`fused_backward_kernel` (the column kernel; the gradient-row pass is reported separately).
usage: python scripts/make_ncu_traffic.py gpurun_out/prof_x_step.ncu-rep <workload name> <frames per launch> [out.json]"""
import collections
import csv
import json
import subprocess
import sys


def per_kernel(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index('Kernel Name')
    scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1.0, 'ms': 1e3}
    acc = collections.defaultdict(lambda: collections.defaultdict(list))
    for r in rows[2:]:
        name = r[ki].split('(')[0].replace('void ', '').replace('bevpool::', '')
        for m in ('dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__time_duration.sum'):
            i = hdr.index(m)
            acc[name][m].append(float(r[i].replace(',', '')) * scale.get(units[i], 1.0))
    return {k: {m: sum(v) / len(v) for m, v in d.items()} | {'launches': len(d['gpu__time_duration.sum'])} for k, d in acc.items()}


def main():
    path, workload, frames = sys.argv[1], sys.argv[2], int(sys.argv[3])
    out_path = sys.argv[4] if len(sys.argv) > 4 else 'profiles/ncu_traffic.json'
    k = per_kernel(path)
    tot = lambda pred: sum(v['dram__bytes_read.sum'] + v['dram__bytes_write.sum'] for n, v in k.items() if pred(n))
    res = {'source': path, 'what': 'dram__bytes_read.sum + dram__bytes_write.sum per launch (ncu --set full, cold L2 per launch)',
           'fused_forward_kernel': {'workload': workload, 'frames_per_launch': frames,
                                    'dram_bytes_per_launch': tot(lambda n: n.startswith(('frustum_reduce', 'pool_forward_share', 'pool_forward_fixup'))),
                                    'kernels': [n for n in k if n.startswith(('frustum_reduce', 'pool_forward_share', 'pool_forward_fixup'))]},
           'fused_backward_kernel': {'workload': workload, 'frames_per_launch': frames,
                                     'dram_bytes_per_launch': tot(lambda n: n.startswith('fused_backward')),
                                     'kernels': [n for n in k if n.startswith('fused_backward')]},
           'per_kernel': {n: {'dram_read_bytes': v['dram__bytes_read.sum'], 'dram_write_bytes': v['dram__bytes_write.sum'],
                              'time_us': v['gpu__time_duration.sum'], 'launches_averaged': v['launches']} for n, v in k.items()}}
    json.dump(res, open(out_path, 'w'), indent=1)
    print(json.dumps({n: res[n]['dram_bytes_per_launch'] for n in ('fused_forward_kernel', 'fused_backward_kernel')}))


if __name__ == '__main__':
    main()
