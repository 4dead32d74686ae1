"""Per-kernel table of one step in an ncu launch list (csv with gpu__time_duration / dram bytes).
usage: python scripts/launch_table.py <launches.csv> [name of the kernel that opens a step]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr = rows[hi]
ki, mi, vi, ii = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('ID')
launch = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    d = launch.setdefault(r[ii], {'name': r[ki]})
    d[r[mi]] = float(r[vi].replace(',', ''))
ls = list(launch.values())
# one step = from a plan key kernel to the next; the timed loop comes first, per-stage timing loops after it
key = sys.argv[2] if len(sys.argv) > 2 else 'plan_key'          # kernel that opens a step ('vox_cell' for the LiDAR branch)
starts = [i for i, l in enumerate(ls) if key in l['name']]
a, b = (starts[3], starts[4]) if key == 'plan_key' else (starts[-2], starts[-1])   # camera: a warm-up step of the main loop
# (earlier ones are the probe / set-up steps); other probes: the last complete step (the timed loop comes last)
tot = sum(l.get('gpu__time_duration.sum', 0) for l in ls[a:b])
print(f'one step = {b - a} launches, {tot / 1e3:.1f} us (ncu: cold caches, serialised)')
for l in ls[a:b]:
    t = l.get('gpu__time_duration.sum', 0)
    mb = (l.get('dram__bytes_read.sum', 0) + l.get('dram__bytes_write.sum', 0)) / 1e6
    name = l['name'].split('(')[0].replace('void ', '').replace('bevpool::', '')[:60]
    print(f'  {t / 1e3:8.1f} us {100 * t / tot:5.1f}%  dram {mb:7.1f} MB  {mb / max(t, 1) * 1e6:6.0f} GB/s  {name}')
