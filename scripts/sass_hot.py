"""Per-instruction view of one kernel of an ncu capture: where the warp instructions and the stall samples go.
usage: python scripts/sass_hot.py <report.ncu-rep> <kernel regex> [bucket]"""
import csv, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
bucket = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + rx, '--print-source', 'sass'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
# several launches of the kernel may be in the report: keep the first
hdr = next(i for i, r in enumerate(rows) if 'Source' in r and 'Instructions Executed' in r)
h = rows[hdr]
ia, ie, isamp = h.index('Source'), h.index('Instructions Executed'), h.index('# Samples')
data = []
for r in rows[hdr + 1:]:
    if len(r) <= ie or not r[ie].isdigit():
        break
    data.append((r[ia].strip(), int(r[ie]), int(r[isamp])))
tot, ts = sum(d[1] for d in data), sum(d[2] for d in data)
print('warp instructions', tot, 'samples', ts, 'sass lines', len(data))
for i in range(0, len(data), bucket):
    e = sum(d[1] for d in data[i:i + bucket]); sm = sum(d[2] for d in data[i:i + bucket])
    print(f'{i:5d} {100 * e / tot:5.1f}% inst {100 * sm / ts:5.1f}% samples  {data[i][0][:50]}')
print()
for i in sorted(sorted(range(len(data)), key=lambda i: -data[i][2])[:30]):
    print(i, data[i][1], data[i][2], data[i][0][:90])
