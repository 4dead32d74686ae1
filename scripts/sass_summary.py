#!/usr/bin/env python
"""SASS opcode summary per kernel of libbevpool_sm100.so (cuobjdump -sass): the mnemonics that show what the kernels
are made of -- UTMALDG / UTMASTG (TMA tensor-map load / store), UBLKCP (1-D TMA bulk copy), SYNCS (mbarrier), LDGSTS
(cp.async), FFMA2 (packed fp32 FMA), ATOM / RED / ATOMS (atomics), MATCH / VOTE, LDG / STG widths.
usage: python scripts/sass_summary.py [path/to/lib.so] > profiles/<tag>_sass_summary.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else 'mm_training_b200/libbevpool_sm100.so'
out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
kern, counts = None, collections.OrderedDict()
WATCH = ['UTMALDG', 'UTMASTG', 'UBLKCP', 'SYNCS', 'LDGSTS', 'FFMA2', 'FFMA', 'FMUL', 'FADD', 'ATOMG', 'ATOM', 'RED', 'ATOMS', 'MATCH',
         'VOTE', 'SHFL', 'LDG', 'STG', 'LDS', 'STS', 'BAR', 'MUFU', 'ACQBULK', 'UTMACMDFLUSH']
for ln in out.splitlines():
    m = re.match(r'\s*Function : (\S+)', ln)
    if m:
        kern = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip().split('(')[0]
        kern = kern.replace('void ', '').replace('bevpool::', '')
        counts[kern] = collections.Counter()
        continue
    m = re.match(r'\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)', ln)
    if m and kern:
        op, mods = m.group(1), m.group(2)
        counts[kern]['_total'] += 1
        if op in WATCH:
            counts[kern][op] += 1
            if op in ('LDG', 'STG', 'LDS', 'STS') and '.128' in mods:
                counts[kern][op + '.128'] += 1
print(f'# {lib}: SASS opcode counts per kernel (static instruction counts, sm_100a)')
for k, c in counts.items():
    items = ' '.join(f'{o}={n}' for o, n in sorted(c.items()) if o != '_total')
    print(f'{k}\n    instrs={c["_total"]} {items}')
tot = collections.Counter()
for c in counts.values():
    tot.update(c)
print('TOTAL ' + ' '.join(f'{o}={n}' for o, n in sorted(tot.items())))
