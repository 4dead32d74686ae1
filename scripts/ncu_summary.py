#!/usr/bin/env python
"""Summarise .ncu-rep files: key raw metrics + opcode mix from the source page.
usage: python scripts/ncu_summary.py gpurun_out/prof_x.ncu-rep [...]"""
import collections
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'lts__t_sectors_srcunit_tex_op_read.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'launch__grid_size', 'launch__block_size',
        'launch__waves_per_multiprocessor', 'sm__maximum_warps_per_active_cycle_pct',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__cycles_active.avg']


def raw(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    last = rows[-1]
    print('kernel:', last[hdr.index('Kernel Name')][:100])
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f'  {w:70s} {rows[1][i]:>12s} {last[i]}')
    stall = [h for h in hdr if 'smsp__average_warps_issue_stalled' in h and 'per_issue_active' in h and 'not_issued' not in h]
    vals = sorted(((float(last[hdr.index(h)]), h) for h in stall), reverse=True)[:7]
    print('  stalls (warps per issue-active):',
          ', '.join(f"{h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}={v:.2f}" for v, h in vals))


def source(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    his = [i for i, r in enumerate(rows) if 'Source' in r and 'Instructions Executed' in r]
    if not his:
        return
    hi = his[-1]
    hdr = rows[hi]
    si, ii, st = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('Warp Stall Sampling (All Samples)')
    ops, stalls, tot = collections.Counter(), collections.Counter(), 0
    for r in rows[hi + 1:]:
        if len(r) <= ii or not r[0].startswith('0x'):
            continue
        try:
            n, s = int(r[ii]), int(r[st])
        except ValueError:
            continue
        toks = r[si].split()
        op = (toks[1] if toks[0].startswith('@') else toks[0]).split('.')[0]
        ops[op] += n
        stalls[op] += s
        tot += n
    print(f'  opcode mix (last launch, {tot / 1e6:.1f}M warp instructions):')
    for op, n in ops.most_common(14):
        print(f'    {op:8s} {n / 1e6:8.2f}M {100 * n / tot:5.1f}%  stall_samples={stalls[op]}')


for p in sys.argv[1:]:
    if not p.endswith('.ncu-rep'):
        continue
    print('==', p)
    raw(p)
    source(p)


def hot(path, top=18):
    """top SASS instructions by stall samples"""
    out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    his = [i for i, r in enumerate(rows) if 'Source' in r and 'Instructions Executed' in r]
    hi = his[-1]
    hdr = rows[hi]
    si, ii, st = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('Warp Stall Sampling (All Samples)')
    body = [r for r in rows[hi + 1:] if len(r) > ii and r[0].startswith('0x')]
    tot = sum(int(r[st]) for r in body)
    idx = sorted(range(len(body)), key=lambda i: -int(body[i][st]))[:top]
    print(f'  hottest SASS by stall samples (total {tot}):')
    for i in sorted(idx):
        r = body[i]
        print(f'    #{i:4d} {100 * int(r[st]) / tot:5.1f}%  exec={int(r[ii]):9d}  {r[si].strip()[:90]}')


if __name__ == '__main__' and '--hot' in sys.argv:
    for p in sys.argv[1:]:
        if p.endswith('.ncu-rep'):
            hot(p)


def loop(path, frac=0.5):
    """SASS of the hot region: every instruction executed at least `frac` x the most-executed count"""
    out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    his = [i for i, r in enumerate(rows) if 'Source' in r and 'Instructions Executed' in r]
    hi = his[-1]
    hdr = rows[hi]
    si, ii, st = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('Warp Stall Sampling (All Samples)')
    body = [r for r in rows[hi + 1:] if len(r) > ii and r[0].startswith('0x')]
    mx = max(int(r[ii]) for r in body)
    tot = sum(int(r[st]) for r in body)
    for i, r in enumerate(body):
        if int(r[ii]) >= frac * mx:
            print(f'    #{i:4d} exec={int(r[ii]):9d} stall={100 * int(r[st]) / tot:4.1f}%  {r[si].strip()[:100]}')


if __name__ == '__main__' and '--loop' in sys.argv:
    for p in sys.argv[1:]:
        if p.endswith('.ncu-rep'):
            loop(p)
