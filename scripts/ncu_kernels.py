#!/usr/bin/env python
"""Per-kernel table from one .ncu-rep holding many launches (ncu --set full).
usage: python scripts/ncu_kernels.py gpurun_out/prof_x.ncu-rep"""
import csv
import subprocess
import sys

W = [('gpu__time_duration.sum', 'time'), ('dram__bytes_read.sum', 'dram_rd'), ('dram__bytes_write.sum', 'dram_wr'),
     ('lts__t_sector_hit_rate.pct', 'L2hit%'), ('l1tex__t_sector_hit_rate.pct', 'L1hit%'),
     ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occ%'), ('launch__registers_per_thread', 'regs'),
     ('l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'L1tp%'),
     ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'L2tp%'),
     ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'DRAMtp%'),
     ('smsp__inst_executed.sum', 'warp_inst'), ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue%'),
     ('lts__t_sectors_srcunit_tex_op_read.sum', 'L2rd_sectors'), ('launch__grid_size', 'grid'),
     ('launch__block_size', 'block')]


def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index('Kernel Name')
    for r in rows[2:]:
        print(r[ki][:90])
        parts = []
        for w, short in W:
            if w in hdr:
                i = hdr.index(w)
                v = r[i]
                try:
                    v = f'{float(v):.4g}'
                except ValueError:
                    pass
                parts.append(f'{short}={v}{units[i] if units[i] in ("us", "Mbyte", "Gbyte", "Kbyte", "ms") else ""}')
        print('    ' + '  '.join(parts))
        stall = [h for h in hdr if 'smsp__average_warps_issue_stalled' in h and 'per_issue_active' in h and 'not_issued' not in h]
        vals = sorted(((float(r[hdr.index(h)]), h) for h in stall), reverse=True)[:5]
        print('    stalls: ' + ', '.join(
            f"{h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}={v:.2f}" for v, h in vals))


if __name__ == '__main__':
    main(sys.argv[1])
