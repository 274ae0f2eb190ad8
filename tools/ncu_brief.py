#!/usr/bin/env python
"""Print the handful of ncu raw-page metrics used in profiles/ summaries."""
import csv
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'launch__grid_size',
        'launch__block_size', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed.sum']


def main(path, only=None):
    rows = list(csv.reader(open(path)))
    H, units = rows[0], rows[1]
    seen = set()
    for r in rows[2:]:
        name = r[H.index('Kernel Name')].split('(')[0]
        if 'launch__grid_size' in H:          # instantiations shared by several stages
            name += " grid=" + r[H.index('launch__grid_size')]
        if name in seen or (only and only not in name):
            continue
        seen.add(name)
        print("== " + name)
        for w in WANT:
            if w in H:
                print("  %-66s %s %s" % (w, r[H.index(w)], units[H.index(w)]))
        st = [(float(r[i]), h) for i, h in enumerate(H)
              if 'pcsamp_warps_issue_stalled' in h and 'not_issued' not in h
              and r[i] not in ('', 'n/a')]
        tot = sum(v for v, _ in st) or 1.0
        print("  stalls: " + ", ".join("%s %.0f%%" % (
            h.replace('smsp__pcsamp_warps_issue_stalled_', ''), 100 * v / tot)
            for v, h in sorted(st, reverse=True)[:6]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
