"""Summarise `ncu -i X.ncu-rep --page raw --csv` into one markdown bullet per captured launch (key metrics only).
Usage: python tools/summarize_ncu_raw.py raw.csv > summary.md"""
import csv
import sys

KEYS = ['Kernel Name', 'launch__grid_size', 'launch__block_size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_uniform.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active']


def main():
    rd = csv.reader(open(sys.argv[1], newline=''))
    head = next(rd)
    units = next(rd)
    idx = {h: i for i, h in enumerate(head)}
    for row in rd:
        if len(row) < len(head):
            continue
        parts = []
        for k in KEYS:
            if k in idx:
                u = units[idx[k]]
                parts.append('%s = %s%s' % (k, row[idx[k]], (' ' + u) if u else ''))
        print('- ' + '; '.join(parts))


if __name__ == '__main__':
    main()
