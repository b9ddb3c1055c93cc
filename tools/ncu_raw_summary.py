"""Summarise an `ncu --page raw --csv` export: per kernel time, occupancy, pipe utilisation, DRAM bytes, stall mix."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
def g(r, k):
    return r[hdr.index(k)] if k in hdr else 'n/a'
for r in rows[2:]:
    print('-----', g(r, 'Kernel Name')[:70], 'grid', g(r, 'launch__grid_size'), 'block', g(r, 'launch__block_size'))
    for k in ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem',
              'launch__occupancy_limit_registers', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
              'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
              'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
              'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
              'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
              'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__cycles_elapsed.max']:
        if k in hdr:
            print('  %-62s %s %s' % (k, g(r, k), units[hdr.index(k)]))
    st = [(float(r[i]), h) for i, h in enumerate(hdr) if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio')]
    print('  stalls (warps per issue):', ', '.join('%s %.2f' % (h.split('stalled_')[1].split('_per_')[0], v) for v, h in sorted(st, reverse=True)[:7]))
