"""Print the handful of ncu raw-page metrics we track for a kernel (first kernel row of `ncu --page raw --csv`)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
vals = rows[2] if len(rows) > 2 else rows[1]
d = dict(zip(hdr, vals))
keys = ['gpu__time_duration.sum', 'sm__inst_executed.sum', 'sm__inst_executed.avg.per_cycle_active', 'smsp__issue_active.avg.pct', 'dram__bytes_read.sum',
        'dram__bytes_write.sum', 'launch__registers_per_thread', 'smsp__warps_active.avg.per_cycle_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'sm__cycles_active.avg', 'lts__t_sector_hit_rate.pct', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed_pipe_fma', 'smsp__inst_executed_pipe_lsu', 'sm__throughput.avg.pct', 'gpu__compute_memory_throughput']
for k in keys:
    for h in hdr:
        if h.startswith(k):
            print(h, d[h])
print('--- stalls per issue')
for h in hdr:
    if 'issue_stalled' in h and h.endswith('per_issue_active.ratio'):
        try:
            v = float(d[h])
        except ValueError:
            continue
        if v > 0.04:
            print('  ', h.replace('smsp__average_warps_issue_stalled_', '').replace('smsp__average_warp_latency_issue_stalled_', '').replace('_per_issue_active.ratio', ''), round(v, 3))
