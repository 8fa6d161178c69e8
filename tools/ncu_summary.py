"""Print the metrics we track from an .ncu-rep (run here, no GPU needed): python tools/ncu_summary.py file.ncu-rep"""
import csv, subprocess, sys
KEYS = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps',
        'smsp__inst_executed.sum', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'lts__t_bytes.sum', 'sm__cycles_elapsed.max', 'launch__occupancy_limit_blocks', 'sm__maximum_warps_per_active_cycle_pct',
        'smsp__inst_executed_per_warp', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic']
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(k, '=', r[i], units[i])
    for i, h in enumerate(hdr):   # local-memory traffic and cache hit rates (the gate interpreter keeps temporaries in local memory)
        if ('mem_local' in h and h.endswith('.sum')) or h.endswith('hit_rate.pct') or h.startswith('smsp__inst_executed_op_local'):
            if r[i] not in ('', '0', 'n/a'):
                print(h, '=', r[i], units[i])
    stalls = []
    for i, h in enumerate(hdr):
        if 'issue_stalled' in h and h.endswith('per_warp_active.pct') and 'not_issued' not in h:
            v = float(r[i])
            if v >= 3:
                stalls.append((v, h.replace('smsp__warp_issue_stalled_', '').replace('_per_warp_active.pct', '')))
    print('stalls(% of warp-active cycles):', ', '.join('%s=%.0f' % (n, v) for v, n in sorted(stalls, reverse=True)))
    per_issue = []
    for i, h in enumerate(hdr):
        if 'issue_stalled' in h and h.endswith('_per_issue_active.ratio'):
            try:
                v = float(r[i])
            except ValueError:
                continue
            if v >= 0.3:
                per_issue.append((v, h.split('issue_stalled_')[1].replace('_per_issue_active.ratio', '')))
    print('stalls (warps per issue):', ', '.join('%s %.2f' % (n, v) for v, n in sorted(per_issue, reverse=True)))
    print('---')
