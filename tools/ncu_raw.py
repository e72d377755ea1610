"""Print the key metrics of an `ncu --page raw --csv` dump. usage: ncu_raw.py file.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
keys = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed', 'sm__cycles_elapsed.avg', 'smsp__cycles_active.avg',
        'l1tex__average_t_sectors_per_request_pipe_lsu_mem_global_op_ld.ratio', 'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_op_read.sum', 'lts__t_sectors_op_write.sum']
for k in keys:
    if k in hdr:
        i = hdr.index(k)
        print('%-78s %-10s %s' % (k, units[i], data[0][i]))
# sectors per request of the LSU global accesses (derived: the ratio metric itself is not part of --set full); the frames
# themselves move through TMA (no LSU requests), so these are the few per-lane accesses that remain (flow / grad_out loads,
# grid stores, the per-pixel fallback path)
col = {h: i for i, h in enumerate(hdr)}
for op in ('ld', 'st', 'red'):
    rq, sc = 'l1tex__t_requests_pipe_lsu_mem_global_op_%s.sum' % op, 'l1tex__t_sectors_pipe_lsu_mem_global_op_%s.sum' % op
    if rq in col and sc in col:
        try:
            r, t = float(data[0][col[rq]]), float(data[0][col[sc]])
            if r > 0:
                print('%-78s %-10s %.2f  (%d requests, %d sectors)' % ('sectors per request, LSU global %s (derived)' % op, 'ratio', t / r, r, t))
        except ValueError:
            pass
print('--- warp stall breakdown (per warp active, %)')
st = []
for i, h in enumerate(hdr):
    if h.startswith('smsp__average_warp') and h.endswith('_per_issue_active.ratio') or ('warp_issue_stalled' in h and h.endswith('per_warp_active.pct')):
        try:
            st.append((float(data[0][i]), h))
        except ValueError:
            pass
for v, h in sorted(st, reverse=True)[:12]:
    print('   %8.2f  %s' % (v, h))
