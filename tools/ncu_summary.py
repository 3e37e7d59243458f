"""Print the key ncu metrics of a report (raw page) -- used to write profiles/*.md."""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
KEYS = ['gpu__time_duration.sum','launch__registers_per_thread','launch__grid_size','launch__block_size','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__warps_active.avg.per_cycle_active','smsp__warps_eligible.avg.per_cycle_active','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','smsp__thread_inst_executed_per_inst_executed.ratio','smsp__thread_inst_executed_per_inst_executed.pct','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__t_sector_hit_rate.pct','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum','l1tex__t_requests_pipe_lsu_mem_local_op_st.sum','lts__throughput.avg.pct_of_peak_sustained_elapsed','lts__t_sector_hit_rate.pct','dram__bytes_read.sum','dram__bytes_write.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','sm__sass_thread_inst_executed_op_fadd_pred_on.sum','sm__sass_thread_inst_executed_op_fmul_pred_on.sum','sm__sass_thread_inst_executed_op_ffma_pred_on.sum','sm__sass_thread_inst_executed_op_dfma_pred_on.sum']
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print('##', d.get('Kernel Name'))
    for k in KEYS:
        if k in d: print(f'| {k} | {d[k]} | {units[hdr.index(k)]} |')
    for h in hdr:
        if 'average_warps_issue_stalled' in h and 'not_issued' not in h and float(d[h] or 0) > 0.1:
            print(f'| {h} | {d[h]} | |')
