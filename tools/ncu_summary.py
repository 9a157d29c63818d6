"""Print selected raw-page metrics of an .ncu-rep (run where ncu is installed; no GPU needed)."""
import csv, subprocess, sys, io
rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
pat = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct',
       'sm__throughput.avg.pct', 'sm__warps_active.avg.pct', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
       'sm__pipe_tensor_cycles_active.avg.pct', 'sm__inst_executed_pipe_fma.avg.pct', 'sm__pipe_fma_cycles_active.avg.pct',
       'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__issue_active.avg.pct', 'sm__inst_executed_pipe_lsu.avg.pct',
       'l1tex__throughput.avg.pct', 'lts__throughput.avg.pct', 'lts__t_sector_hit_rate.pct', 'smsp__average_warp_latency_issue_stalled',
       'smsp__average_warps_issue_stalled', 'smsp__inst_executed.sum', 'sm__inst_executed_pipe_alu.avg.pct', 'smsp__thread_inst_executed_per_inst_executed.ratio',
       'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__cycles_elapsed.avg ', 'sm__cycles_active.avg']
cols = [i for i, h in enumerate(hdr) if any(h.startswith(p) for p in pat)]
for i in cols:
    print(f'{hdr[i]:90s} {units[i]:14s} ' + ' | '.join(r[i] for r in rows[2:]))
