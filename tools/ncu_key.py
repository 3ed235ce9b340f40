"""Developer tool: key metrics of the first kernel in `ncu -i X.ncu-rep --page raw --csv` output (argv[1])."""
import csv
import sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, r = rows[0], rows[2]
want = ['Kernel Name', 'gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'memory_l1_wavefronts_shared_ideal', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'sm__cycles_elapsed.max', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed']
for i, h in enumerate(hdr):
    if h in want or ('issue_stalled' in h and 'per_issue_active' in h and float(r[i] or 0) > 0.05):
        print(h, r[i])
