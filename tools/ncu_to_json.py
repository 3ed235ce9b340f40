"""Developer tool: key metrics of every kernel in `ncu -i X.ncu-rep --page raw --csv` output (argv[1]) as JSON (stdout)."""
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__inst_executed.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'launch__shared_mem_per_block_dynamic', 'launch__grid_size', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']
out = []
for r in rows[2:]:
    if len(r) != len(hdr):
        continue
    d = {}
    for i, h in enumerate(hdr):
        if h in want or ('issue_stalled' in h and 'per_issue_active' in h):
            try:
                d[h] = float(r[i].replace(',', ''))
            except ValueError:
                d[h] = r[i]
    d = {k: v for k, v in d.items() if not ('issue_stalled' in k and isinstance(v, float) and v < 0.05)}
    out.append(d)
json.dump(out, sys.stdout, indent=1)
