"""Reduce an `ncu -i X.ncu-rep --page raw --csv` dump to the columns the roofline discussion uses.
usage: python tools/summarize_ncu.py raw.csv out.csv"""
import csv, sys
PREFIXES = ('gpu__time_duration', 'dram__bytes', 'dram__throughput', 'gpu__dram_throughput', 'sm__warps_active',
            'launch__registers', 'launch__occupancy', 'launch__shared', 'launch__grid_size', 'launch__block_size',
            'sm__throughput', 'l1tex__throughput', 'lts__throughput', 'smsp__issue_active', 'sm__inst_executed_pipe_fp64',
            'smsp__inst_executed.sum', 'smsp__average_warps_issue_stalled', 'l1tex__data_bank_conflicts', 'sm__pipe_fp64',
            'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__t_sectors_pipe_lsu_mem_global',
            'l1tex__t_requests_pipe_lsu_mem_global', 'smsp__thread_inst_executed_per_inst_executed')
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
keep = [i for i, h in enumerate(hdr) if h in ('ID', 'Kernel Name', 'Block Size', 'Grid Size') or any(h.startswith(p) for p in PREFIXES)]
with open(sys.argv[2], 'w', newline='') as f:
    w = csv.writer(f)
    for r in rows:
        w.writerow([r[i] for i in keep])
print(len(rows) - 2, "kernels,", len(keep), "columns")
