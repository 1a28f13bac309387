"""Summarise an `ncu --page raw --csv` export: one block of key metrics per captured kernel."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sectors_op_atom.sum', 'lts__t_sectors_op_red.sum',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct']
want += [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio')]
idx = [hdr.index(w) for w in want if w in hdr]
for r in rows[2:]:
    print('----')
    for i in idx:
        v = r[i]
        if hdr[i].startswith('smsp__average_warps_issue_stalled'):
            try:
                if float(v) < 0.3: continue
            except ValueError: pass
        print(f"  {hdr[i]} = {v} {units[i]}")
