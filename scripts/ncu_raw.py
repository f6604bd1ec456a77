#!/usr/bin/env python
"""Key raw metrics of every kernel in an .ncu-rep.  usage: ncu_raw.py report.ncu-rep"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__cycles_active.avg']
want += [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio')]
for r in rows[2:]:
    print('==', r[hdr.index('Kernel Name')][:90])
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            try:
                if w.startswith('smsp__average_warps_issue_stalled') and float(r[i]) < 0.3:
                    continue
            except ValueError:
                pass
            print('   %-88s %s %s' % (w, r[i], rows[1][i]))
