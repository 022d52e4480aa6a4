import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[0]
want=['Kernel Name','launch__grid_size','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','sm__cycles_elapsed.max','smsp__cycles_active.avg','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts.sum','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_active']
want+= [h for h in hdr if 'issue_stalled' in h and 'per_issue_active' in h]
idx=[hdr.index(w) for w in want if w in hdr]
for r in rows[2:]:
    print('=====')
    for i in idx:
        v=r[i]
        try:
            f=float(v.replace(',',''))
            if 'stalled' in hdr[i] and f<0.15: continue
            v='%.4g'%f
        except: pass
        print('  ',hdr[i].replace('smsp__average_warps_issue_stalled_','stall:').replace('_per_issue_active.ratio',''), v)
