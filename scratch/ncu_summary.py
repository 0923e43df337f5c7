import csv,sys,subprocess
rep=sys.argv[1]
out=subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[0]
keys=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','launch__registers_per_thread','launch__occupancy_limit','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__data_bank_conflicts_pipe_lsu_mem_shared','l1tex__data_pipe_lsu_wavefronts_mem_shared','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_','smsp__average_warp','smsp__warp_issue_stalled','launch__grid_size','launch__block_size','sm__cycles_elapsed.max','lts__t_sector_hit_rate','l1tex__t_sector_hit_rate','smsp__cycles_active.avg','sm__pipe_fma','sm__pipe_alu','smsp__inst_executed_pipe']
for r in rows[2:]:
    d=dict(zip(hdr,r))
    print('---',d.get('Kernel Name','')[:70])
    for h in hdr:
        if any(h.startswith(k) for k in keys):
            v=d[h]
            if v not in ('0','0.000000','n/a',''): print('  ',h,v)
