import csv,sys,subprocess
rep=sys.argv[1]
out=subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
h=rows[0]
want=['Kernel Name','gpu__time_duration.sum','launch__grid_size','launch__block_size','dram__bytes_read.sum','dram__bytes_write.sum','dram__bytes_read.sum.per_second','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','lts__t_sector_hit_rate.pct','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct','l1tex__t_sector_hit_rate.pct','sm__throughput.avg.pct_of_peak_sustained_elapsed','launch__registers_per_thread','launch__waves_per_multiprocessor','lts__throughput.avg.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_tensor.sum','smsp__inst_executed.sum','launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']
for r in rows[2:]:
    for i,n in enumerate(h):
        if n in want or ('stalled' in n and 'ratio' in n and float(r[i] or 0)>0.8): print(f"{n:80s} {rows[1][i]:12s} {r[i][:100]}")
    print('---')
