mkdir -p gpurun_out
R=r35
python tools/attn_prof.py > gpurun_out/${R}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum --clock-control none --csv --log-file gpurun_out/${R}_attn_launches.csv python tools/attn_prof.py > gpurun_out/${R}_ncu1.log 2>&1
grep -E "pkernel|dq_ds|rowdot" gpurun_out/${R}_attn_launches.csv | tail -20 | awk -F'","' '{print substr($5,1,40), $13, $15}'
