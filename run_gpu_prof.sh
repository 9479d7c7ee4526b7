mkdir -p gpurun_out
R=r40
python tools/step_prof.py 2 > gpurun_out/${R}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 200 -c 198 --csv --log-file gpurun_out/${R}_launches.csv python tools/step_prof.py 2 > gpurun_out/${R}_ncu1.log 2>&1
tail -n 2 gpurun_out/${R}_ncu1.log; grep -c "gpu__time_duration" gpurun_out/${R}_launches.csv
