mkdir -p gpurun_out
python tools/attn_prof.py > gpurun_out/r13_attn_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"pkernel" -s 6 -c 3 -o gpurun_out/r13_attn -f python tools/attn_prof.py > gpurun_out/r13_ncu.log 2>&1
tail -5 gpurun_out/r13_ncu.log
