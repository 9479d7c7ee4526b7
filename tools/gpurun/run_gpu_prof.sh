mkdir -p gpurun_out
R=${R:-r02o}
# ncu --set full of the two attention kernels that changed this round (each command first runs plain, then under ncu)
python tools/attn_prof.py > gpurun_out/${R}_plain1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:bwd_tc_kernel -s 1 -c 1 -o gpurun_out/${R}_bwd_tc python tools/attn_prof.py > gpurun_out/${R}_ncu1.log 2>&1
ICKB200_DECODE_GRAPH=0 python tools/bench_predict.py --variant K --beam 5 --reps 1 > gpurun_out/${R}_plain2.log 2>&1 && \
ICKB200_DECODE_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:mha_decode_tma_mma -s 20 -c 1 -o gpurun_out/${R}_beam_xattn python tools/bench_predict.py --variant K --beam 5 --reps 1 > gpurun_out/${R}_ncu2.log 2>&1
tail -n 2 gpurun_out/${R}_ncu1.log gpurun_out/${R}_ncu2.log; ls -la gpurun_out/*.ncu-rep | tail -3
