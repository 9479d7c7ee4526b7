mkdir -p gpurun_out
R=r77
ICKB200_DECODE_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:mha_decode_rows -s 6 -c 1 -o gpurun_out/${R}_decode_rows python tools/bench_predict.py --variant K --reps 1 > gpurun_out/${R}_ncu1.log 2>&1
ICKB200_DECODE_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:fwd_pkernel -s 9 -c 1 -o gpurun_out/${R}_beam_xattn python tools/bench_predict.py --variant K --beam 5 --reps 1 > gpurun_out/${R}_ncu2.log 2>&1
tail -n 2 gpurun_out/${R}_ncu1.log gpurun_out/${R}_ncu2.log | cut -c1-150; ls -la gpurun_out/*.ncu-rep
