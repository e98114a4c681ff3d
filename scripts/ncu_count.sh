# ncu --set full capture of the count kernel on the C2 bench (run under gpurun, 1 GPU)
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"count_dedup" -s 1 -c 1 -o gpurun_out/prof_r1c $CMD > gpurun_out/ncu3.log 2>&1
tail -2 gpurun_out/ncu3.log
