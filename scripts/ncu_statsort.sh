# ncu --set full capture of stat_sort_kernel (set-up half of the count stage) on the C2 bench (run under gpurun, 1 GPU)
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain_ss.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"stat_sort" -s 2 -c 1 -o gpurun_out/prof_statsort $CMD > gpurun_out/ncu_ss.log 2>&1
tail -2 gpurun_out/ncu_ss.log
