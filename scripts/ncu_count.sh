CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"count_dedup|net_group|net_pass2" -s 3 -c 3 -o gpurun_out/prof_r1b $CMD > gpurun_out/ncu3.log 2>&1
tail -2 gpurun_out/ncu3.log
