# launch list (device time per launch) of one warm step of the C2 bench (run under gpurun, 1 GPU)
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain4.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_tmp.csv $CMD > gpurun_out/ncu4.log 2>&1
tail -1 gpurun_out/ncu4.log
