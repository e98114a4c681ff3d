# ncu --set full capture of the count kernel on the C2 bench (run under gpurun, 1 GPU):  [KERN=count_fine] [BENCH_ARGS="--rows ..."] bash scripts/ncu_count.sh TAG
TAG=${1:-r02_count}
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras ${BENCH_ARGS}"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"${KERN:-count_dedup}" -s 1 -c 1 -o gpurun_out/${TAG} $CMD > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log
