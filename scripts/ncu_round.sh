# Round profile: (1) plain run, (2) launch list of the C2 bench, (3) --set full of the top kernels of one warm step.
# Run under gpurun on ONE GPU:  gpurun -- bash scripts/ncu_round.sh r02
TAG=${1:-r02}
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras"
$CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1
# matches per step: rank_pack 1, gemm 1, count_dedup 1, net_sort+net_pairs+net_label 3, net_rowsum<0> 1 = 7 (one 4 GiB strip holds every unique row of C2); skip setup (1) + warm-up step (7) - the first match of the step is rank_pack
$CMD > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"count_dedup|gemm_i8_tc|net_sort|net_pairs|net_label|rank_pack|net_rowsum_kernel<0>" -s 6 -c 7 -o gpurun_out/${TAG}_full $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -n 1 gpurun_out/${TAG}_ncu_launches.log; tail -n 1 gpurun_out/${TAG}_ncu_full.log
