# ncu --set full of the tcgen05 GEMM at the FANTOM5 shape (K = 1829 -> 15 K blocks); run under gpurun, 1 GPU
CMD="python bench.py --shape full --loci 400 --perms 15 --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain5.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"gemm_i8_tc|rank_pack" -s 2 -c 3 -o gpurun_out/prof_gemm_full $CMD > gpurun_out/ncu5.log 2>&1
tail -2 gpurun_out/ncu5.log
