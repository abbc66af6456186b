mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"candidate_kernel|clip_kernel|order2_finalize|scatter_kernel|heavy" -s 24 -c 24 -o gpurun_out/r01b $CMD > gpurun_out/ncu_full.log 2>&1
echo rc=$?
tail -3 gpurun_out/ncu_full.log
