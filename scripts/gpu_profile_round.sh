#!/bin/bash
# One GPU-box visit that produces everything profiles/ cites for a round:
#   1. the -m gpu tests, 2. the default bench line (N=1) and the reference arm, 3. the ncu launch list of the same bench
#   command (gpu__time_duration.sum per launch), 4. one ncu --set full capture of the dominant kernels.
# Usage (from the repo root, through gpurun):  bash scripts/gpu_profile_round.sh r01h
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
CMD="python bench.py --steps 10 --warmup 3"
$CMD > $out/bench_${tag}.json 2> $out/bench_${tag}.err || { tail -5 $out/bench_${tag}.err; exit 1; }
python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_ref_${tag}.json 2> $out/bench_ref_${tag}.err
SHORT="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $out/launches_${tag}.csv $SHORT > $out/ncu_launches_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"clip_sh_kernel|clip_mom_kernel|candidate_single|scatter_kernel|scatter_long|order2_finalize" -s 6 -c 12 -o $out/${tag}_kernels $SHORT > $out/ncu_full_${tag}.log 2>&1
tail -2 $out/ncu_full_${tag}.log
ncu --set full --clock-control none -k regex:"apply_rec_kernel|grad_c2l_rec_kernel|gc_clip_kernel|gc_filter_kernel|gc_candidate_kernel" -s 2 -c 7 -o $out/${tag}_apply $SHORT > $out/ncu_apply_${tag}.log 2>&1
tail -1 $out/ncu_apply_${tag}.log
tail -c 400 $out/bench_${tag}.json
