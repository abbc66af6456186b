# developer tool: time apply_rec_kernel / grad_c2l_rec_kernel of configs[1] under ncu for the default library and the variant libraries given
for lib in "" "$@"; do
  echo "lib=$lib"
  XGRID_B200_LIB=$lib timeout 300 ncu -k regex:rec_kernel --metrics gpu__time_duration.sum,smsp__inst_executed.sum,l1tex__t_sector_hit_rate.pct,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread --clock-control none -c 2 python scripts/profile_apply.py 2 2>&1 | grep -E "gpu__time|inst_executed|hit_rate|throughput|registers" 
done
