timeout 300 python scripts/first_light.py f64 > gpurun_out/fl_f64b.log 2>&1; tail -5 gpurun_out/fl_f64b.log
export GRM_CUDA_SYRK_I8=2cta
CMD="python bench.py --workload c2 --steps 2 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/plain_2cta.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:syrk_i8_2cta -s 3 -c 1 -o gpurun_out/prof_syrk_2cta_c2 $CMD > gpurun_out/ncu_2cta.log 2>&1
tail -2 gpurun_out/ncu_2cta.log
