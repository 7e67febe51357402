timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu.log 2>&1; tail -3 gpurun_out/t_gpu.log
CMD="python bench.py --workload c2 --steps 2 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"syrk|pack_dosage|rowstats|colsum_planes|centre_rows|scale_mirror|detect" -c 200 --csv --log-file gpurun_out/launches_c2.csv $CMD > gpurun_out/ncu1.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:syrk_i8 -s 3 -c 1 -o gpurun_out/prof_syrk_c2 $CMD > gpurun_out/ncu2.log 2>&1
CMD2="python scripts/first_light.py f64"
$CMD2 > gpurun_out/plain_f64.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:syrk_f64 -s 6 -c 1 -o gpurun_out/prof_syrk_f64 $CMD2 > gpurun_out/ncu_f64.log 2>&1
tail -3 gpurun_out/ncu_f64.log
