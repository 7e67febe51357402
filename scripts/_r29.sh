CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/plain_c3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:syrk_i8_2cta -s 3 -c 1 -o gpurun_out/prof_syrk_c3 $CMD > gpurun_out/ncu_c3a.log 2>&1
$CMD > gpurun_out/plain_c3c.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"syrk|pack_dosage|rowstats|colsum_planes|centre_rows|scale_mirror|detect|finalize" -c 100 --csv --log-file gpurun_out/launches_c3.csv $CMD > gpurun_out/ncu_c3.log 2>&1
timeout 600 python bench.py > gpurun_out/bench_c3.log 2>&1; tail -c 300 gpurun_out/bench_c3.log
timeout 300 python bench.py --impl reference > gpurun_out/bench_ref.log 2>&1; tail -c 300 gpurun_out/bench_ref.log
