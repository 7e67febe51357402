timeout 500 python scripts/c4_demo.py > gpurun_out/c4_full.log 2>&1; tail -2 gpurun_out/c4_full.log
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/plain_c3.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"syrk|pack_dosage|rowstats|colsum_planes|centre_rows|scale_mirror|detect|finalize" -c 100 --csv --log-file gpurun_out/launches_c3.csv $CMD > gpurun_out/ncu_c3.log 2>&1
tail -2 gpurun_out/ncu_c3.log
