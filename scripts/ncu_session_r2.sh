#!/bin/bash
# Round-2 profiler session (1 GPU).  Every step has its own timeout; ncu only instruments this library's kernels.
set -u
OUT=gpurun_out
STEP="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-f64 --no-libpeaks --no-verify"
MINE='regex:pack_|syrk_|rowstats|colsum_planes|centre_rows|scale_mirror|locus_|mask_codes'
timeout 150 $STEP > $OUT/r2_plain_c3.log 2>&1 || { echo "plain run failed"; tail -3 $OUT/r2_plain_c3.log; exit 1; }
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -k "$MINE" -c 80 --csv --log-file $OUT/launches_r2_c3.csv $STEP > $OUT/r2_ncu_launch.log 2>&1; echo "launch list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:syrk_i8_wide -s 3 -c 1 -f -o $OUT/prof_r2_syrk_wide_c3 $STEP > $OUT/r2_ncu_wide.log 2>&1; echo "wide rc=$?"
C2E="python bench.py --workload c2 --steps 1 --warmup 3 --e2e-steps 1 --no-cpu --no-f64 --no-libpeaks --no-verify"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:pack_codes -s 2 -c 1 -f -o $OUT/prof_r2_pack_codes_c2 $C2E > $OUT/r2_ncu_packcodes.log 2>&1; echo "pack_codes rc=$?"
C2F="python bench.py --workload c2 --steps 1 --warmup 3 --no-e2e --no-cpu --no-libpeaks --no-verify"
timeout 240 ncu --set full --clock-control none --import-source on -k regex:syrk_f64_tma -s 2 -c 1 -f -o $OUT/prof_r2_syrk_f64_tma $C2F > $OUT/r2_ncu_f64.log 2>&1; echo "f64 rc=$?"
ls -la $OUT/*r2*.ncu-rep $OUT/launches_r2_c3.csv 2>&1 | tail -6
