export GRM_CUDA_SYRK_I8=2cta
timeout 300 python scripts/first_light.py int_small > gpurun_out/fl_2cta_small.log 2>&1; tail -7 gpurun_out/fl_2cta_small.log
timeout 300 python scripts/first_light.py int_c2 > gpurun_out/fl_2cta_c2.log 2>&1; grep -E "syrk|int_mm" gpurun_out/fl_2cta_c2.log | cut -c1-200
timeout 300 python bench.py --no-e2e --no-cpu > gpurun_out/bench_c3_2cta.log 2>&1; tail -c 600 gpurun_out/bench_c3_2cta.log
unset GRM_CUDA_SYRK_I8
timeout 300 python bench.py --no-e2e --no-cpu > gpurun_out/bench_c3_1cta.log 2>&1; tail -c 600 gpurun_out/bench_c3_1cta.log
