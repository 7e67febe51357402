timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu.log 2>&1; tail -5 gpurun_out/t_gpu.log
timeout 300 python scripts/first_light.py int_c2 > gpurun_out/fl_c2d.log 2>&1; grep -E "pack|syrk" gpurun_out/fl_c2d.log | cut -c1-160
GRM_CUDA_PACK=ldg timeout 300 python scripts/first_light.py int_c2 > gpurun_out/fl_c2e.log 2>&1; grep -E "\] pack" gpurun_out/fl_c2e.log | cut -c1-160
timeout 500 python bench.py --no-e2e --no-cpu > gpurun_out/bench_c3_tma.log 2>&1; grep -o '"stage_ms": {[^}]*}' gpurun_out/bench_c3_tma.log; grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_c3_tma.log
GRM_CUDA_PACK=ldg timeout 500 python bench.py --no-e2e --no-cpu > gpurun_out/bench_c3_ldg.log 2>&1; grep -o '"stage_ms": {[^}]*}' gpurun_out/bench_c3_ldg.log; grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_c3_ldg.log
