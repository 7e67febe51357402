timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu.log 2>&1; tail -3 gpurun_out/t_gpu.log
GRM_CUDA_SYRK_I8=2cta timeout 300 python scripts/first_light.py int_small > gpurun_out/fl_2cta_small.log 2>&1; tail -8 gpurun_out/fl_2cta_small.log
GRM_CUDA_SYRK_I8=2cta timeout 300 python scripts/first_light.py int_c2 > gpurun_out/fl_2cta_c2.log 2>&1; grep -E "syrk|int_mm" gpurun_out/fl_2cta_c2.log | cut -c1-200
GRM_CUDA_SYRK_I8=2cta timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu_2cta.log 2>&1; tail -3 gpurun_out/t_gpu_2cta.log
timeout 300 python bench.py --no-e2e --no-cpu > gpurun_out/bench_c3_1cta.log 2>&1; tail -c 900 gpurun_out/bench_c3_1cta.log
GRM_CUDA_SYRK_I8=2cta timeout 300 python bench.py --no-e2e --no-cpu > gpurun_out/bench_c3_2cta.log 2>&1; tail -c 900 gpurun_out/bench_c3_2cta.log
