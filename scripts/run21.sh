timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu.log 2>&1; tail -4 gpurun_out/t_gpu.log
timeout 500 python scripts/c4_demo.py > gpurun_out/c4_full.log 2>&1; tail -1 gpurun_out/c4_full.log
