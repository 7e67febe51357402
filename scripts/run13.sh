timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu.log 2>&1; tail -5 gpurun_out/t_gpu.log
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 500 python bench.py > gpurun_out/bench_c3.log 2>&1; tail -c 400 gpurun_out/bench_c3.log
