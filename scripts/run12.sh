timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu.log 2>&1; tail -3 gpurun_out/t_gpu.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1 && timeout 400 compute-sanitizer --tool memcheck --error-exitcode 3 python __graft_entry__.py smoke > gpurun_out/sanitizer_memcheck.log 2>&1; echo "sanitizer rc=$?"; tail -6 gpurun_out/sanitizer_memcheck.log
timeout 400 python bench.py --workload c2 > gpurun_out/bench_c2.log 2>&1; tail -c 300 gpurun_out/bench_c2.log
