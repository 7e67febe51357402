timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu.log 2>&1; tail -3 gpurun_out/t_gpu.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py > gpurun_out/dist_check2.log 2>&1; grep -E "rank [01]/" gpurun_out/dist_check2.log
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_c3_n2.log 2>&1; tail -c 1200 gpurun_out/bench_c3_n2.log
timeout 300 python scripts/first_light.py int_c2 > gpurun_out/fl_c2c.log 2>&1; grep -E "syrk" gpurun_out/fl_c2c.log | cut -c1-200
