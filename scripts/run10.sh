timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu.log 2>&1; tail -5 gpurun_out/t_gpu.log
nvidia-smi topo -m | head -12
for N in 8 4 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_c3_n$N.log 2>&1; tail -c 400 gpurun_out/bench_c3_n$N.log; echo
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 scripts/dist_check.py > gpurun_out/dist_check8.log 2>&1; grep -E "rank [0-7]/" gpurun_out/dist_check8.log | sort | head -30
