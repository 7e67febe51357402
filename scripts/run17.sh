timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu.log 2>&1; tail -4 gpurun_out/t_gpu.log
C5_N=6000 C5_P=3000 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 scripts/c5_demo.py > gpurun_out/c5_small.log 2>&1; tail -2 gpurun_out/c5_small.log
