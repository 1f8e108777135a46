#!/bin/bash
# 2-GPU: the sharded gen_data test, then the bench under torchrun (strong scaling = default, then weak)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_distributed.py -m gpu -q > gpurun_out/r02g_pytest_2gpu.txt 2>&1; tail -5 gpurun_out/r02g_pytest_2gpu.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 --no-workloads > gpurun_out/r02g_bench_2gpu_strong.log 2>&1; grep '^{' gpurun_out/r02g_bench_2gpu_strong.log | cut -c1-700 || tail -20 gpurun_out/r02g_bench_2gpu_strong.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 3 --warmup 3 --no-workloads --scaling weak > gpurun_out/r02g_bench_2gpu_weak.log 2>&1; grep '^{' gpurun_out/r02g_bench_2gpu_weak.log | cut -c1-400 || tail -20 gpurun_out/r02g_bench_2gpu_weak.log
