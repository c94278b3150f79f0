#!/bin/bash
# GPU call 21 (two GPUs): the final code on the multi-GPU paths -- in-library test, 2-rank torchrun bench line
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_round2.py -m gpu -q --timeout 300 -k "multi_gpu" > gpurun_out/c21_pytest_2gpu.log 2>&1
tail -3 gpurun_out/c21_pytest_2gpu.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/c21_bench_n2.json 2> gpurun_out/c21_bench_n2.err
tail -c 2500 gpurun_out/c21_bench_n2.json
tail -3 gpurun_out/c21_bench_n2.err
