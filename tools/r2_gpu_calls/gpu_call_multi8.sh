#!/bin/bash
# 8-GPU call: in-library scaling 1/2/4/8 on one box, torchrun bench at 8 and 4, multi-GPU tests
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/m8_topo.txt 2>&1
: > gpurun_out/m8_ab.log
for k in 1 2 4 8; do
  AB_TAG=inlib_$k AB_NGPUS=$k timeout 600 python tools/ab.py 53 5 >> gpurun_out/m8_ab.log 2>&1
done
AB_TAG=inlib_8_n105 AB_NGPUS=8 timeout 600 python tools/ab.py 105 3 >> gpurun_out/m8_ab.log 2>&1
cat gpurun_out/m8_ab.log
for k in 8 4; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $k --master-addr 127.0.0.1 --master-port 2952$k bench.py --gpus $k --steps 5 --warmup 3 > gpurun_out/m8_bench_n$k.json 2> gpurun_out/m8_bench_n$k.err
  tail -c 900 gpurun_out/m8_bench_n$k.json
done
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -q --timeout 600 -k "multi_gpu" > gpurun_out/m8_pytest.log 2>&1
tail -3 gpurun_out/m8_pytest.log
