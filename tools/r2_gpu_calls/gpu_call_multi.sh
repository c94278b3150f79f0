#!/bin/bash
# multi-GPU call: NG = number of GPUs of this box (2, 4 or 8)
NG=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/m${NG}_topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -q --timeout 600 -k "multi_gpu or cost_balanced" > gpurun_out/m${NG}_pytest.log 2>&1
tail -3 gpurun_out/m${NG}_pytest.log
: > gpurun_out/m${NG}_ab.log
for k in 1 2 4 8; do
  if [ $k -le $NG ]; then AB_TAG=inlib_$k AB_NGPUS=$k timeout 600 python tools/ab.py 53 5 >> gpurun_out/m${NG}_ab.log 2>&1; fi
done
cat gpurun_out/m${NG}_ab.log
for k in 2 4 8; do
  if [ $k -le $NG ]; then
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $k --master-addr 127.0.0.1 --master-port 2951$k bench.py --gpus $k --steps 5 --warmup 3 > gpurun_out/m${NG}_bench_n$k.json 2> gpurun_out/m${NG}_bench_n$k.err
    tail -c 1500 gpurun_out/m${NG}_bench_n$k.json
  fi
done
timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/m${NG}_bench_n1.json 2> gpurun_out/m${NG}_bench_n1.err
