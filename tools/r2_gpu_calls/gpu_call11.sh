#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/c11_ab.log
for f in 1 10; do
  AB_TAG=prim_cut_$f QCF_PRIM_CUT=$f timeout 600 python tools/ab.py 53 5 >> gpurun_out/c11_ab.log 2>&1
  QCF_PRIM_CUT=$f timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -q -s --timeout 900 -k "benchmark_size and 53" > gpurun_out/c11_parity_prim_$f.log 2>&1
  echo "prim cut $f: $(grep -E 'N=1007' gpurun_out/c11_parity_prim_$f.log | cut -c1-140) $(tail -1 gpurun_out/c11_parity_prim_$f.log)" >> gpurun_out/c11_ab.log
done
for f in 1e-1 1; do
  AB_TAG=pair_cut_$f QCF_PAIR_CUT=$f timeout 600 python tools/ab.py 53 5 >> gpurun_out/c11_ab.log 2>&1
  QCF_PAIR_CUT=$f timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -q -s --timeout 900 -k "benchmark_size and 53" > gpurun_out/c11_parity_pair_$f.log 2>&1
  echo "pair cut $f: $(grep -E 'N=1007' gpurun_out/c11_parity_pair_$f.log | cut -c1-140) $(tail -1 gpurun_out/c11_parity_pair_$f.log)" >> gpurun_out/c11_ab.log
done
cat gpurun_out/c11_ab.log | cut -c1-210
