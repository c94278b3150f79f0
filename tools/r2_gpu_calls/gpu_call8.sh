#!/bin/bash
# GPU call 8: primitive-pair cut sweep (parity measured for each) + full suite at the new defaults
mkdir -p gpurun_out
: > gpurun_out/c8_ab.log
for f in 1e-4 1e-3 1e-2 1e-1; do
  AB_TAG=prim_cut_$f QCF_PRIM_CUT=$f timeout 600 python tools/ab.py 53 5 >> gpurun_out/c8_ab.log 2>&1
  QCF_PRIM_CUT=$f timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -q -s --timeout 900 -k "benchmark_size and 53" > gpurun_out/c8_parity_$f.log 2>&1
  echo "prim cut $f: $(grep -E 'N=1007' gpurun_out/c8_parity_$f.log | cut -c1-140) $(tail -1 gpurun_out/c8_parity_$f.log)" >> gpurun_out/c8_ab.log
done
cat gpurun_out/c8_ab.log
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/c8_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c8_pytest.log
tail -6 gpurun_out/c8_pytest.log
