#!/bin/bash
# GPU call 7: RED threshold factor sweep with the N = 1007 parity measured for each, full suite on the final code
mkdir -p gpurun_out
: > gpurun_out/c7_ab.log
for f in 0 0.01 0.1 1; do
  AB_TAG=eps_factor_$f QCF_RED_EPS_FACTOR=$f timeout 600 python tools/ab.py 53 5 >> gpurun_out/c7_ab.log 2>&1
  QCF_RED_EPS_FACTOR=$f timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -q -s --timeout 900 -k "benchmark_size and 53" > gpurun_out/c7_parity_$f.log 2>&1
  echo "factor $f: $(grep -E 'N=1007' gpurun_out/c7_parity_$f.log | cut -c1-120) $(tail -1 gpurun_out/c7_parity_$f.log)" >> gpurun_out/c7_ab.log
done
AB_TAG=det QCF_DETERMINISTIC=1 timeout 600 python tools/ab.py 53 5 >> gpurun_out/c7_ab.log 2>&1
cat gpurun_out/c7_ab.log
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/c7_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c7_pytest.log
tail -6 gpurun_out/c7_pytest.log
timeout 900 python bench.py --scf > gpurun_out/c7_bench_n1.json 2> gpurun_out/c7_bench_n1.err
tail -c 300 gpurun_out/c7_bench_n1.json
