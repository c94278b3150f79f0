#!/bin/bash
# GPU call 3: memory-latency measures of the block kernel, one at a time (variants A..E), whole build + per-class profile
mkdir -p gpurun_out
AB_TAG=A_all timeout 600 python tools/ab.py 53 5 > gpurun_out/c3_ab.log 2>&1
for v in B C D E; do
  AB_TAG=$v QCF_LIB=qchem-rs_b200/libqcfock_$v.so timeout 600 python tools/ab.py 53 5 >> gpurun_out/c3_ab.log 2>&1
done
AB_TAG=prof_A QCF_PROFILE=1 timeout 600 python tools/ab.py 53 2 250 > gpurun_out/c3_profile_A.log 2>&1
for v in B C D E; do
  AB_TAG=prof_$v QCF_PROFILE=1 QCF_LIB=qchem-rs_b200/libqcfock_$v.so timeout 600 python tools/ab.py 53 2 250 > gpurun_out/c3_profile_$v.log 2>&1
done
cat gpurun_out/c3_ab.log
timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -q --timeout 900 -x -k "benchmark_size or deterministic_mode or graph" > gpurun_out/c3_pytest.log 2>&1
tail -3 gpurun_out/c3_pytest.log
