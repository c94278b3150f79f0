#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/c6_ab.log
for v in eps15 eps14; do
  AB_TAG=$v QCF_LIB=qchem-rs_b200/libqcfock_$v.so timeout 600 python tools/ab.py 53 5 >> gpurun_out/c6_ab.log 2>&1
  QCF_LIB=qchem-rs_b200/libqcfock_$v.so timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -q -s --timeout 900 -k "benchmark_size and 53" > gpurun_out/c6_parity_$v.log 2>&1
  grep -E "N=1007|passed|failed" gpurun_out/c6_parity_$v.log
done
AB_TAG=prof_eps15 QCF_PROFILE=1 QCF_LIB=qchem-rs_b200/libqcfock_eps15.so timeout 600 python tools/ab.py 53 2 250 > gpurun_out/c6_profile_eps15.log 2>&1
cat gpurun_out/c6_ab.log
