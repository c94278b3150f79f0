#!/bin/bash
# GPU call 14: where does the block kernels' time go?  Timing-only variants (WRONG results by construction):
# Boys rows always row 0 (L1 hit), density gathers always element 0, no atomics at all.
mkdir -p gpurun_out
: > gpurun_out/c14_ab.log
AB_TAG=base timeout 600 python tools/ab.py 53 4 >> gpurun_out/c14_ab.log 2>&1
AB_TAG=no_atomics QCF_RED_EPS_FACTOR=1e30 timeout 600 python tools/ab.py 53 4 >> gpurun_out/c14_ab.log 2>&1
for v in fakeboys fakegather fakeboth; do
  AB_TAG=$v QCF_LIB=qchem-rs_b200/libqcfock_$v.so timeout 600 python tools/ab.py 53 4 >> gpurun_out/c14_ab.log 2>&1
done
AB_TAG=fakeboth_no_atomics QCF_RED_EPS_FACTOR=1e30 QCF_LIB=qchem-rs_b200/libqcfock_fakeboth.so timeout 600 python tools/ab.py 53 4 >> gpurun_out/c14_ab.log 2>&1
AB_TAG=prof_fakeboth_no_atomics QCF_PROFILE=1 QCF_RED_EPS_FACTOR=1e30 QCF_LIB=qchem-rs_b200/libqcfock_fakeboth.so timeout 600 python tools/ab.py 53 2 250 > gpurun_out/c14_profile_fakeall.log 2>&1
AB_TAG=prof_no_atomics QCF_PROFILE=1 QCF_RED_EPS_FACTOR=1e30 timeout 600 python tools/ab.py 53 2 250 > gpurun_out/c14_profile_noatomics.log 2>&1
cat gpurun_out/c14_ab.log | cut -c1-170
bash tools/gpu_call13.sh
