#!/bin/bash
# GPU call 2: validate the AoS / 256-bit-load build, register-cap variants of the block kernel, ncu on the low-L classes
mkdir -p gpurun_out
AB_TAG=aos_default timeout 600 python tools/ab.py 53 5 > gpurun_out/c2_ab.log 2>&1
for v in mb4 mb5 mb6 mb8; do
  AB_TAG=$v QCF_LIB=qchem-rs_b200/libqcfock_$v.so timeout 600 python tools/ab.py 53 5 >> gpurun_out/c2_ab.log 2>&1
done
AB_TAG=prof_default QCF_PROFILE=1 timeout 600 python tools/ab.py 53 2 250 > gpurun_out/c2_profile_default.log 2>&1
for v in mb4 mb5 mb6 mb8; do
  AB_TAG=prof_$v QCF_PROFILE=1 QCF_LIB=qchem-rs_b200/libqcfock_$v.so timeout 600 python tools/ab.py 53 2 250 > gpurun_out/c2_profile_$v.log 2>&1
done
cat gpurun_out/c2_ab.log
timeout 1500 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -q --timeout 900 -x > gpurun_out/c2_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c2_pytest.log
tail -4 gpurun_out/c2_pytest.log
# ncu: the (ss|ss) and (ps|ss) launches of one build, stream launches (no graph), after the same command ran clean
QCF_NO_GRAPH=1 timeout 600 python tools/ab.py 53 1 > gpurun_out/c2_ncu_plain.log 2>&1 && \
QCF_NO_GRAPH=1 timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name-base mangled \
  -k regex:'eri_jk_kernelILi[01]ELi0ELi0ELi0ELi1ELi1E' -s 26 -c 12 -o gpurun_out/c2_prof_lowL python tools/ab.py 53 1 > gpurun_out/c2_ncu.log 2>&1
tail -3 gpurun_out/c2_ncu.log
