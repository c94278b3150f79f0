#!/bin/bash
# GPU call 4: SoA default restored; ILP variants F/G (AoS-era builds, compare per class with c3 variant C);
# runtime-knob sweep on the default build
mkdir -p gpurun_out
: > gpurun_out/c4_ab.log
AB_TAG=soa_default timeout 600 python tools/ab.py 53 5 >> gpurun_out/c4_ab.log 2>&1
for v in F G; do
  AB_TAG=$v QCF_LIB=qchem-rs_b200/libqcfock_$v.so timeout 600 python tools/ab.py 53 5 >> gpurun_out/c4_ab.log 2>&1
  AB_TAG=prof_$v QCF_PROFILE=1 QCF_LIB=qchem-rs_b200/libqcfock_$v.so timeout 600 python tools/ab.py 53 2 250 > gpurun_out/c4_profile_$v.log 2>&1
done
AB_TAG=prof_soa QCF_PROFILE=1 timeout 600 python tools/ab.py 53 2 250 > gpurun_out/c4_profile_soa.log 2>&1
AB_TAG=block128 QCF_BLOCK=128 timeout 600 python tools/ab.py 53 4 >> gpurun_out/c4_ab.log 2>&1
AB_TAG=block32 QCF_BLOCK=32 timeout 600 python tools/ab.py 53 4 >> gpurun_out/c4_ab.log 2>&1
AB_TAG=order1 QCF_ORDER=1 timeout 600 python tools/ab.py 53 4 >> gpurun_out/c4_ab.log 2>&1
AB_TAG=streams4 QCF_STREAMS=4 timeout 600 python tools/ab.py 53 4 >> gpurun_out/c4_ab.log 2>&1
AB_TAG=streams12 QCF_STREAMS=12 timeout 600 python tools/ab.py 53 4 >> gpurun_out/c4_ab.log 2>&1
AB_TAG=kpt128 QCF_KETS_PER_THREAD=128 timeout 600 python tools/ab.py 53 4 >> gpurun_out/c4_ab.log 2>&1
AB_TAG=ctas592 QCF_TARGET_CTAS=592 timeout 600 python tools/ab.py 53 4 >> gpurun_out/c4_ab.log 2>&1
AB_TAG=psmin9 QCF_PS_MIN=9 timeout 600 python tools/ab.py 53 4 >> gpurun_out/c4_ab.log 2>&1
cat gpurun_out/c4_ab.log
