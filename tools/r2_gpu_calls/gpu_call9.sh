#!/bin/bash
# GPU call 9: what one rank of an 8-rank run does, on one GPU (rank 0 of 8): knob sweep + serialised per-class profile
mkdir -p gpurun_out
: > gpurun_out/c9_ab.log
export AB_RANK=0 AB_WORLD=8
AB_TAG=r0of8_default timeout 600 python tools/ab.py 53 6 >> gpurun_out/c9_ab.log 2>&1
AB_TAG=r0of8_streams16 QCF_STREAMS=16 timeout 600 python tools/ab.py 53 6 >> gpurun_out/c9_ab.log 2>&1
AB_TAG=r0of8_streams4 QCF_STREAMS=4 timeout 600 python tools/ab.py 53 6 >> gpurun_out/c9_ab.log 2>&1
AB_TAG=r0of8_kpt16 QCF_KETS_PER_THREAD=16 timeout 600 python tools/ab.py 53 6 >> gpurun_out/c9_ab.log 2>&1
AB_TAG=r0of8_kpt64 QCF_KETS_PER_THREAD=64 timeout 600 python tools/ab.py 53 6 >> gpurun_out/c9_ab.log 2>&1
AB_TAG=r0of8_ctas592 QCF_TARGET_CTAS=592 timeout 600 python tools/ab.py 53 6 >> gpurun_out/c9_ab.log 2>&1
AB_TAG=r0of8_ctas148 QCF_TARGET_CTAS=148 timeout 600 python tools/ab.py 53 6 >> gpurun_out/c9_ab.log 2>&1
AB_TAG=r0of8_psmin18 QCF_PS_MIN=18 timeout 600 python tools/ab.py 53 6 >> gpurun_out/c9_ab.log 2>&1
AB_TAG=r0of8_psoff QCF_PS_MIN=100000 timeout 600 python tools/ab.py 53 6 >> gpurun_out/c9_ab.log 2>&1
AB_TAG=r0of8_order1 QCF_ORDER=1 timeout 600 python tools/ab.py 53 6 >> gpurun_out/c9_ab.log 2>&1
AB_TAG=r0of8_nograph QCF_NO_GRAPH=1 timeout 600 python tools/ab.py 53 6 >> gpurun_out/c9_ab.log 2>&1
AB_TAG=r0of8_profile QCF_PROFILE=1 timeout 600 python tools/ab.py 53 2 250 > gpurun_out/c9_profile_r0of8.log 2>&1
unset AB_RANK AB_WORLD
AB_TAG=full_profile QCF_PROFILE=1 timeout 600 python tools/ab.py 53 2 250 > gpurun_out/c9_profile_full.log 2>&1
cat gpurun_out/c9_ab.log
