#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/c10_ab.log
export AB_RANK=0 AB_WORLD=8
AB_TAG=r0of8_newdefault timeout 600 python tools/ab.py 53 6 >> gpurun_out/c10_ab.log 2>&1
AB_TAG=r0of8_streams24 QCF_STREAMS=24 timeout 600 python tools/ab.py 53 6 >> gpurun_out/c10_ab.log 2>&1
AB_TAG=r0of8_streams32 QCF_STREAMS=32 timeout 600 python tools/ab.py 53 6 >> gpurun_out/c10_ab.log 2>&1
AB_TAG=r0of8_ctas74 QCF_TARGET_CTAS=74 timeout 600 python tools/ab.py 53 6 >> gpurun_out/c10_ab.log 2>&1
AB_TAG=r0of8_s32_c74 QCF_STREAMS=32 QCF_TARGET_CTAS=74 timeout 600 python tools/ab.py 53 6 >> gpurun_out/c10_ab.log 2>&1
AB_TAG=r0of8_s32_kpt64 QCF_STREAMS=32 QCF_KETS_PER_THREAD=64 timeout 600 python tools/ab.py 53 6 >> gpurun_out/c10_ab.log 2>&1
export AB_WORLD=2
AB_TAG=r0of2_newdefault timeout 600 python tools/ab.py 53 5 >> gpurun_out/c10_ab.log 2>&1
AB_TAG=r0of2_olddefault QCF_STREAMS=8 QCF_TARGET_CTAS=296 timeout 600 python tools/ab.py 53 5 >> gpurun_out/c10_ab.log 2>&1
unset AB_RANK AB_WORLD
AB_TAG=full_streams16 QCF_STREAMS=16 timeout 600 python tools/ab.py 53 5 >> gpurun_out/c10_ab.log 2>&1
AB_TAG=full_ctas148 QCF_TARGET_CTAS=148 timeout 600 python tools/ab.py 53 5 >> gpurun_out/c10_ab.log 2>&1
AB_TAG=full_default timeout 600 python tools/ab.py 53 5 >> gpurun_out/c10_ab.log 2>&1
cat gpurun_out/c10_ab.log | cut -c1-170
