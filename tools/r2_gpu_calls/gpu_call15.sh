#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/c15_ab.log
AB_TAG=base timeout 600 python tools/ab.py 53 5 >> gpurun_out/c15_ab.log 2>&1
AB_TAG=prio1_block_high QCF_PRIO=1 timeout 600 python tools/ab.py 53 5 >> gpurun_out/c15_ab.log 2>&1
AB_TAG=prio2_slab_high QCF_PRIO=2 timeout 600 python tools/ab.py 53 5 >> gpurun_out/c15_ab.log 2>&1
AB_TAG=prio1_s16 QCF_PRIO=1 QCF_STREAMS=16 timeout 600 python tools/ab.py 53 5 >> gpurun_out/c15_ab.log 2>&1
AB_TAG=prio2_s16 QCF_PRIO=2 QCF_STREAMS=16 timeout 600 python tools/ab.py 53 5 >> gpurun_out/c15_ab.log 2>&1
AB_RANK=0 AB_WORLD=8 AB_TAG=r0of8_base timeout 600 python tools/ab.py 53 6 >> gpurun_out/c15_ab.log 2>&1
AB_RANK=0 AB_WORLD=8 AB_TAG=r0of8_prio1 QCF_PRIO=1 timeout 600 python tools/ab.py 53 6 >> gpurun_out/c15_ab.log 2>&1
AB_RANK=0 AB_WORLD=8 AB_TAG=r0of8_prio2 QCF_PRIO=2 timeout 600 python tools/ab.py 53 6 >> gpurun_out/c15_ab.log 2>&1
cat gpurun_out/c15_ab.log | cut -c1-150
