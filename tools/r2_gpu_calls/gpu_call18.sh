#!/bin/bash
# GPU call 18: WIDE as a template parameter; launch knobs for small problems (caffeine, N = 190, N = 513)
mkdir -p gpurun_out
: > gpurun_out/c18_ab.log
AB_TAG=n53 timeout 600 python tools/ab.py 53 5 >> gpurun_out/c18_ab.log 2>&1
AB_TAG=n27_default timeout 600 python tools/ab.py 27 5 >> gpurun_out/c18_ab.log 2>&1
AB_TAG=n27_s16_c148 QCF_STREAMS=16 QCF_TARGET_CTAS=148 timeout 600 python tools/ab.py 27 5 >> gpurun_out/c18_ab.log 2>&1
AB_TAG=n27_s32_c74_k32 QCF_STREAMS=32 QCF_TARGET_CTAS=74 QCF_KETS_PER_THREAD=32 timeout 600 python tools/ab.py 27 5 >> gpurun_out/c18_ab.log 2>&1
AB_TAG=caffeine_default AB_MOL=caffeine timeout 600 python tools/ab.py 0 8 >> gpurun_out/c18_ab.log 2>&1
AB_TAG=caffeine_s16_c148 AB_MOL=caffeine QCF_STREAMS=16 QCF_TARGET_CTAS=148 timeout 600 python tools/ab.py 0 8 >> gpurun_out/c18_ab.log 2>&1
AB_TAG=caffeine_s32_c74_k32 AB_MOL=caffeine QCF_STREAMS=32 QCF_TARGET_CTAS=74 QCF_KETS_PER_THREAD=32 timeout 600 python tools/ab.py 0 8 >> gpurun_out/c18_ab.log 2>&1
AB_TAG=caffeine_s32_c296 AB_MOL=caffeine QCF_STREAMS=32 timeout 600 python tools/ab.py 0 8 >> gpurun_out/c18_ab.log 2>&1
AB_TAG=n10_default timeout 600 python tools/ab.py 10 8 >> gpurun_out/c18_ab.log 2>&1
AB_TAG=n10_s32_c74_k32 QCF_STREAMS=32 QCF_TARGET_CTAS=74 QCF_KETS_PER_THREAD=32 timeout 600 python tools/ab.py 10 8 >> gpurun_out/c18_ab.log 2>&1
AB_TAG=n10_s32_c296 QCF_STREAMS=32 timeout 600 python tools/ab.py 10 8 >> gpurun_out/c18_ab.log 2>&1
AB_RANK=0 AB_WORLD=8 AB_TAG=r0of8 timeout 600 python tools/ab.py 53 6 >> gpurun_out/c18_ab.log 2>&1
cat gpurun_out/c18_ab.log | cut -c1-160
