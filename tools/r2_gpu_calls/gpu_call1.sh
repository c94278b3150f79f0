#!/bin/bash
# first GPU call of round 2: fixture, A/B timings of the new build path, per-launch profile, the whole GPU test suite
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/c1_smi.txt 2>&1
timeout 600 python tools/make_bench_density.py 53 6 gpurun_out/ > gpurun_out/c1_density.log 2>&1
AB_TAG=default timeout 600 python tools/ab.py 53 5 > gpurun_out/c1_ab.log 2>&1
AB_TAG=ps_off QCF_PS_MIN=1000000 timeout 600 python tools/ab.py 53 5 >> gpurun_out/c1_ab.log 2>&1
AB_TAG=nograph QCF_NO_GRAPH=1 timeout 600 python tools/ab.py 53 5 >> gpurun_out/c1_ab.log 2>&1
AB_TAG=det QCF_DETERMINISTIC=1 timeout 600 python tools/ab.py 53 5 >> gpurun_out/c1_ab.log 2>&1
AB_TAG=kpt32 QCF_KETS_PER_THREAD=32 timeout 600 python tools/ab.py 53 5 >> gpurun_out/c1_ab.log 2>&1
AB_TAG=streams16 QCF_STREAMS=16 timeout 600 python tools/ab.py 53 5 >> gpurun_out/c1_ab.log 2>&1
AB_TAG=profile QCF_PROFILE=1 timeout 600 python tools/ab.py 53 2 60 > gpurun_out/c1_profile.log 2>&1
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 -rA > gpurun_out/c1_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c1_pytest.log
tail -5 gpurun_out/c1_pytest.log; cat gpurun_out/c1_ab.log
