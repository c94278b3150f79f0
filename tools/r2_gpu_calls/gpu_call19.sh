#!/bin/bash
# GPU call 19: shared-memory exchange rows (KROWS) -- A/B at N = 1007 in one process, then the whole GPU suite with the
# KROWS instantiation forced on for every launch it supports
mkdir -p gpurun_out
timeout 400 python tools/r2_gpu_calls/not_kept/ab_krows.py 53 5 > gpurun_out/c19_ab.log 2>&1
cat gpurun_out/c19_ab.log | cut -c1-200
QCF_KROWS_MAX_PRIM=1000000 timeout 420 python -m pytest tests -m gpu -q -s -k "not multi_gpu" > gpurun_out/c19_tests_krows_forced.log 2>&1
tail -25 gpurun_out/c19_tests_krows_forced.log | cut -c1-200
