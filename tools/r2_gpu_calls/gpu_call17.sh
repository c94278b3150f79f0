#!/bin/bash
# GPU call 17: adaptive scan width -- small molecules (BASELINE configs) and the usual sizes; rank share of 8
mkdir -p gpurun_out
timeout 900 python tools/config_times.py > gpurun_out/c17_baseline_configs.jsonl 2> gpurun_out/c17_baseline_configs.err
cat gpurun_out/c17_baseline_configs.jsonl | cut -c1-260
: > gpurun_out/c17_ab.log
AB_TAG=n53 timeout 600 python tools/ab.py 53 5 >> gpurun_out/c17_ab.log 2>&1
AB_TAG=n27 timeout 600 python tools/ab.py 27 5 >> gpurun_out/c17_ab.log 2>&1
AB_TAG=n10 timeout 600 python tools/ab.py 10 5 >> gpurun_out/c17_ab.log 2>&1
AB_RANK=0 AB_WORLD=8 AB_TAG=r0of8 timeout 600 python tools/ab.py 53 6 >> gpurun_out/c17_ab.log 2>&1
AB_RANK=0 AB_WORLD=8 AB_TAG=r0of8_n27 timeout 600 python tools/ab.py 27 6 >> gpurun_out/c17_ab.log 2>&1
cat gpurun_out/c17_ab.log | cut -c1-170
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/c17_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c17_pytest.log
tail -3 gpurun_out/c17_pytest.log
