#!/bin/bash
# GPU call 22: pipelined host staging (P and G cross PCIe in 4 pieces overlapped with the pinned-buffer copies) --
# full GPU suite, e2e vs kernel time, final bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/c22_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c22_pytest.log
tail -3 gpurun_out/c22_pytest.log
AB_TAG=staged4 timeout 300 python tools/ab.py 53 7 > gpurun_out/c22_ab.log 2>&1
AB_TAG=caffeine AB_MOL=caffeine timeout 300 python tools/ab.py 0 7 >> gpurun_out/c22_ab.log 2>&1
cut -c1-260 gpurun_out/c22_ab.log
timeout 500 python bench.py --scf > gpurun_out/c22_bench_n1.json 2> gpurun_out/c22_bench_n1.err
tail -c 1200 gpurun_out/c22_bench_n1.json
