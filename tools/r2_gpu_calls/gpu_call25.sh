#!/bin/bash
# GPU call 25: per-launch bra split as the default from four ranks on -- split tests, every rank of an 8- and a 4-rank run
# on one GPU, new default against the per-group split (QCF_SPLIT_MIN_BRAS=0)
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_round2.py -m gpu -q --timeout 200 -k "cost_balanced or graph_replay" > gpurun_out/c25_pytest.log 2>&1
tail -3 gpurun_out/c25_pytest.log
AB_SPLIT_CONFIGS="default,per_group" timeout 200 python tools/ab_split.py 8 4 > gpurun_out/c25_split.log 2>&1
AB_SPLIT_CONFIGS="default,per_group,per_launch_296" timeout 200 python tools/ab_split.py 4 4 >> gpurun_out/c25_split.log 2>&1
cut -c1-330 gpurun_out/c25_split.log
