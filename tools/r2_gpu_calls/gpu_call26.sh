#!/bin/bash
# GPU call 26: per-launch split balanced on the measured per-class quartet times (QCF_SPLIT_COST=1), eight ranks on one GPU
mkdir -p gpurun_out
AB_SPLIT_CONFIGS="measured_148,measured_296,measured_592,measured_1184" timeout 150 python tools/ab_split.py 8 3 > gpurun_out/c26_split.log 2>&1
cut -c1-330 gpurun_out/c26_split.log
