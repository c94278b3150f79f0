#!/bin/bash
# GPU call 28: per-launch split at two ranks (threshold 592 = 1184 / 2) against the per-group split, rank by rank on one GPU
mkdir -p gpurun_out
AB_SPLIT_CONFIGS="per_group,per_launch_592,per_launch_296" timeout 80 python tools/ab_split.py 2 4 > gpurun_out/c28_split_w2.log 2>&1
cut -c1-300 gpurun_out/c28_split_w2.log
