#!/bin/bash
# GPU call 24: per-launch bra split (QCF_SPLIT_MIN_BRAS) against the per-group split, every rank of an 8-rank run on one GPU
mkdir -p gpurun_out
timeout 200 python tools/ab_split.py 8 4 > gpurun_out/c24_split.log 2>&1
cut -c1-330 gpurun_out/c24_split.log
