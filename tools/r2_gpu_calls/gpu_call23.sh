#!/bin/bash
# GPU call 23: launch-shape knobs for one rank's share of an 8-rank (and 4-rank) run, on one GPU
mkdir -p gpurun_out
timeout 200 python tools/ab_rankshare.py 8 6 > gpurun_out/c23_rankshare.log 2>&1
cut -c1-200 gpurun_out/c23_rankshare.log
