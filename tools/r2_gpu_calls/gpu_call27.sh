#!/bin/bash
# GPU call 27: last check of the final binary -- smoke() and the GPU suite without the four slow size-parity tests
mkdir -p gpurun_out
timeout 40 python __graft_entry__.py smoke > gpurun_out/c27_smoke.log 2>&1; tail -1 gpurun_out/c27_smoke.log
timeout 85 python -m pytest tests -m gpu -q -x --timeout 80 -k "not benchmark_size and not at_513 and not caffeine and not one_hot" > gpurun_out/c27_pytest.log 2>&1
tail -3 gpurun_out/c27_pytest.log
