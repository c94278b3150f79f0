#!/bin/bash
# GPU call 16: final committed state -- full GPU suite, smoke, bench line, BASELINE molecular configs, UHF build time
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/c16_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c16_pytest.log
tail -4 gpurun_out/c16_pytest.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/c16_smoke.log 2>&1; tail -1 gpurun_out/c16_smoke.log
timeout 900 python bench.py --scf > gpurun_out/c16_bench_n1.json 2> gpurun_out/c16_bench_n1.err
tail -c 200 gpurun_out/c16_bench_n1.json
timeout 900 python tools/config_times.py > gpurun_out/c16_baseline_configs.jsonl 2> gpurun_out/c16_baseline_configs.err
cat gpurun_out/c16_baseline_configs.jsonl | cut -c1-330
timeout 600 python tools/uhf_time.py > gpurun_out/c16_uhf_time.json 2> gpurun_out/c16_uhf_time.err
cat gpurun_out/c16_uhf_time.json | cut -c1-400
