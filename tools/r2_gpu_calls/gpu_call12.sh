#!/bin/bash
# GPU call 12: final code -- full GPU suite, smoke, bench line (+ device SCF), ncu launch list, top-kernel capture, other sizes
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/c12_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c12_pytest.log
tail -5 gpurun_out/c12_pytest.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/c12_smoke.log 2>&1; tail -2 gpurun_out/c12_smoke.log
timeout 900 python bench.py --scf > gpurun_out/c12_bench_n1.json 2> gpurun_out/c12_bench_n1.err
tail -c 300 gpurun_out/c12_bench_n1.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/c12_bench_reference.json 2> gpurun_out/c12_bench_reference.err
tail -c 400 gpurun_out/c12_bench_reference.json
: > gpurun_out/c12_ab.log
AB_TAG=final_n53 timeout 600 python tools/ab.py 53 5 >> gpurun_out/c12_ab.log 2>&1
AB_TAG=final_det QCF_DETERMINISTIC=1 timeout 600 python tools/ab.py 53 5 >> gpurun_out/c12_ab.log 2>&1
AB_TAG=final_n27 timeout 600 python tools/ab.py 27 4 >> gpurun_out/c12_ab.log 2>&1
AB_TAG=final_n105 timeout 900 python tools/ab.py 105 3 >> gpurun_out/c12_ab.log 2>&1
AB_TAG=final_n158 timeout 1200 python tools/ab.py 158 2 >> gpurun_out/c12_ab.log 2>&1
AB_TAG=final_profile QCF_PROFILE=1 timeout 600 python tools/ab.py 53 2 250 > gpurun_out/c12_profile_final.log 2>&1
cat gpurun_out/c12_ab.log | cut -c1-200
timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/c12_plain.log 2>&1 && \
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2_launches_bench.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/c12_ncu_list.log 2>&1
QCF_NO_GRAPH=1 timeout 600 python tools/ab.py 53 1 > gpurun_out/c12_ncu_plain2.log 2>&1 && \
QCF_NO_GRAPH=1 timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name-base mangled \
  -k regex:'eri_jk_kernelILi1ELi0ELi0ELi0ELi1ELi1E' -s 20 -c 10 -o gpurun_out/r2_prof_block1000_final python tools/ab.py 53 1 > gpurun_out/c12_ncu_full.log 2>&1
tail -2 gpurun_out/c12_ncu_full.log
