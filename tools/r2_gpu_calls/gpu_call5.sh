#!/bin/bash
# GPU call 5: knob sweep on the restored build, full GPU test-suite, bench line, ncu launch list + top-kernel capture
mkdir -p gpurun_out
: > gpurun_out/c5_ab.log
AB_TAG=default timeout 600 python tools/ab.py 53 5 >> gpurun_out/c5_ab.log 2>&1
AB_TAG=copy1 QCF_COPY_THREADS=1 timeout 600 python tools/ab.py 53 5 >> gpurun_out/c5_ab.log 2>&1
AB_TAG=block128 QCF_BLOCK=128 timeout 600 python tools/ab.py 53 4 >> gpurun_out/c5_ab.log 2>&1
AB_TAG=block32 QCF_BLOCK=32 timeout 600 python tools/ab.py 53 4 >> gpurun_out/c5_ab.log 2>&1
AB_TAG=order1 QCF_ORDER=1 timeout 600 python tools/ab.py 53 4 >> gpurun_out/c5_ab.log 2>&1
AB_TAG=kpt128 QCF_KETS_PER_THREAD=128 timeout 600 python tools/ab.py 53 4 >> gpurun_out/c5_ab.log 2>&1
AB_TAG=kpt96_ctas148 QCF_KETS_PER_THREAD=96 QCF_TARGET_CTAS=148 timeout 600 python tools/ab.py 53 4 >> gpurun_out/c5_ab.log 2>&1
AB_TAG=n27 timeout 600 python tools/ab.py 27 4 >> gpurun_out/c5_ab.log 2>&1
AB_TAG=n105 timeout 900 python tools/ab.py 105 3 >> gpurun_out/c5_ab.log 2>&1
cat gpurun_out/c5_ab.log
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/c5_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c5_pytest.log
tail -4 gpurun_out/c5_pytest.log
timeout 900 python bench.py --scf > gpurun_out/c5_bench_n1.json 2> gpurun_out/c5_bench_n1.err
tail -c 600 gpurun_out/c5_bench_n1.json
timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/c5_plain.log 2>&1 && \
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2_launches_bench.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/c5_ncu_list.log 2>&1
QCF_NO_GRAPH=1 timeout 600 python tools/ab.py 53 1 > gpurun_out/c5_ncu_plain2.log 2>&1 && \
QCF_NO_GRAPH=1 timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name-base mangled \
  -k regex:'eri_jk_kernelILi1ELi0ELi0ELi0ELi1ELi1E' -s 20 -c 10 -o gpurun_out/r2_prof_block1000 python tools/ab.py 53 1 > gpurun_out/c5_ncu_full.log 2>&1
tail -2 gpurun_out/c5_ncu_full.log
