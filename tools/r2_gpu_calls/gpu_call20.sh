#!/bin/bash
# GPU call 20: final code -- full GPU suite, smoke, device SCF to epsilon = 1e-8 with incremental builds (three switch-over
# thresholds), ncu --set full of the slab kernel launches (d-bra classes)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/c20_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c20_pytest.log
tail -4 gpurun_out/c20_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/c20_smoke.log 2>&1; tail -2 gpurun_out/c20_smoke.log
timeout 400 python bench.py --scf --scf-epsilon 1e-8 --no-cpu-baseline > gpurun_out/c20_bench_scf_eps1e-8.json 2> gpurun_out/c20_bench_scf.err
QCF_INC_RMS=1e-3 timeout 400 python bench.py --scf --scf-epsilon 1e-8 --no-cpu-baseline > gpurun_out/c20_bench_scf_eps1e-8_incrms1e-3.json 2>> gpurun_out/c20_bench_scf.err
QCF_INC_RMS=1e-2 timeout 400 python bench.py --scf --scf-epsilon 1e-8 --scf-full-every 12 --no-cpu-baseline > gpurun_out/c20_bench_scf_eps1e-8_incrms1e-2_every12.json 2>> gpurun_out/c20_bench_scf.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/c20_bench_scf_eps*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f, "ms_per_step", round(d["ms_per_step"], 2))
    for k, v in d["config"]["scf"].items():
        if isinstance(v, dict):
            print("  ", k, "it", v["iterations"], "steps_wall_s", round(v["steps_wall_s"], 3), "E", v["e_total"], "last builds", v["build_ms"][-10:])
PY
QCF_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --kernel-name-base mangled \
  -k regex:'eri_jk_slab_kernel' -c 22 -o /tmp/r2_prof_slab_final python tools/ab.py 53 1 > gpurun_out/c20_ncu_slab.log 2>&1
tail -2 gpurun_out/c20_ncu_slab.log
python tools/ncu_summary.py /tmp/r2_prof_slab_final.ncu-rep > gpurun_out/r2_final_slab_kernels.txt 2>&1
cut -c1-400 gpurun_out/r2_final_slab_kernels.txt | head -30
ls -la /tmp/*.ncu-rep
