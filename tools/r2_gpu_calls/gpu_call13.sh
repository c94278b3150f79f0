#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -q --timeout 600 -k "one_hot or cli_driver or incremental or device_resident" > gpurun_out/c13_pytest.log 2>&1
tail -4 gpurun_out/c13_pytest.log
timeout 600 python - > gpurun_out/c13_scf.log 2>&1 <<'PY'
import sys, time
sys.path.insert(0, '.')
import qcpkg; pkg = qcpkg.load()
bs = pkg.BasisSet.load('data/basis/6-31G_st.json')
system = pkg.MolecularSystem.from_atoms(pkg.molecules.water_cluster(53), bs)
with pkg.engine.FockEngine(system, tau=1e-12) as eng:
    ints = eng.one_electron()
    cfg = pkg.hf.HartreeFockConfig(60, 1e-6)
    for every in (0, 8, 0, 8):
        t0 = time.perf_counter(); out = pkg.hf.restricted_hartree_fock_device(system, cfg, ints, eng, full_rebuild_every=every); dt = time.perf_counter() - t0
        print(f"full_every={every} iterations={out.iterations} wall={dt:.3f} init={out.init_s:.3f} steps_wall={sum(s['wall_ms'] for s in out.steps)*1e-3:.3f} "
              f"build={sum(s['build_ms'] for s in out.steps)*1e-3:.3f} linalg={sum(s['linalg_ms'] for s in out.steps)*1e-3:.3f} E={out.total_energy():.9f}", flush=True)
        print('   wall_ms per step', [round(s['wall_ms']) for s in out.steps])
PY
cat gpurun_out/c13_scf.log | cut -c1-400
