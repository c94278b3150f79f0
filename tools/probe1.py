import sys, time, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from helpers import water_cluster, oracle_lib, load_system
from qchem_rs_b200 import hf, engine

ns = [int(x) for x in sys.argv[1].split(',')]
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
for n in ns:
    system = water_cluster(n)
    fb = system.flat()
    t0 = time.time(); ints = oracle_lib.one_electron(fb); t1 = time.time()
    print(f"n={n} N={fb.n_basis} 1e ints {t1-t0:.2f}s", flush=True)
    with engine.FockEngine(system, tau=1e-12) as eng:
        if n == ns[0]:
            print("fp64 peak TF", eng.fp64_peak_tflops())
        class B:
            def rhf(self, P):
                G = eng.rhf(P); st = eng.stats()
                print(f"  build kernel_ms={st['kernel_ms']:.2f} total_ms={st['total_ms']:.2f} quartets={st['quartets']:.3e}/{st['quartets_total']:.3e} "
                      f"model_TF={st['model_flops']/st['kernel_ms']/1e9:.3f} launches={st['launches']} pairs={st['n_pairs']} groups={st['n_groups']}", flush=True)
                return G
        out = hf.restricted_hartree_fock(system, hf.HartreeFockConfig(iters, 1e-8), ints, B())
