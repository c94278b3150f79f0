import time
import sys, os
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
import qcpkg
pkg = qcpkg.load()
from qchem_rs_b200 import hf, engine, molecules
from qchem_rs_b200.basis import BasisSet, MolecularSystem


def water_cluster(n, basis="6-31G_st"):
    bs = BasisSet.load(ROOT / "data" / "basis" / f"{basis}.json")
    return MolecularSystem.from_atoms(molecules.water_cluster(n), bs)

ns = [int(x) for x in sys.argv[1].split(',')]
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
for n in ns:
    system = water_cluster(n)
    fb = system.flat()
    print(f"n={n} N={fb.n_basis}", flush=True)
    with engine.FockEngine(system, tau=1e-12) as eng:
        ints = eng.one_electron()
        if n == ns[0]:
            print("fp64 peak TF", eng.fp64_peak_tflops())
        class B:
            def rhf(self, P):
                G = eng.rhf(P); st = eng.stats()
                print(f"  build kernel_ms={st['kernel_ms']:.2f} total_ms={st['total_ms']:.2f} quartets={st['quartets']:.3e}/{st['quartets_total']:.3e} "
                      f"model_TF={st['model_flops']/st['kernel_ms']/1e9:.3f} launches={st['launches']} pairs={st['n_pairs']} groups={st['n_groups']}", flush=True)
                return G
        out = hf.restricted_hartree_fock(system, hf.HartreeFockConfig(iters, 1e-8), ints, B())
