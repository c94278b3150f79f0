"""UHF Fock build (alpha/beta exchange digestion, NK = 2 kernels) on the N = 1007 water cluster: device time per build."""
import sys, os, json
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
import qcpkg
pkg = qcpkg.load()
from qchem_rs_b200 import hf, engine, molecules
from qchem_rs_b200.basis import BasisSet, MolecularSystem
n = int(sys.argv[1]) if len(sys.argv) > 1 else 53
bs = BasisSet.load(ROOT / "data" / "basis" / "6-31G_st.json")
system = MolecularSystem.from_atoms(molecules.water_cluster(n), bs)
with engine.FockEngine(system, tau=1e-12) as eng:
    ints = eng.one_electron()
    seen = {}
    class Tap:
        def rhf(self, P):
            seen['P'] = P.copy(); return eng.rhf(P)
    hf.restricted_hartree_fock(system, hf.HartreeFockConfig(6, 1e-14), ints, Tap())
    P = seen['P']
    t_r, t_u = [], []
    for _ in range(4):
        g = eng.rhf(P); t_r.append(eng.stats()['kernel_ms'])
    for _ in range(4):
        ga, gb = eng.uhf(0.5 * P, 0.5 * P); st = eng.stats(); t_u.append(st['kernel_ms'])
    err = float(np.max(np.abs(ga - g)))     # closed shell: G_alpha = J[P] - K[P/2] = G_rhf
    print(json.dumps({"workload": f"(H2O)_{n} 6-31G* N={st['n_basis']}", "rhf_build_ms": min(t_r[1:]), "uhf_build_ms": min(t_u[1:]),
                      "quartets": st['quartets'], "max_abs_Ga_minus_Grhf": err}))
