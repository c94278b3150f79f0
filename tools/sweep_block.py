"""Sweep of the block kernel's CTA size and ket chunking on the N = 1007 build (no rebuild needed)."""
import sys, os
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
import qcpkg
pkg = qcpkg.load()
from qchem_rs_b200 import hf, engine, molecules
from qchem_rs_b200.basis import BasisSet, MolecularSystem
bs = BasisSet.load(ROOT / "data" / "basis" / "6-31G_st.json")
system = MolecularSystem.from_atoms(molecules.water_cluster(53), bs)
with engine.FockEngine(system, tau=1e-12) as eng:
    ints = eng.one_electron()
    seen = {}
    class Tap:
        def rhf(self, P):
            seen['P'] = P.copy(); return eng.rhf(P)
    hf.restricted_hartree_fock(system, hf.HartreeFockConfig(6, 1e-14), ints, Tap())
P = seen['P']
for block in (32, 64, 128):
    for kpt in (16, 32, 64):
        os.environ["QCF_KETS_PER_THREAD"] = str(kpt)
        with engine.FockEngine(system, tau=1e-12, block_threads=block) as eng:
            t = []
            for _ in range(4):
                eng.rhf(P); t.append(eng.stats()['kernel_ms'])
            print(f"block={block} kpt={kpt} kernel_ms={min(t[1:]):.2f}", flush=True)
