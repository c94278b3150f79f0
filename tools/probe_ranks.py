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

n = int(sys.argv[1]); world = int(sys.argv[2])
system = water_cluster(n)
with engine.FockEngine(system, tau=1e-12) as eng:
    ints = eng.one_electron()
    seen = {}
    class Tap:
        def rhf(self, P):
            seen['P'] = P.copy(); return eng.rhf(P)
    hf.restricted_hartree_fock(system, hf.HartreeFockConfig(3, 1e-14), ints, Tap())
    P = seen['P']
    for _ in range(2): eng.rhf(P)
    st = eng.stats(); print("full", st['kernel_ms'], st['total_ms'], st['quartets'])
for r in range(world):
    with engine.FockEngine(system, tau=1e-12, rank=r, world_size=world) as eng:
        for _ in range(3): eng.rhf(P)
        st = eng.stats(); print("rank", r, f"kernel_ms={st['kernel_ms']:.2f} total_ms={st['total_ms']:.2f} q={st['quartets']:.3e} flops={st['model_flops']:.3e}")
