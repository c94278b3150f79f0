import os
os.environ["QCF_PROFILE"] = "1"
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

system = water_cluster(53)
P = np.load('/tmp/P.npy')
res = {}
for w in (1, 8):
    with engine.FockEngine(system, tau=1e-12, rank=0, world_size=w) as eng:
        for _ in range(2): eng.rhf(P)
        res[w] = {(r['la'], r['lb'], r['kab'], r['lc'], r['ld'], r['kcd']): r for r in eng.launch_profile()}
tot1 = sum(r['ms'] for r in res[1].values()); tot8 = sum(r['ms'] for r in res[8].values())
print("serialized totals", tot1, tot8, "ideal8", tot1 / 8)
rows = []
for k, r1 in res[1].items():
    r8 = res[8].get(k)
    if r8: rows.append((r8['ms'] - r1['ms'] / 8, k, r1['ms'], r8['ms'], r1['nbra'], r1['nket']))
rows.sort(reverse=True)
print("excess_ms key ms1 ms8 nbra nket")
for e, k, m1, m8, nb, nk in rows[:40]:
    print(f"{e:7.3f} ({k[0]}{k[1]}|{k[3]}{k[4]}) K={k[2]}x{k[5]} ms1={m1:7.3f} ms8={m8:7.3f} ratio={m1/max(m8,1e-9):5.2f} nbra={nb} nket={nk}")
print("sum excess", sum(r[0] for r in rows))
