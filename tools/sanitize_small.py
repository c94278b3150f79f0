"""Smallest end-to-end case for compute-sanitizer: (H2O)_2 / 6-31G* (all 21 classes), one RHF and one UHF build."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
import qcpkg
pkg = qcpkg.load()
bs = pkg.BasisSet.load(ROOT / "data" / "basis" / "6-31G_st.json")
system = pkg.MolecularSystem.from_atoms(pkg.molecules.water_cluster(2), bs)
n = system.n_basis()
rng = np.random.default_rng(0)
a = rng.normal(size=(n, n)); P = 0.5 * (a + a.T)
with pkg.engine.FockEngine(system, tau=1e-12) as eng:
    S, T, V = eng.one_electron()
    G = eng.rhf(P)
    Ga, Gb = eng.uhf(P, 0.5 * P)
    print("ok", float(np.abs(G).max()), float(np.abs(Ga).max()), eng.stats()["quartets"])
