"""Shared test helpers: system construction and the oracle-backed Fock builders."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import qcpkg  # noqa: E402

qcpkg.load()
from qchem_rs_b200.basis import BasisSet, MolecularSystem, Atom, Shell  # noqa: E402
from qchem_rs_b200 import molecules  # noqa: E402
from oracle import oracle_lib  # noqa: E402

DATA = ROOT / "data"
GOLD = ROOT / "tests" / "golden"


def load_system(mol: str, basis: str) -> MolecularSystem:
    bs = BasisSet.load(DATA / "basis" / f"{basis}.json")
    return MolecularSystem.load(DATA / "mol" / f"{mol}.json", bs)


def water_cluster(n: int, basis: str = "6-31G_st") -> MolecularSystem:
    bs = BasisSet.load(DATA / "basis" / f"{basis}.json")
    return MolecularSystem.from_atoms(molecules.water_cluster(n), bs)


def spd_random_system():
    """The 4-centre s/p/d single-primitive system of tests/golden/spd_random.json."""
    doc = json.loads((GOLD / "spd_random.json").read_text())
    atoms = [Atom(z, np.array(c)) for z, c in zip(doc["Z"], doc["centers"])]
    system = MolecularSystem(atoms)
    for l, c, a in doc["shells"]:
        system.shells.append(Shell(int(l), np.array([a]), np.array([1.0]), "gto_cartesian"))
        system.shell_atom.append(int(c))
    return system, doc


def random_symmetric_density(n: int, seed: int, scale: float = 1.0) -> np.ndarray:
    rng = np.random.default_rng(seed)
    a = rng.normal(size=(n, n)) * scale
    return 0.5 * (a + a.T)
