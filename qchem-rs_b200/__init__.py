"""qcfock: B200-native direct-SCF Fock-build engine for qchem-rs (host-side Python mirror).

The directory name carries a hyphen, so import it through `qcpkg.load()` at the repo root (it
registers the package as `qchem_rs_b200`).
"""
from . import basis, diis, distributed, engine, hf, molecules  # noqa: F401
from .basis import BasisSet, MolecularSystem, Atom  # noqa: F401
from .hf import (HartreeFockConfig, restricted_hartree_fock, unrestricted_hartree_fock)  # noqa: F401

