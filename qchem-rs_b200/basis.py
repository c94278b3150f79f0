"""Basis-set and molecule loaders: the host-side mirror of the two `molint` loader calls the
reference makes before an SCF run.

Reference call sites (the `molint` crate itself is an un-vendored path dependency,
`Cargo.toml:12`, so only the call sites exist):

* ``BasisSet::load(path)``                 -> qchem-cli/src/main.rs:76, :120
* ``MolecularSystem::load(path, &basis)``  -> qchem-cli/src/main.rs:77, :121
* ``system.atoms[i].ordinal / .position``  -> core/src/hf/rhf.rs:36, :116-117
* ``system.n_basis()``                     -> core/src/hf/rhf.rs:37

File formats:

* basis: MolSSI-BSE "complete" JSON; ``elements[Z].electron_shells[*]`` with string-encoded
  ``exponents`` and ``coefficients[k]`` (one row per entry of ``angular_momentum``; fused SP shells
  have ``angular_momentum: [0, 1]``).
* molecule: JSON array of ``{"element": "<Z>", "position": [x, y, z]}``; positions are used as
  bohr with no unit conversion (rhf.rs:116-117 feeds them straight into 1/r).

Conventions this engine fixes (molint is absent, so they are choices, documented in DESIGN.md):

* fused SP shells are split into one s shell followed by one p shell, file order kept;
* basis functions are atom-major, shells in file order, Cartesian components in the order
  x,y,z / xx,xy,xz,yy,yz,zz (lexicographic, 6d);
* every primitive Cartesian Gaussian x^i y^j z^k exp(-a r^2) carries its own normalisation
  N(a;i,j,k); contraction coefficients are used as tabulated (no renormalisation of the contraction);
* the flat ``coefs`` array handed to the engine holds  c_k * N(a_k; l,0,0); the remaining
  per-component factor sqrt((2l-1)!!/((2i-1)!!(2j-1)!!(2k-1)!!)) is applied by the engine.
"""
from __future__ import annotations

import ctypes
import json
import math
from dataclasses import dataclass, field
from pathlib import Path
from typing import List, Sequence

import numpy as np

MAX_L = 2  # the CUDA engine has s, p and (Cartesian) d kernels


def ncart(l: int) -> int:
    return (l + 1) * (l + 2) // 2


def dfact(n: int) -> float:
    """(n)!! with (-1)!! = 1."""
    r = 1.0
    while n > 1:
        r *= n
        n -= 2
    return r


def prim_norm(alpha: float, l: int) -> float:
    """Normalisation of x^l exp(-a r^2)."""
    return (2.0 * alpha / math.pi) ** 0.75 * (4.0 * alpha) ** (l / 2.0) / math.sqrt(dfact(2 * l - 1))


def cart_components(l: int):
    return [(i, j, l - i - j) for i in range(l, -1, -1) for j in range(l - i, -1, -1)]


def component_scale(l: int) -> List[float]:
    """N(a;i,j,k) / N(a;l,0,0) for each Cartesian component, independent of a."""
    return [math.sqrt(dfact(2 * l - 1) / (dfact(2 * i - 1) * dfact(2 * j - 1) * dfact(2 * k - 1)))
            for (i, j, k) in cart_components(l)]


@dataclass
class Shell:
    l: int
    exponents: np.ndarray      # (K,)
    coefficients: np.ndarray   # (K,) as tabulated
    function_type: str = "gto"


@dataclass
class BasisSet:
    """Mirror of `molint::basis::BasisSet` as far as the reference uses it (main.rs:76)."""
    name: str
    elements: dict  # Z (int) -> list[Shell], SP shells already split

    @staticmethod
    def load(path) -> "BasisSet":
        with open(path) as fh:
            doc = json.load(fh)
        elements = {}
        for z, entry in doc["elements"].items():
            shells: List[Shell] = []
            for sh in entry.get("electron_shells", []):
                exps = np.array([float(x) for x in sh["exponents"]], dtype=np.float64)
                ftype = sh.get("function_type", "gto")
                for l, row in zip(sh["angular_momentum"], sh["coefficients"]):
                    coef = np.array([float(x) for x in row], dtype=np.float64)
                    keep = coef != 0.0
                    shells.append(Shell(int(l), exps[keep].copy(), coef[keep].copy(), ftype))
            elements[int(z)] = shells
        return BasisSet(doc.get("name", Path(path).stem), elements)


@dataclass
class Atom:
    """Mirror of `molint::system::Atom` (fields used at rhf.rs:36, :116-117)."""
    ordinal: int
    position: np.ndarray  # (3,) bohr


@dataclass
class MolecularSystem:
    """Mirror of `molint::system::MolecularSystem` (rhf.rs:36-37, main.rs:77)."""
    atoms: List[Atom]
    shells: List[Shell] = field(default_factory=list)
    shell_atom: List[int] = field(default_factory=list)

    @staticmethod
    def load(path, basis: BasisSet) -> "MolecularSystem":
        with open(path) as fh:
            doc = json.load(fh)
        atoms = [Atom(int(a["element"]), np.array(a["position"], dtype=np.float64)) for a in doc]
        return MolecularSystem.from_atoms(atoms, basis)

    @staticmethod
    def from_atoms(atoms: Sequence[Atom], basis: BasisSet) -> "MolecularSystem":
        sys_ = MolecularSystem(list(atoms))
        for ia, atom in enumerate(sys_.atoms):
            if atom.ordinal not in basis.elements:
                raise KeyError(f"basis set {basis.name} has no element Z={atom.ordinal}")
            for sh in basis.elements[atom.ordinal]:
                if sh.l > MAX_L:
                    raise ValueError(f"angular momentum l={sh.l} > {MAX_L} is not supported")
                if sh.l >= 2 and sh.function_type == "gto_spherical":
                    raise ValueError("spherical d shells are not supported (Cartesian 6d only)")
                sys_.shells.append(sh)
                sys_.shell_atom.append(ia)
        return sys_

    # -- derived sizes -------------------------------------------------------------------------
    def n_basis(self) -> int:
        return sum(ncart(s.l) for s in self.shells)

    def n_electrons(self) -> int:
        return sum(a.ordinal for a in self.atoms)   # rhf.rs:36 (neutral molecules only)

    def shell_offsets(self) -> np.ndarray:
        off = np.zeros(len(self.shells) + 1, dtype=np.int64)
        for i, s in enumerate(self.shells):
            off[i + 1] = off[i] + ncart(s.l)
        return off

    def flat(self) -> "FlatBasis":
        return FlatBasis(self)


class CBasis(ctypes.Structure):
    """ctypes image of `qcf_basis` (include/qcfock.h); the oracle uses the same layout."""
    _fields_ = [
        ("n_atoms", ctypes.c_int),
        ("Z", ctypes.POINTER(ctypes.c_int)),
        ("xyz", ctypes.POINTER(ctypes.c_double)),
        ("n_shells", ctypes.c_int),
        ("shell_atom", ctypes.POINTER(ctypes.c_int)),
        ("shell_l", ctypes.POINTER(ctypes.c_int)),
        ("shell_nprim", ctypes.POINTER(ctypes.c_int)),
        ("shell_prim_off", ctypes.POINTER(ctypes.c_int)),
        ("exps", ctypes.POINTER(ctypes.c_double)),
        ("coefs", ctypes.POINTER(ctypes.c_double)),
        ("cartesian", ctypes.c_int),
    ]


class FlatBasis:
    """Flat arrays in the layout of `qcf_basis`; keeps the numpy buffers alive."""

    def __init__(self, system: MolecularSystem):
        self.Z = np.array([a.ordinal for a in system.atoms], dtype=np.int32)
        self.xyz = np.ascontiguousarray(
            np.array([a.position for a in system.atoms], dtype=np.float64).reshape(-1))
        self.shell_atom = np.array(system.shell_atom, dtype=np.int32)
        self.shell_l = np.array([s.l for s in system.shells], dtype=np.int32)
        self.shell_nprim = np.array([len(s.exponents) for s in system.shells], dtype=np.int32)
        off = np.zeros(len(system.shells), dtype=np.int32)
        if len(off) > 1:
            off[1:] = np.cumsum(self.shell_nprim)[:-1]
        self.shell_prim_off = off
        self.exps = np.concatenate([s.exponents for s in system.shells]).astype(np.float64)
        self.coefs = np.concatenate([
            s.coefficients * np.array([prim_norm(a, s.l) for a in s.exponents])
            for s in system.shells]).astype(np.float64)
        self.n_basis = system.n_basis()
        ip = ctypes.POINTER(ctypes.c_int)
        dp = ctypes.POINTER(ctypes.c_double)
        self.c = CBasis(
            len(self.Z), self.Z.ctypes.data_as(ip), self.xyz.ctypes.data_as(dp),
            len(self.shell_l), self.shell_atom.ctypes.data_as(ip), self.shell_l.ctypes.data_as(ip),
            self.shell_nprim.ctypes.data_as(ip), self.shell_prim_off.ctypes.data_as(ip),
            self.exps.ctypes.data_as(dp), self.coefs.ctypes.data_as(dp), 1)

    def ref(self):
        return ctypes.byref(self.c)
