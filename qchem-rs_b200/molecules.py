"""Inputs the BASELINE configs need but the reference's data/mol/ does not hold:
O2, caffeine and (H2O)_n clusters, emitted in the reference's molecule JSON format
(`[{"element": "<Z>", "position": [x, y, z]}, ...]`, bohr; data/mol/water.json).
"""
from __future__ import annotations

import json
import math
from pathlib import Path

import numpy as np

from .basis import Atom

ANGSTROM = 1.8897261246257702  # bohr per angstrom


def to_json(atoms, path=None) -> str:
    doc = [{"element": str(a.ordinal), "position": [float(x) for x in a.position]} for a in atoms]
    text = json.dumps(doc, indent=1)
    if path is not None:
        Path(path).write_text(text)
    return text


def oxygen():
    """O2, O-O 2.2828 bohr on z."""
    return [Atom(8, np.array([0.0, 0.0, 0.0])), Atom(8, np.array([0.0, 0.0, 2.2828]))]


def caffeine():
    """1,3,7-trimethylxanthine, idealised planar ring geometry (regular hexagon fused with a regular
    pentagon, bond 1.39 A; C=O 1.22, N-CH3 1.46, C-H 1.09 A), 24 atoms, bohr."""
    R = 1.39
    def d(deg): return np.array([math.cos(math.radians(deg)), math.sin(math.radians(deg)), 0.0])
    ring = {"C5": R * d(30), "C6": R * d(90), "N1": R * d(150), "C2": R * d(210), "N3": R * d(270), "C4": R * d(330)}
    x0 = R * math.cos(math.radians(30))
    apo = R / (2 * math.tan(math.radians(36)))
    rc = R / (2 * math.sin(math.radians(36)))
    cen = np.array([x0 + apo, 0.0, 0.0])
    ring.update({"N7": cen + rc * d(72), "C8": cen + rc * d(0), "N9": cen + rc * d(-72)})
    atoms = [(6 if k[0] == "C" else 7, v) for k, v in ring.items()]
    atoms.append((8, ring["C6"] + 1.22 * d(90)))
    atoms.append((8, ring["C2"] + 1.22 * d(210)))
    atoms.append((1, ring["C8"] + 1.09 * d(0)))
    z = np.array([0.0, 0.0, 1.0])
    for name, ang in (("N1", 150), ("N3", 270), ("N7", 72)):
        u = d(ang)
        cm = ring[name] + 1.46 * u
        atoms.append((6, cm))
        e2 = np.cross(u, z)
        for phi in (90.0, 210.0, 330.0):
            hdir = (1.0 / 3.0) * u + (2.0 * math.sqrt(2.0) / 3.0) * (
                math.cos(math.radians(phi)) * z + math.sin(math.radians(phi)) * e2)
            atoms.append((1, cm + 1.09 * hdir))
    return [Atom(zz, np.asarray(p) * ANGSTROM) for zz, p in atoms]


def water_cluster(n: int):
    """(H2O)_n exactly as SURVEY.md 8d specifies: r(OH) = 1.8088 bohr, HOH = 104.52 deg, molecules on
    a simple-cubic lattice a = 5.86 bohr filled lexicographically in a ceil(n^(1/3))^3 box, a uniform
    random rotation (unit quaternion) and a uniform +-0.2 bohr jitter per molecule,
    numpy.random.default_rng(seed = 20240 + n)."""
    rng = np.random.default_rng(20240 + n)
    roh, ang, a = 1.8088, math.radians(104.52), 5.86
    m = int(math.ceil(n ** (1.0 / 3.0) - 1e-9))
    while m ** 3 < n:
        m += 1
    local = np.array([[0.0, 0.0, 0.0],
                      [roh * math.sin(ang / 2), 0.0, roh * math.cos(ang / 2)],
                      [-roh * math.sin(ang / 2), 0.0, roh * math.cos(ang / 2)]])
    atoms = []
    count = 0
    for ix in range(m):
        for iy in range(m):
            for iz in range(m):
                if count >= n:
                    break
                q = rng.normal(size=4)
                q /= np.linalg.norm(q)
                w, x, y, zq = q
                rot = np.array([
                    [1 - 2 * (y * y + zq * zq), 2 * (x * y - zq * w), 2 * (x * zq + y * w)],
                    [2 * (x * y + zq * w), 1 - 2 * (x * x + zq * zq), 2 * (y * zq - x * w)],
                    [2 * (x * zq - y * w), 2 * (y * zq + x * w), 1 - 2 * (x * x + y * y)]])
                shift = np.array([ix, iy, iz], dtype=float) * a + rng.uniform(-0.2, 0.2, size=3)
                pos = local @ rot.T + shift
                atoms.append(Atom(8, pos[0]))
                atoms.append(Atom(1, pos[1]))
                atoms.append(Atom(1, pos[2]))
                count += 1
    return atoms
