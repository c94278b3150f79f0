#!/usr/bin/env python
"""Generates tests/golden/*.json -- golden vectors for the CPU oracle and the CUDA engine.

The reference (qchem-rs) holds no tests or fixtures and its integral crate `molint` is absent
(Cargo.toml:12), so these vectors come from an INDEPENDENT implementation written here:
closed-form Cartesian-Gaussian integrals after Taketa, Huzinaga and O-ohata (J. Phys. Soc. Japan 21,
2313 (1966)) -- binomial-prefactor / B-array sums, no Hermite recursions -- with the Boys function
taken from scipy's confluent hypergeometric function, F_m(T) = 1F1(m+1/2; m+3/2; -T)/(2m+1).
It shares no code and no algorithm with oracle/qc_oracle.cpp (McMurchie-Davidson) or the CUDA
kernels.  Inputs are read from the repository's data/ files (same schema as the reference's).

Run:  python tests/golden/make_golden.py      (about a minute; pure Python)
"""
import itertools
import json
import math
import sys
from pathlib import Path

import numpy as np
from scipy.special import hyp1f1, comb, factorial, factorial2

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import qcpkg  # noqa: E402

qcpkg.load()
from qchem_rs_b200.basis import BasisSet, MolecularSystem, Atom, cart_components, ncart  # noqa: E402


def boys(m, T):
    return hyp1f1(m + 0.5, m + 1.5, -T) / (2 * m + 1)


def fact2(n):
    return 1.0 if n <= 0 else float(factorial2(n))


def norm(alpha, l, m, n):
    return math.sqrt(2 ** (2 * (l + m + n) + 1.5) * alpha ** (l + m + n + 1.5)
                     / fact2(2 * l - 1) / fact2(2 * m - 1) / fact2(2 * n - 1) / math.pi ** 1.5)


def binomial_prefactor(s, ia, ib, xpa, xpb):
    total = 0.0
    for t in range(s + 1):
        if s - ia <= t <= ib:
            total += comb(ia, s - t) * comb(ib, t) * xpa ** (ia - s + t) * xpb ** (ib - t)
    return total


def overlap_1d(l1, l2, pax, pbx, gamma):
    total = 0.0
    for i in range(1 + (l1 + l2) // 2):
        total += binomial_prefactor(2 * i, l1, l2, pax, pbx) * fact2(2 * i - 1) / (2 * gamma) ** i
    return total


def prim_overlap(a1, lmn1, A, a2, lmn2, B):
    if min(lmn2) < 0:
        return 0.0
    g = a1 + a2
    P = (a1 * A + a2 * B) / g
    pre = (math.pi / g) ** 1.5 * math.exp(-a1 * a2 * np.dot(A - B, A - B) / g)
    w = 1.0
    for k in range(3):
        w *= overlap_1d(lmn1[k], lmn2[k], P[k] - A[k], P[k] - B[k], g)
    return pre * w


def prim_kinetic(a1, lmn1, A, a2, lmn2, B):
    l2, m2, n2 = lmn2
    t0 = a2 * (2 * (l2 + m2 + n2) + 3) * prim_overlap(a1, lmn1, A, a2, lmn2, B)
    t1 = -2 * a2 ** 2 * (prim_overlap(a1, lmn1, A, a2, (l2 + 2, m2, n2), B)
                         + prim_overlap(a1, lmn1, A, a2, (l2, m2 + 2, n2), B)
                         + prim_overlap(a1, lmn1, A, a2, (l2, m2, n2 + 2), B))
    t2 = -0.5 * (l2 * (l2 - 1) * prim_overlap(a1, lmn1, A, a2, (l2 - 2, m2, n2), B)
                 + m2 * (m2 - 1) * prim_overlap(a1, lmn1, A, a2, (l2, m2 - 2, n2), B)
                 + n2 * (n2 - 1) * prim_overlap(a1, lmn1, A, a2, (l2, m2, n2 - 2), B))
    return t0 + t1 + t2


def a_array(l1, l2, pa, pb, cp, g):
    imax = l1 + l2 + 1
    arr = [0.0] * imax
    for i in range(imax):
        for r in range(i // 2 + 1):
            for u in range((i - 2 * r) // 2 + 1):
                idx = i - 2 * r - u
                arr[idx] += ((-1) ** i * binomial_prefactor(i, l1, l2, pa, pb) * (-1) ** u * factorial(i)
                             * cp ** (i - 2 * r - 2 * u) * (0.25 / g) ** (r + u)
                             / factorial(r) / factorial(u) / factorial(i - 2 * r - 2 * u))
    return arr


def prim_nuclear(a1, lmn1, A, a2, lmn2, B, C):
    g = a1 + a2
    P = (a1 * A + a2 * B) / g
    rab2 = np.dot(A - B, A - B)
    rcp2 = np.dot(C - P, C - P)
    arrs = [a_array(lmn1[k], lmn2[k], P[k] - A[k], P[k] - B[k], P[k] - C[k], g) for k in range(3)]
    total = 0.0
    for i, ax in enumerate(arrs[0]):
        for j, ay in enumerate(arrs[1]):
            for k, az in enumerate(arrs[2]):
                total += ax * ay * az * boys(i + j + k, rcp2 * g)
    return -2 * math.pi / g * math.exp(-a1 * a2 * rab2 / g) * total


def fact_ratio2(a, b):
    return factorial(a) / factorial(b) / factorial(a - 2 * b)


def b0(i, r, g):
    return fact_ratio2(i, r) * (4 * g) ** (r - i)


def fb(i, l1, l2, p, a, b, r, g):
    return binomial_prefactor(i, l1, l2, p - a, p - b) * b0(i, r, g)


def b_array(l1, l2, l3, l4, p, a, b, q, c, d, g1, g2, delta):
    imax = l1 + l2 + l3 + l4 + 1
    arr = [0.0] * imax
    for i1 in range(l1 + l2 + 1):
        for i2 in range(l3 + l4 + 1):
            for r1 in range(i1 // 2 + 1):
                for r2 in range(i2 // 2 + 1):
                    for u in range((i1 + i2) // 2 - r1 - r2 + 1):
                        idx = i1 + i2 - 2 * (r1 + r2) - u
                        n = i1 + i2 - 2 * (r1 + r2)
                        arr[idx] += (fb(i1, l1, l2, p, a, b, r1, g1) * (-1) ** i2 * fb(i2, l3, l4, q, c, d, r2, g2)
                                     * (-1) ** u * fact_ratio2(n, u) * (q - p) ** (n - 2 * u) / delta ** (n - u))
    return arr


def prim_eri(aa, la, A, ab, lb, B, ac, lc, C, ad, ld, D):
    rab2 = np.dot(A - B, A - B)
    rcd2 = np.dot(C - D, C - D)
    g1, g2 = aa + ab, ac + ad
    P = (aa * A + ab * B) / g1
    Q = (ac * C + ad * D) / g2
    rpq2 = np.dot(P - Q, P - Q)
    delta = 0.25 * (1 / g1 + 1 / g2)
    arrs = [b_array(la[k], lb[k], lc[k], ld[k], P[k], A[k], B[k], Q[k], C[k], D[k], g1, g2, delta) for k in range(3)]
    total = 0.0
    for i, bx in enumerate(arrs[0]):
        for j, by in enumerate(arrs[1]):
            for k, bz in enumerate(arrs[2]):
                total += bx * by * bz * boys(i + j + k, 0.25 * rpq2 / delta)
    return (2 * math.pi ** 2.5 / (g1 * g2 * math.sqrt(g1 + g2)) * math.exp(-aa * ab * rab2 / g1)
            * math.exp(-ac * ad * rcd2 / g2) * total)


class Fn:
    """One contracted Cartesian basis function (tabulated coefficients x per-primitive norm)."""

    def __init__(self, lmn, center, exps, coefs):
        self.lmn, self.A, self.exps = lmn, np.asarray(center, float), list(exps)
        self.c = [c * norm(a, *lmn) for a, c in zip(exps, coefs)]


def functions(system):
    fns = []
    for sh, ia in zip(system.shells, system.shell_atom):
        for lmn in cart_components(sh.l):
            fns.append(Fn(lmn, system.atoms[ia].position, sh.exponents, sh.coefficients))
    return fns


def contracted(fn_list, prim):
    total = 0.0
    for idx in itertools.product(*[range(len(f.exps)) for f in fn_list]):
        w = 1.0
        args = []
        for f, i in zip(fn_list, idx):
            w *= f.c[i]
            args += [f.exps[i], f.lmn, f.A]
        total += w * prim(*args)
    return total


def one_electron(system):
    fns = functions(system)
    n = len(fns)
    S, T, V = np.zeros((n, n)), np.zeros((n, n)), np.zeros((n, n))
    for i in range(n):
        for j in range(i + 1):
            S[i, j] = S[j, i] = contracted([fns[i], fns[j]], prim_overlap)
            T[i, j] = T[j, i] = contracted([fns[i], fns[j]], prim_kinetic)
            v = 0.0
            for at in system.atoms:
                v += at.ordinal * contracted([fns[i], fns[j]],
                                             lambda a1, l1, A, a2, l2, B, C=at.position: prim_nuclear(a1, l1, A, a2, l2, B, C))
            V[i, j] = V[j, i] = v
    return S, T, V


def eri_unique(system):
    """(ij|kl) for i>=j, k>=l, ij>=kl as a flat list in that loop order."""
    fns = functions(system)
    n = len(fns)
    vals = []
    for i in range(n):
        for j in range(i + 1):
            for k in range(i + 1):
                for l in range(k + 1):
                    if k == i and l > j:
                        continue
                    vals.append(contracted([fns[i], fns[j], fns[k], fns[l]], prim_eri))
    return vals


def main():
    out = Path(__file__).resolve().parent
    data = ROOT / "data"

    # 1. H2 / STO-3G  (Szabo-Ostlund table) and H2O / STO-3G: S, T, V and every unique ERI
    for mol in ("hydrogen", "water"):
        bs = BasisSet.load(data / "basis" / "STO-3G.json")
        system = MolecularSystem.load(data / "mol" / f"{mol}.json", bs)
        S, T, V = one_electron(system)
        doc = {"molecule": mol, "basis": "STO-3G", "n_basis": system.n_basis(),
               "overlap": S.tolist(), "kinetic": T.tolist(), "nuclear": V.tolist(),
               "eri_unique_order": "for i: for j<=i: for k<=i: for l<=k: skip (k==i and l>j)",
               "eri_unique": eri_unique(system)}
        (out / f"{mol}_sto3g.json").write_text(json.dumps(doc))
        print(mol, "done", len(doc["eri_unique"]), "unique ERIs")

    # 2. a low-symmetry 3-centre system with s, p and Cartesian d shells: one-electron matrices and
    #    selected shell quartets covering d classes (single primitives, so pure Python stays fast)
    rng = np.random.default_rng(7)
    centers = rng.uniform(-1.2, 1.2, size=(4, 3))
    shells = []          # (l, center index, exponent)
    for l in (0, 1, 2):
        for c in range(4):
            shells.append((l, c, float(rng.uniform(0.4, 1.8))))
    quartets = [(8, 9, 10, 11), (8, 4, 9, 0), (8, 8, 9, 9), (10, 5, 6, 1), (11, 0, 1, 2), (9, 10, 4, 5),
                (4, 5, 6, 7), (4, 0, 5, 1), (8, 1, 2, 3), (11, 10, 9, 3), (8, 9, 4, 1), (10, 6, 11, 7),
                (4, 4, 4, 4), (8, 8, 8, 8), (9, 5, 9, 5), (0, 1, 2, 3)]
    blocks = []
    for q in quartets:
        comps = [cart_components(shells[s][0]) for s in q]
        blk = np.zeros([len(c) for c in comps])
        for idx in itertools.product(*[range(len(c)) for c in comps]):
            fl = [Fn(comps[k][idx[k]], centers[shells[q[k]][1]], [shells[q[k]][2]], [1.0]) for k in range(4)]
            blk[idx] = contracted(fl, prim_eri)
        blocks.append({"shells": list(q), "values": blk.tolist()})
        print("quartet", q, "done")
    atoms = [Atom(z, centers[i]) for i, z in enumerate((8, 1, 6, 7))]
    sysd = MolecularSystem(atoms)
    from qchem_rs_b200.basis import Shell
    for l, c, a in shells:
        sysd.shells.append(Shell(l, np.array([a]), np.array([1.0]), "gto_cartesian"))
        sysd.shell_atom.append(c)
    S, T, V = one_electron(sysd)
    doc = {"centers": centers.tolist(), "Z": [8, 1, 6, 7], "shells": [list(s) for s in shells],
           "overlap": S.tolist(), "kinetic": T.tolist(), "nuclear": V.tolist(), "quartets": blocks}
    (out / "spd_random.json").write_text(json.dumps(doc))

    # 3. Boys function table
    Ts = [0.0, 1e-9, 1e-3, 0.1, 0.5, 1.0, 3.3, 7.9, 12.0, 18.5, 25.0, 30.0, 33.0, 35.9, 36.1, 40.0, 55.0, 80.0, 150.0, 1e3, 1e5]
    import mpmath
    mpmath.mp.dps = 40
    tab = []
    for T in Ts:
        row = []
        for m in range(0, 13):
            if T == 0:
                row.append(1.0 / (2 * m + 1))
            else:
                row.append(float(mpmath.hyp1f1(m + 0.5, m + 1.5, -T) / (2 * m + 1)))
        tab.append(row)
    (out / "boys.json").write_text(json.dumps({"T": Ts, "F": tab}))
    print("boys done")


if __name__ == "__main__":
    main()
