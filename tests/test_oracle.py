"""Pins the CPU oracle (oracle/qc_oracle.cpp).  The reference has no tests or golden vectors
(SURVEY.md 4, 8c), so the pins are: the Szabo-Ostlund H2/STO-3G table, golden vectors from the
independent closed-form implementation in tests/golden/make_golden.py, and invariances."""
import json

import numpy as np
import pytest

from helpers import GOLD, load_system, spd_random_system, random_symmetric_density, oracle_lib
from qchem_rs_b200 import hf
from qchem_rs_b200.basis import ncart


def unique_eri(eri):
    n = eri.shape[0]
    vals = []
    for i in range(n):
        for j in range(i + 1):
            for k in range(i + 1):
                for l in range(k + 1):
                    if k == i and l > j:
                        continue
                    vals.append(eri[i, j, k, l])
    return np.array(vals)


def test_boys_against_mpmath_table():
    doc = json.loads((GOLD / "boys.json").read_text())
    for T, row in zip(doc["T"], doc["F"]):
        F = oracle_lib.boys(12, T)
        np.testing.assert_allclose(F, row, rtol=2e-14, atol=1e-300)


def test_h2_sto3g_szabo_ostlund():
    """S12=0.6593, T11=0.7600, T12=0.2365, V11=-1.8804, (11|11)=0.7746, (11|22)=0.5697,
    (21|11)=0.4441, (21|21)=0.2970, eps=(-0.5782,+0.6703), E_tot=-1.1167 (Szabo & Ostlund, ch. 3)."""
    fb = load_system("hydrogen", "STO-3G").flat()
    S, T, V = oracle_lib.one_electron(fb)
    eri = oracle_lib.eri_tensor(fb)
    assert S[0, 1] == pytest.approx(0.6593, abs=1e-4)
    assert T[0, 0] == pytest.approx(0.7600, abs=1e-4)
    assert T[0, 1] == pytest.approx(0.2365, abs=1e-4)
    assert V[0, 0] == pytest.approx(-1.8804, abs=1e-4)
    assert eri[0, 0, 0, 0] == pytest.approx(0.7746, abs=1e-4)
    assert eri[0, 0, 1, 1] == pytest.approx(0.5697, abs=1e-4)
    assert eri[1, 0, 0, 0] == pytest.approx(0.4441, abs=1e-4)
    assert eri[1, 0, 1, 0] == pytest.approx(0.2970, abs=1e-4)
    # closed-shell energy of the symmetry-determined orbital (1s_a + 1s_b)/sqrt(2(1+S))
    h = T + V
    c = np.ones(2) / np.sqrt(2 * (1 + S[0, 1]))
    P = 2 * np.outer(c, c)
    G = oracle_lib.DenseFock(fb).rhf(P)
    e_el = 0.5 * np.trace(P @ (2 * h + G))
    e_tot = e_el + oracle_lib.nuclear_repulsion(fb)
    assert e_tot == pytest.approx(-1.1167143, abs=2e-6)
    F = h + G
    eps = np.linalg.eigvalsh(np.linalg.solve(np.linalg.cholesky(S), np.linalg.solve(np.linalg.cholesky(S), F).T))
    assert eps[0] == pytest.approx(-0.5782, abs=2e-4)
    assert eps[1] == pytest.approx(0.6703, abs=2e-4)


@pytest.mark.parametrize("mol", ["hydrogen", "water"])
def test_sto3g_against_closed_form_golden(mol):
    doc = json.loads((GOLD / f"{mol}_sto3g.json").read_text())
    fb = load_system(mol, "STO-3G").flat()
    assert fb.n_basis == doc["n_basis"]
    S, T, V = oracle_lib.one_electron(fb)
    np.testing.assert_allclose(S, doc["overlap"], atol=1e-13)
    np.testing.assert_allclose(T, doc["kinetic"], atol=1e-12)
    np.testing.assert_allclose(V, doc["nuclear"], atol=1e-11)
    eri = oracle_lib.eri_tensor(fb)
    np.testing.assert_allclose(unique_eri(eri), doc["eri_unique"], atol=1e-12)


def test_spd_quartets_against_closed_form_golden():
    system, doc = spd_random_system()
    fb = system.flat()
    S, T, V = oracle_lib.one_electron(fb)
    np.testing.assert_allclose(S, doc["overlap"], atol=1e-13)
    np.testing.assert_allclose(T, doc["kinetic"], atol=1e-12)
    np.testing.assert_allclose(V, doc["nuclear"], atol=1e-11)
    assert np.allclose(np.diag(S), 1.0, atol=1e-13)      # every Cartesian component normalised
    for blk in doc["quartets"]:
        a, b, c, d = blk["shells"]
        got = oracle_lib.eri_shell_quartet(fb, a, b, c, d)
        np.testing.assert_allclose(got, np.array(blk["values"]), atol=2e-13, rtol=1e-11)


def test_eri_eightfold_symmetry():
    fb = load_system("water", "STO-3G").flat()
    eri = oracle_lib.eri_tensor(fb)
    for perm in [(1, 0, 2, 3), (0, 1, 3, 2), (2, 3, 0, 1), (3, 2, 1, 0)]:
        np.testing.assert_allclose(eri, eri.transpose(perm), atol=1e-14)


def test_dense_equals_direct_at_tau_zero_rhf_and_uhf():
    """mode (i) reference-faithful N^4 contraction == mode (ii) direct digestion."""
    system, _ = spd_random_system()
    fb = system.flat()
    n = fb.n_basis
    dense = oracle_lib.DenseFock(fb)
    direct = oracle_lib.DirectFock(fb, tau=0.0)
    P = random_symmetric_density(n, 1)
    np.testing.assert_allclose(direct.rhf(P), dense.rhf(P), atol=1e-11)
    Pa, Pb = random_symmetric_density(n, 2), random_symmetric_density(n, 3)
    Ga, Gb = direct.uhf(Pa, Pb)
    np.testing.assert_allclose(Ga, dense.uhf_one(Pa, Pb), atol=1e-11)
    np.testing.assert_allclose(Gb, dense.uhf_one(Pb, Pa), atol=1e-11)


def test_schwarz_bounds_every_integral():
    system, _ = spd_random_system()
    fb = system.flat()
    Q = oracle_lib.schwarz(fb)
    ns = len(fb.shell_l)
    rng = np.random.default_rng(0)
    for _ in range(60):
        a, b, c, d = rng.integers(0, ns, size=4)
        blk = oracle_lib.eri_shell_quartet(fb, a, b, c, d)
        assert np.max(np.abs(blk)) <= Q[a, b] * Q[c, d] * (1 + 1e-12) + 1e-300


def test_water_sto3g_scf_and_invariance():
    """Full SCF with the reference's loop; tr(PS) = n_electrons; E_tot invariant under rotation +
    translation of the molecule."""
    from qchem_rs_b200.basis import MolecularSystem, BasisSet, Atom
    from helpers import DATA
    system = load_system("water", "STO-3G")
    fb = system.flat()
    ints = oracle_lib.one_electron(fb)
    out = hf.restricted_hartree_fock(system, hf.HartreeFockConfig(100, 1e-8), ints, oracle_lib.DenseFock(fb))
    assert out is not None
    assert np.trace(out.density @ ints[0]) == pytest.approx(10.0, abs=1e-10)
    th = 0.7
    rot = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1.0]]) @ \
          np.array([[1, 0, 0], [0, np.cos(0.4), -np.sin(0.4)], [0, np.sin(0.4), np.cos(0.4)]])
    bs = BasisSet.load(DATA / "basis" / "STO-3G.json")
    moved = MolecularSystem.from_atoms([Atom(a.ordinal, rot @ a.position + np.array([0.3, -1.1, 2.0])) for a in system.atoms], bs)
    fb2 = moved.flat()
    out2 = hf.restricted_hartree_fock(moved, hf.HartreeFockConfig(100, 1e-8), oracle_lib.one_electron(fb2), oracle_lib.DenseFock(fb2))
    assert out2.total_energy() == pytest.approx(out.total_energy(), abs=1e-9)
