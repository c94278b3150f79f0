"""GPU parity tests (run with -m gpu on a B200).  Every check goes through the C ABI of
libqcfock.so and compares with the CPU oracle / the committed golden vectors.

Tolerances (north_star): |dE_total| < 1e-8 Eh, max |dF_ij| < 1e-9, same SCF iteration count."""
import json

import numpy as np
import pytest

from helpers import GOLD, load_system, water_cluster, spd_random_system, random_symmetric_density, oracle_lib
from qchem_rs_b200 import hf, engine

pytestmark = pytest.mark.gpu

F_TOL = 1e-9
E_TOL = 1e-8


@pytest.fixture(scope="module")
def spd():
    system, doc = spd_random_system()
    return system, doc, engine.FockEngine(system, tau=engine.QCF_TAU_NONE)


def test_boys_matches_mpmath_table(spd):
    _, _, eng = spd
    doc = json.loads((GOLD / "boys.json").read_text())
    T = np.array(doc["T"])
    ref = np.array(doc["F"])
    for mmax in range(0, 9):
        F = eng.boys(mmax, T)
        np.testing.assert_allclose(F, ref[:, :mmax + 1], rtol=5e-14, atol=1e-300)


def test_boys_dense_sweep_against_oracle(spd):
    _, _, eng = spd
    T = np.concatenate([np.linspace(0, 70, 3501), np.linspace(63.9, 64.1, 201), np.geomspace(40, 1e4, 200)])
    F = eng.boys(8, T)
    ref = np.array([oracle_lib.boys(8, t) for t in T])
    np.testing.assert_allclose(F, ref, rtol=1e-13, atol=1e-300)


def test_quartets_match_closed_form_golden(spd):
    _, doc, eng = spd
    for blk in doc["quartets"]:
        a, b, c, d = blk["shells"]
        got = eng.eri_quartet(a, b, c, d)
        np.testing.assert_allclose(got, np.array(blk["values"]), atol=5e-13, rtol=1e-11)


def test_all_class_quartets_match_oracle(spd):
    """Every ordered combination of angular momenta (s,p,d)^4 = 81 quartets, incl. permuted orders."""
    system, _, eng = spd
    fb = system.flat()
    first = {0: 0, 1: 4, 2: 8}      # first shell index of each l in the spd_random system
    rng = np.random.default_rng(5)
    for la in range(3):
        for lb in range(3):
            for lc in range(3):
                for ld in range(3):
                    q = [first[l] + int(rng.integers(0, 4)) for l in (la, lb, lc, ld)]
                    got = eng.eri_quartet(*q)
                    ref = oracle_lib.eri_shell_quartet(fb, *q)
                    np.testing.assert_allclose(got, ref, atol=5e-13, rtol=1e-11, err_msg=str(q))


def test_contracted_quartets_water_sto3g():
    system = load_system("water", "STO-3G")
    fb = system.flat()
    ns = len(system.shells)
    with engine.FockEngine(system, tau=engine.QCF_TAU_NONE) as eng:
        for a in range(ns):
            for b in range(ns):
                for c in range(ns):
                    for d in range(ns):
                        np.testing.assert_allclose(eng.eri_quartet(a, b, c, d), oracle_lib.eri_shell_quartet(fb, a, b, c, d),
                                                   atol=5e-13, rtol=1e-11)


def test_schwarz_matches_oracle(spd):
    """qcf_schwarz reports the bounds in the engine's component-scaled space (every Cartesian
    component carries the x^l normalisation; the density is scaled to match), so the oracle blocks
    are divided by the component factors before the comparison."""
    from qchem_rs_b200.basis import component_scale
    system, _, eng = spd
    fb = system.flat()
    ns = len(fb.shell_l)
    ref = np.zeros((ns, ns))
    for a in range(ns):
        for b in range(ns):
            blk = oracle_lib.eri_shell_quartet(fb, a, b, a, b)
            sa, sb = np.array(component_scale(int(fb.shell_l[a]))), np.array(component_scale(int(fb.shell_l[b])))
            diag = np.einsum("ijij->ij", blk) / np.outer(sa, sb) ** 2
            ref[a, b] = np.sqrt(np.max(np.abs(diag)))
    np.testing.assert_allclose(eng.schwarz(), ref, rtol=1e-11, atol=1e-14)


@pytest.mark.parametrize("name", ["water_sto3g", "spd_random", "benzene_631g", "water3_631gs"])
@pytest.mark.parametrize("tau", [engine.QCF_TAU_NONE, 1e-12])
def test_fock_rhf_uhf_jk_against_dense_oracle(name, tau):
    """G(P) against the reference-faithful dense contraction (rhf.rs:152-167, uhf.rs:210-227),
    random symmetric densities (no symmetry of a converged density to hide errors)."""
    if name == "water_sto3g":
        system = load_system("water", "STO-3G")
    elif name == "spd_random":
        system = spd_random_system()[0]
    elif name == "benzene_631g":
        system = load_system("benzene", "6-31G")
    else:
        system = water_cluster(3)
    fb = system.flat()
    n = fb.n_basis
    dense = oracle_lib.DenseFock(fb)
    with engine.FockEngine(system, tau=tau) as eng:
        P = random_symmetric_density(n, 11)
        np.testing.assert_allclose(eng.rhf(P), dense.rhf(P), atol=F_TOL)
        Pa, Pb = random_symmetric_density(n, 12), random_symmetric_density(n, 13)
        Ga, Gb = eng.uhf(Pa, Pb)
        np.testing.assert_allclose(Ga, dense.uhf_one(Pa, Pb), atol=F_TOL)
        np.testing.assert_allclose(Gb, dense.uhf_one(Pb, Pa), atol=F_TOL)
        (J,), (K,) = eng.jk([P])
        np.testing.assert_allclose(J, np.einsum("ijkl,kl->ij", dense.eri, P), atol=F_TOL)
        np.testing.assert_allclose(K, np.einsum("ikjl,kl->ij", dense.eri, P), atol=F_TOL)
        st = eng.stats()
        assert st["quartets"] > 0 and st["quartets"] <= st["quartets_total"]
        if tau == engine.QCF_TAU_NONE:
            assert st["quartets"] == st["quartets_total"]


def test_empty_and_degenerate_inputs():
    """Zero density gives a zero matrix; a single s shell (one quartet, all degeneracy factors) works."""
    from qchem_rs_b200.basis import MolecularSystem, Atom, Shell
    system = MolecularSystem([Atom(1, np.zeros(3))])
    system.shells.append(Shell(0, np.array([0.8]), np.array([1.0])))
    system.shell_atom.append(0)
    fb = system.flat()
    with engine.FockEngine(system, tau=engine.QCF_TAU_NONE) as eng:
        assert np.all(eng.rhf(np.zeros((1, 1))) == 0.0)
        P = np.array([[2.0]])
        np.testing.assert_allclose(eng.rhf(P), oracle_lib.DenseFock(fb).rhf(P), atol=1e-13)


@pytest.mark.parametrize("world", [2, 3, 8])
def test_bra_partition_sums_to_full_build(world):
    """'Distributed without a cluster' (SURVEY.md 4): the rank partition of the bra-pair list run
    sequentially on one GPU sums to the single-rank result."""
    system = water_cluster(2)
    n = system.n_basis()
    P = random_symmetric_density(n, 3)
    with engine.FockEngine(system, tau=1e-12) as eng:
        full = eng.rhf(P)
        q_full = eng.stats()["quartets"]
    acc = np.zeros_like(full)
    q = 0
    for r in range(world):
        with engine.FockEngine(system, tau=1e-12, rank=r, world_size=world) as eng:
            acc += eng.rhf(P)
            q += eng.stats()["quartets"]
    np.testing.assert_allclose(acc, full, atol=1e-11)
    assert q == q_full


def _scf_pair(system, kind, tau=1e-12, eps=1e-8, dense=True, **kw):
    fb = system.flat()
    ints = oracle_lib.one_electron(fb)
    cfg = hf.HartreeFockConfig(100, eps)
    ref_builder = oracle_lib.DenseFock(fb) if dense else oracle_lib.DirectFock(fb, tau=0.0)
    with engine.FockEngine(system, tau=tau) as eng:
        if kind == "rhf":
            ref = hf.restricted_hartree_fock(system, cfg, ints, ref_builder, keep_history=True)
            got = hf.restricted_hartree_fock(system, cfg, ints, eng, keep_history=True)
        else:
            ref = hf.unrestricted_hartree_fock(system, cfg, ints, ref_builder, **kw)
            got = hf.unrestricted_hartree_fock(system, cfg, ints, eng, **kw)
    return ref, got


@pytest.mark.parametrize("mol,basis", [("water", "STO-3G"), ("benzene_d6h", "6-31G"), ("ethylene", "6-31G")])
def test_rhf_scf_parity(mol, basis):
    ref, got = _scf_pair(load_system(mol, basis), "rhf")
    assert ref is not None and got is not None
    assert got.iterations == ref.iterations
    assert abs(got.total_energy() - ref.total_energy()) < E_TOL
    assert np.max(np.abs(got.fock - ref.fock)) < F_TOL
    for (_, _, _, fr), (_, _, _, fg) in zip(ref.history, got.history):
        assert np.max(np.abs(fr - fg)) < F_TOL


def test_rhf_scf_h2_szabo_ostlund():
    """H2/STO-3G.  The Fock matrices agree to 1 ulp on identical densities and the converged energy
    is the Szabo-Ostlund value, but the iteration COUNT is not asserted: for N = 2 all DIIS error
    matrices are proportional, the reference's B matrix (diis.rs:40-51) is singular from the 4th
    sample on and its QR solve amplifies 1e-16 to 1e-3 (SURVEY.md 3.3, 7 'hard parts')."""
    ref, got = _scf_pair(load_system("hydrogen", "STO-3G"), "rhf")
    assert ref is not None and got is not None
    assert got.total_energy() == pytest.approx(-1.1167143, abs=2e-6)
    assert abs(got.total_energy() - ref.total_energy()) < E_TOL
    for (_, _, _, fr), (_, _, _, fg) in list(zip(ref.history, got.history))[:3]:
        assert np.max(np.abs(fr - fg)) < 1e-14


def test_rhf_benzene_reference_geometry_trajectory():
    """benzene/6-31G with the reference's own data/mol/benzene.json (C ring radius 1.5165 bohr -- a
    compressed, unphysical geometry; the Hueckel guess density has elements ~1e3 and the first SCF
    steps are chaotic, so 1e-16 relative differences are amplified along the trajectory: the two
    converged energies differ by ~1e-8 although every single Fock build agrees to ~2e-12).  Parity is
    therefore asserted one step at a time on the ORACLE's trajectory, G_gpu(P_k) vs G_oracle(P_k) with
    the tolerance scaled by max(1, max|P_k|); the free-running GPU SCF must take the same number of
    iterations (+-1) and land within 1e-6 Eh."""
    system = load_system("benzene", "6-31G")
    fb = system.flat()
    ints = oracle_lib.one_electron(fb)
    dense = oracle_lib.DenseFock(fb)
    worst = [0.0]
    with engine.FockEngine(system, tau=1e-12) as eng:
        class Both:
            def rhf(self, P):
                g_ref = dense.rhf(P)
                worst[0] = max(worst[0], float(np.max(np.abs(eng.rhf(P) - g_ref))) / max(1.0, float(np.max(np.abs(P)))))
                return g_ref
        ref = hf.restricted_hartree_fock(system, hf.HartreeFockConfig(100, 1e-8), ints, Both())
        got = hf.restricted_hartree_fock(system, hf.HartreeFockConfig(100, 1e-8), ints, eng)
    assert ref is not None and got is not None
    assert 0.0 < worst[0] < F_TOL
    # free-running comparison: same count in every run observed so far, but the trajectory is chaotic and the
    # FP64 atomics are unordered, so one iteration of slack and 1e-6 Eh are allowed (the per-step bound above
    # is the parity claim)
    assert abs(got.iterations - ref.iterations) <= 1
    assert abs(got.total_energy() - ref.total_energy()) < 1e-6


def test_uhf_scf_parity_o2_reference_semantics():
    """O2/6-31G with the reference's UHF semantics n_alpha = n_beta = 8 (uhf.rs:43-45).  Eight electrons per spin
    half-fill the degenerate pi* pair, so which combination gets occupied is decided by round-off in the
    eigensolver: two builders that agree to 1e-12 per Fock build take different numbers of iterations (observed
    9 vs 14) and can even land on different UHF solutions (-149.46138 or the symmetry-broken -149.51350 Eh; the
    oracle does the same when its summation order is changed, and FP64 atomics are unordered).  A free-running
    comparison is therefore meaningless for this input; asserted instead: parity of every single Fock build,
    < 1e-9, along BOTH trajectories (oracle-driven and GPU-driven), and that both runs converge."""
    system = load_system("oxygen", "6-31G")
    fb = system.flat()
    ints = oracle_lib.one_electron(fb)
    dense = oracle_lib.DenseFock(fb)
    worst = [0.0, 0.0]
    with engine.FockEngine(system, tau=1e-12) as eng:
        class Both:
            def __init__(self, drive_gpu):
                self.drive_gpu = drive_gpu

            def uhf(self, Pa, Pb):
                ra, rb = dense.uhf(Pa, Pb)
                ga, gb = eng.uhf(Pa, Pb)
                k = 1 if self.drive_gpu else 0
                worst[k] = max(worst[k], float(np.max(np.abs(ga - ra))), float(np.max(np.abs(gb - rb))))
                return (ga, gb) if self.drive_gpu else (ra, rb)
        ref = hf.unrestricted_hartree_fock(system, hf.HartreeFockConfig(100, 1e-8), ints, Both(False))
        got = hf.unrestricted_hartree_fock(system, hf.HartreeFockConfig(100, 1e-8), ints, Both(True))
    assert ref is not None and got is not None
    assert 0.0 < worst[0] < F_TOL and 0.0 < worst[1] < F_TOL
    for out in (ref, got):
        assert -149.6 < out.total_energy() < -149.4


def test_uhf_o2_triplet_extension_trajectory():
    """Labelled extension: true triplet O2 (9 alpha, 7 beta) exercises Ka != Kb.  The reference's loop
    (DIIS(2,8) without a conditioning guard) stalls at ~1e-4 for this state on the oracle as well, so
    parity is asserted step by step along the oracle's first 15 iterations."""
    system = load_system("oxygen", "6-31G")
    fb = system.flat()
    ints = oracle_lib.one_electron(fb)
    dense = oracle_lib.DenseFock(fb)
    worst = [0.0]

    with engine.FockEngine(system, tau=1e-12) as eng:
        class Both:
            def uhf(self, Pa, Pb):
                ra, rb = dense.uhf(Pa, Pb)
                ga, gb = eng.uhf(Pa, Pb)
                worst[0] = max(worst[0], float(np.max(np.abs(ga - ra))), float(np.max(np.abs(gb - rb))))
                return ra, rb
        hf.unrestricted_hartree_fock(system, hf.HartreeFockConfig(14, 1e-12), ints, Both(), n_alpha=9, n_beta=7)
    assert 0.0 < worst[0] < F_TOL


def test_rhf_scf_parity_water_dimer_631gs():
    """d-shell classes through a whole SCF: (H2O)_2 / 6-31G*, N = 38."""
    ref, got = _scf_pair(water_cluster(2), "rhf")
    assert ref is not None and got is not None
    assert got.iterations == ref.iterations
    assert abs(got.total_energy() - ref.total_energy()) < E_TOL
    assert np.max(np.abs(got.fock - ref.fock)) < F_TOL


@pytest.mark.parametrize("name", ["water_sto3g", "spd_random", "benzene_631g", "water3_631gs", "caffeine_631gs"])
def test_one_electron_matrices_match_oracle(name):
    """qcf_one_electron (stand-in for molint::overlap/kinetic/nuclear, rhf.rs:41-43) vs the oracle."""
    if name == "water_sto3g":
        system = load_system("water", "STO-3G")
    elif name == "spd_random":
        system = spd_random_system()[0]
    elif name == "benzene_631g":
        system = load_system("benzene", "6-31G")
    elif name == "caffeine_631gs":
        system = load_system("caffeine", "6-31G_st")
    else:
        system = water_cluster(3)
    fb = system.flat()
    S0, T0, V0 = oracle_lib.one_electron(fb)
    with engine.FockEngine(system, tau=1e-12) as eng:
        S, T, V = eng.one_electron()
    np.testing.assert_allclose(S, S0, atol=1e-12, rtol=1e-12)
    np.testing.assert_allclose(T, T0, atol=1e-11, rtol=1e-12)
    np.testing.assert_allclose(V, V0, atol=1e-10, rtol=1e-12)


def test_caffeine_631gs_rhf_against_direct_oracle():
    """BASELINE config 4: caffeine / 6-31G* RHF (N = 230, d-shell classes).  The reference-faithful dense
    path would need 22 GB for the N^4 tensor, so the oracle's direct-SCF form (tau = 0, no screening) is the
    checker: parity of G for the first SCF density (one CPU build, ~10 s), then the GPU SCF must converge."""
    system = load_system("caffeine", "6-31G_st")
    fb = system.flat()
    assert fb.n_basis == 230
    direct = oracle_lib.DirectFock(fb, tau=0.0)
    worst = []
    with engine.FockEngine(system, tau=1e-12) as eng:
        ints = eng.one_electron()

        class Both:
            def __init__(self):
                self.n = 0

            def rhf(self, P):
                g = eng.rhf(P)
                if self.n < 1:
                    worst.append(float(np.max(np.abs(g - direct.rhf(P)))) / max(1.0, float(np.max(np.abs(P)))))
                self.n += 1
                return g
        out = hf.restricted_hartree_fock(system, hf.HartreeFockConfig(100, 1e-7), ints, Both())
        st = eng.stats()
    assert len(worst) == 1 and max(worst) < F_TOL
    assert out is not None and out.iterations < 60
    assert -680.0 < out.total_energy() < -670.0      # RHF/6-31G* caffeine is about -676 Eh
    assert st["prim_pairs_kept"] <= st["prim_pairs"]


def test_water_cluster_5_scf_against_direct_oracle():
    """(H2O)_5 / 6-31G* (N = 95): full SCF, GPU vs the oracle's direct path inside the same harness."""
    system = water_cluster(5)
    fb = system.flat()
    ints = oracle_lib.one_electron(fb)
    cfg = hf.HartreeFockConfig(100, 1e-7)
    ref = hf.restricted_hartree_fock(system, cfg, ints, oracle_lib.DirectFock(fb, tau=0.0))
    with engine.FockEngine(system, tau=1e-12) as eng:
        got = hf.restricted_hartree_fock(system, cfg, ints, eng)
    assert ref is not None and got is not None
    assert got.iterations == ref.iterations
    assert abs(got.total_energy() - ref.total_energy()) < E_TOL
    assert np.max(np.abs(got.fock - ref.fock)) < F_TOL


def test_device_resident_paths_match_host_calls():
    """qcf_build_rhf_dev / qcf_build_uhf_dev on torch's current stream (the multi-GPU product path,
    qchem-rs_b200/distributed.py::DeviceFock, world size 1 here) against the host-buffer C-ABI calls."""
    import torch
    from qchem_rs_b200 import distributed
    system = water_cluster(2)
    n = system.n_basis()
    P, Pb = random_symmetric_density(n, 5), random_symmetric_density(n, 6)
    dev = torch.device("cuda", 0)
    with engine.FockEngine(system, tau=1e-12) as eng:
        fock = distributed.DeviceFock(eng, dev)
        g_host = eng.rhf(P)
        np.testing.assert_allclose(fock.rhf(P), g_host, atol=1e-12)
        fock.hP[0].copy_(torch.from_numpy(P))
        np.testing.assert_allclose(fock.rhf_pinned().numpy(), g_host, atol=1e-12)
        ga, gb = eng.uhf(P, Pb)
        da, db = fock.uhf(P, Pb)
        np.testing.assert_allclose(da, ga, atol=1e-12)
        np.testing.assert_allclose(db, gb, atol=1e-12)
        prof = eng.launch_profile()
        assert len(prof) > 0 and sum(r["quartets"] for r in prof) == eng.stats()["quartets"]


def test_gpu_quartets_permutational_symmetry_and_schwarz_bound(spd):
    """Invariants of the device integrals themselves: 8-fold permutational symmetry and
    |(ab|cd)| <= sqrt(max (ab|ab)) sqrt(max (cd|cd)) for random shell quartets of the s/p/d test system."""
    _, _, eng = spd
    rng = np.random.default_rng(17)
    ns = 12
    qmax = {}

    def q(a, b):
        if (a, b) not in qmax:
            blk = eng.eri_quartet(a, b, a, b)
            qmax[(a, b)] = float(np.sqrt(np.max(np.abs(np.einsum("ijij->ij", blk)))))
        return qmax[(a, b)]
    for _ in range(40):
        a, b, c, d = (int(x) for x in rng.integers(0, ns, size=4))
        v = eng.eri_quartet(a, b, c, d)
        np.testing.assert_allclose(eng.eri_quartet(b, a, c, d), v.transpose(1, 0, 2, 3), atol=1e-13)
        np.testing.assert_allclose(eng.eri_quartet(a, b, d, c), v.transpose(0, 1, 3, 2), atol=1e-13)
        np.testing.assert_allclose(eng.eri_quartet(c, d, a, b), v.transpose(2, 3, 0, 1), atol=1e-13)
        assert np.max(np.abs(v)) <= q(a, b) * q(c, d) * (1 + 1e-10) + 1e-14


def test_scf_energy_is_rotation_and_translation_invariant_on_gpu():
    """(H2O)_2 / 6-31G* (Cartesian d shells): E_total from the engine-driven SCF is invariant under a rigid
    rotation + translation of the molecule.  The reference's Hueckel guess is not rotation-covariant (rhf.rs:139-143
    scales the diagonal too), so the two runs follow different trajectories and stop at slightly different points of
    a first-order energy expression: 5e-8 Eh is asserted (observed 1.4e-9)."""
    from qchem_rs_b200.basis import MolecularSystem, Atom, BasisSet
    from helpers import DATA
    bs = BasisSet.load(DATA / "basis" / "6-31G_st.json")
    atoms = water_cluster(2).atoms
    th, ph = 0.7, -1.1
    rz = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1]])
    rx = np.array([[1, 0, 0], [0, np.cos(ph), -np.sin(ph)], [0, np.sin(ph), np.cos(ph)]])
    rot = rz @ rx
    moved = [Atom(a.ordinal, rot @ a.position + np.array([1.5, -2.0, 0.75])) for a in atoms]
    energies = []
    for at in (atoms, moved):
        system = MolecularSystem.from_atoms(at, bs)
        with engine.FockEngine(system, tau=1e-12) as eng:
            out = hf.restricted_hartree_fock(system, hf.HartreeFockConfig(100, 1e-9), eng.one_electron(), eng)
        assert out is not None
        energies.append(out.total_energy())
    assert abs(energies[0] - energies[1]) < 5e-8
