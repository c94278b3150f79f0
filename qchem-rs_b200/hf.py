"""RHF / UHF drivers: the host-side SCF loops of core/src/hf/rhf.rs and core/src/hf/uhf.rs restated
with every parity-relevant quirk kept.  The two-electron matrix G(P) comes from a
pluggable Fock builder -- the CUDA engine (`engine.FockEngine`) in the product, the CPU oracle in
the tests -- which replaces rhf.rs:45,58-62,67-68 and uhf.rs:55,90-91.  Everything else (guess,
DIIS, eigensolve, energy expression, convergence test) is host code and not the optimisation target.

Quirks kept on purpose:
* Hueckel guess scales the diagonal by 1.75 as well              rhf.rs:139-143, uhf.rs:197-201
* DIIS(4,6) for RHF, DIIS(2,8) per spin for UHF                  rhf.rs:65, uhf.rs:74-77
* loop runs 0..=max_iterations                                   rhf.rs:66, uhf.rs:79
* energy uses the NEW density with G of the OLD density          rhf.rs:84-85
* convergence looks at diag(delta P) only                        rhf.rs:87-88, uhf.rs:126-127
* UHF: n_alpha = n_beta = n_electrons/2, rms halved twice        uhf.rs:43-45, :137, :139
* `iterations` is the loop index at convergence                  rhf.rs:101
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from .diis import Diis


@dataclass
class HartreeFockConfig:          # core/src/hf/mod.rs:9-15
    max_iterations: int = 100
    epsilon: float = 1e-6


@dataclass
class RestrictedHartreeFockOutput:   # rhf.rs:14-30
    orbital_energies: np.ndarray
    electronic_energy: float
    nuclear_repulsion: float
    iterations: int
    density: np.ndarray = None
    fock: np.ndarray = None
    fock_build_seconds: List[float] = None

    def total_energy(self) -> float:
        return self.electronic_energy + self.nuclear_repulsion


@dataclass
class UnrestrictedHartreeFockOutput:  # uhf.rs:15-34
    orbital_energies_alpha: np.ndarray
    orbital_energies_beta: np.ndarray
    electronic_energy: float
    nuclear_repulsion: float
    iterations: int
    density_alpha: np.ndarray = None
    density_beta: np.ndarray = None
    fock_alpha: np.ndarray = None
    fock_beta: np.ndarray = None

    def total_energy(self) -> float:
        return self.electronic_energy + self.nuclear_repulsion


def compute_nuclear_repulsion(atoms) -> float:      # rhf.rs:110-122
    e = 0.0
    for a in range(len(atoms)):
        for b in range(a + 1, len(atoms)):
            e += float(atoms[a].ordinal * atoms[b].ordinal) / float(
                np.linalg.norm(atoms[b].position - atoms[a].position))
    return e


def sorted_eigs(m: np.ndarray):                     # utils.rs:20-36 (ascending eigenvalues)
    w, v = np.linalg.eigh(m)
    order = np.argsort(w, kind="stable")
    return v[:, order], w[order]


def compute_transformation_matrix(overlap):         # rhf.rs:124-131  X = U s^-1/2 U^T
    w, u = np.linalg.eigh(overlap)
    d = u.T @ (overlap @ u)
    return u @ (np.diag(1.0 / np.sqrt(np.diag(d))) @ u.T)


def compute_updated_density(c, n_occ, factor):      # rhf.rs:169-181 (factor 2), uhf.rs:229-241 (1)
    co = c[:, :n_occ]
    return factor * (co @ co.T)


def compute_hueckel_density(h, s, x, n_occ, factor):  # rhf.rs:133-150, uhf.rs:191-208
    d = np.diag(h)
    h_eht = 1.75 * s * (d[:, None] + d[None, :]) / 2.0
    cp, _ = sorted_eigs(x.T @ (h_eht @ x))
    return compute_updated_density(x @ cp, n_occ, factor)


def restricted_hartree_fock(system, config: HartreeFockConfig, integrals, fock_builder,
                            keep_history: bool = False) -> Optional[RestrictedHartreeFockOutput]:
    """rhf.rs:32-108.  `integrals` = (S, T, V); `fock_builder.rhf(P)` returns G = J[P] - K[P]/2."""
    import time
    n_electrons = system.n_electrons()
    nuclear_repulsion = compute_nuclear_repulsion(system.atoms)
    S, T, V = integrals
    h = T + V
    x = compute_transformation_matrix(S)
    density = compute_hueckel_density(h, S, x, n_electrons // 2, 2.0)
    diis = Diis(4, 6)
    times = []
    history = []
    for iteration in range(config.max_iterations + 1):
        t0 = time.perf_counter()
        g = fock_builder.rhf(density)
        times.append(time.perf_counter() - t0)
        fock = h + g
        error = fock @ density @ S - S @ density @ fock
        fock_x = diis.fock(error, fock)
        if fock_x is None:
            raise RuntimeError("DIIS failed")                      # rhf.rs:73 expect()
        cp, orbital_energies = sorted_eigs(x.T @ (fock_x @ x))
        c = x @ cp
        new_density = compute_updated_density(c, n_electrons // 2, 2.0)
        change = new_density - density
        density = density + change * 1.0
        electronic_energy = 0.5 * np.trace(density @ (2.0 * h + g))
        rms = float(np.sqrt(np.sum(np.diag(change) ** 2) / S.shape[0]))
        if keep_history:
            history.append((iteration, electronic_energy, rms, fock.copy()))
        if rms < config.epsilon:
            out = RestrictedHartreeFockOutput(orbital_energies, float(electronic_energy),
                                              nuclear_repulsion, iteration, density, fock, times)
            out.history = history
            return out
    return None


def unrestricted_hartree_fock(system, config: HartreeFockConfig, integrals, fock_builder,
                              n_alpha: Optional[int] = None, n_beta: Optional[int] = None,
                              ) -> Optional[UnrestrictedHartreeFockOutput]:
    """uhf.rs:36-189.  `fock_builder.uhf(Pa, Pb)` returns (Ga, Gb), G_s = J[Pa+Pb] - K[P_s].

    With n_alpha/n_beta left at None the reference semantics apply (n_alpha = n_beta =
    n_electrons/2, uhf.rs:43-45).  Passing them explicitly is a labelled extension (true open shell,
    e.g. triplet O2 with (9, 7)); the reference ignores --charge/--spin-multiplicity (main.rs:115-116).
    """
    n_electrons = system.n_electrons()
    if n_alpha is None:
        n_alpha = n_electrons // 2
    if n_beta is None:
        n_beta = n_electrons // 2
    nuclear_repulsion = compute_nuclear_repulsion(system.atoms)
    S, T, V = integrals
    n = S.shape[0]
    h = T + V
    x = compute_transformation_matrix(S)
    dens = [compute_hueckel_density(h, S, x, n_alpha, 1.0), compute_hueckel_density(h, S, x, n_beta, 1.0)]
    occ = [n_alpha, n_beta]
    diis = [Diis(2, 8), Diis(2, 8)]
    g_keep = [np.zeros((n, n)), np.zeros((n, n))]
    f_keep = [None, None]
    coeffs = [None, None]
    energies = [None, None]
    for iteration in range(config.max_iterations + 1):
        # both spins see the previous iteration's densities (uhf.rs:81-108), so one engine call
        # G_a, G_b = build_uhf(P_a, P_b) serves the two compute_electronic_hamiltonian calls
        ga, gb = fock_builder.uhf(dens[0], dens[1])
        for spin, g in enumerate((ga, gb)):
            fock = h + g
            error = fock @ dens[spin] @ S - S @ dens[spin] @ fock
            fock_x = diis[spin].fock(error, fock)
            if fock_x is None:
                raise RuntimeError(f"DIIS failed in spin {spin}")
            g_keep[spin] = g
            f_keep[spin] = fock
            cp, e = sorted_eigs(x.T @ (fock_x @ x))
            coeffs[spin] = x @ cp
            energies[spin] = e
        density_rms = 0.0
        for spin in range(2):
            new_density = compute_updated_density(coeffs[spin], occ[spin], 1.0)
            change = new_density - dens[spin]
            dens[spin] = dens[spin] + change * 1.0
            density_rms += float(np.sqrt(np.sum(np.diag(change) ** 2) / n))
        density_rms /= 2.0
        if density_rms / 2.0 < config.epsilon:
            e_a = 0.5 * np.trace(dens[0] @ (2.0 * h + g_keep[0]))
            e_b = 0.5 * np.trace(dens[1] @ (2.0 * h + g_keep[1]))
            return UnrestrictedHartreeFockOutput(energies[0], energies[1], float(e_a + e_b),
                                                 nuclear_repulsion, iteration, dens[0], dens[1],
                                                 f_keep[0], f_keep[1])
    return None


class IncrementalFock:
    """Fock builder for the host SCF loops above that uses the engine's difference-density builds:
    G_k = G_{k-1} + G(P_k - P_{k-1}), a full rebuild every `full_every` iterations."""

    def __init__(self, engine, full_every: int = 8):
        self.eng = engine
        self.full_every = max(1, int(full_every))
        self.n = 0
        self.log = []      # (iteration, quartets, kernel_ms) per build

    def _reset(self):
        r = (self.n % self.full_every) == 0
        self.n += 1
        return r

    def rhf(self, P):
        g = self.eng.rhf_incremental(P, reset=self._reset())
        st = self.eng.stats()
        self.log.append((self.n - 1, st["quartets"], st["kernel_ms"]))
        return g

    def uhf(self, Pa, Pb):
        g = self.eng.uhf_incremental(Pa, Pb, reset=self._reset())
        st = self.eng.stats()
        self.log.append((self.n - 1, st["quartets"], st["kernel_ms"]))
        return g


def restricted_hartree_fock_device(system, config: HartreeFockConfig, integrals, engine,
                                   full_rebuild_every: int = 0) -> Optional[RestrictedHartreeFockOutput]:
    """rhf.rs:32-108 with the whole iteration on the GPU (qcf_scf_init / qcf_scf_step):
    P, G, F, the DIIS history and the orbitals stay in HBM; the host sees scalars only."""
    import time
    S, T, V = integrals
    t0 = time.perf_counter()
    engine.scf_init(S, T + V, system.n_electrons() // 2, unrestricted=False, full_rebuild_every=full_rebuild_every)
    init_s = time.perf_counter() - t0
    nuclear_repulsion = compute_nuclear_repulsion(system.atoms)
    steps = []
    for _ in range(config.max_iterations + 1):
        info = engine.scf_step(config.epsilon)
        steps.append(info)
        if info["converged"]:
            out = RestrictedHartreeFockOutput(engine.scf_get("orbital_energies"), info["electronic_energy"], nuclear_repulsion,
                                              info["iteration"], engine.scf_get("density"), engine.scf_get("fock"),
                                              [s["build_ms"] * 1e-3 for s in steps])
            out.steps = steps
            out.init_s = init_s
            return out
    return None


def unrestricted_hartree_fock_device(system, config: HartreeFockConfig, integrals, engine, n_alpha: Optional[int] = None,
                                     n_beta: Optional[int] = None, full_rebuild_every: int = 0
                                     ) -> Optional[UnrestrictedHartreeFockOutput]:
    """uhf.rs:36-189 on the GPU.  n_alpha / n_beta default to the reference semantics (n_electrons / 2 each)."""
    S, T, V = integrals
    n_el = system.n_electrons()
    n_alpha = n_el // 2 if n_alpha is None else n_alpha
    n_beta = n_el // 2 if n_beta is None else n_beta
    engine.scf_init(S, T + V, n_alpha, n_beta, unrestricted=True, full_rebuild_every=full_rebuild_every)
    nuclear_repulsion = compute_nuclear_repulsion(system.atoms)
    for _ in range(config.max_iterations + 1):
        info = engine.scf_step(config.epsilon)
        if info["converged"]:
            return UnrestrictedHartreeFockOutput(engine.scf_get("orbital_energies", 0), engine.scf_get("orbital_energies", 1),
                                                 info["electronic_energy"], nuclear_repulsion, info["iteration"],
                                                 engine.scf_get("density", 0), engine.scf_get("density", 1),
                                                 engine.scf_get("fock", 0), engine.scf_get("fock", 1))
    return None
