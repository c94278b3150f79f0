#!/usr/bin/env python
"""The four molecular BASELINE.json configs end to end on the CUDA engine: full SCF with the reference's loop
(qchem-rs_b200/hf.py), one-electron matrices from qcf_one_electron, per-build device time from qcf_stats.
Prints one JSON line per config (profiles/r1_baseline_configs.jsonl)."""
import json, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
import qcpkg
pkg = qcpkg.load()

def system(mol, basis):
    bs = pkg.BasisSet.load(ROOT / "data" / "basis" / f"{basis}.json")
    return pkg.MolecularSystem.load(ROOT / "data" / "mol" / f"{mol}.json", bs)

CONFIGS = [("H2O STO-3G RHF", "water", "STO-3G", "rhf", {}),
           ("benzene 6-31G RHF (D6h geometry)", "benzene_d6h", "6-31G", "rhf", {}),
           ("O2 triplet 6-31G UHF (9 alpha, 7 beta; labelled extension)", "oxygen", "6-31G", "uhf", {"n_alpha": 9, "n_beta": 7}),
           ("O2 6-31G UHF, reference semantics n_alpha = n_beta = 8", "oxygen", "6-31G", "uhf", {}),
           ("caffeine 6-31G* RHF", "caffeine", "6-31G_st", "rhf", {})]
for name, mol, basis, kind, kw in CONFIGS:
    sysm = system(mol, basis)
    with pkg.engine.FockEngine(sysm, tau=1e-12) as eng:
        ints = eng.one_electron()
        ms, q = [], []
        class Tap:
            def rhf(self, P):
                g = eng.rhf(P); st = eng.stats(); ms.append(st["kernel_ms"]); q.append(st["quartets"]); return g
            def uhf(self, Pa, Pb):
                g = eng.uhf(Pa, Pb); st = eng.stats(); ms.append(st["kernel_ms"]); q.append(st["quartets"]); return g
        cfg = pkg.HartreeFockConfig(60 if "triplet" in name else 100, 1e-6)
        t0 = time.perf_counter()
        out = (pkg.restricted_hartree_fock(sysm, cfg, ints, Tap()) if kind == "rhf"
               else pkg.unrestricted_hartree_fock(sysm, cfg, ints, Tap(), **kw))
        wall = time.perf_counter() - t0
        st = eng.stats()
    line = {"config": name, "n_basis": st["n_basis"], "converged": out is not None,
            "iterations": None if out is None else out.iterations,
            "total_energy_Eh": None if out is None else out.total_energy(),
            "fock_builds": len(ms), "fock_build_ms_median": float(np.median(ms[1:] or ms)),
            "quartets_per_build": int(np.median(q)), "quartets_unscreened": st["quartets_total"], "scf_wall_s": wall}
    print(json.dumps(line), flush=True)
