"""Writes the committed fixture bench.py's reference arm loads: a rank-n_occ factor L (P = L L^T) of the density that
enters the Fock build of SCF iteration `iters` of (H2O)_n / 6-31G*, produced by the reference's RHF loop (hf.py) driving
the CUDA engine.  Run on a GPU box:  python tools/make_bench_density.py 53 6 gpurun_out/  then copy the .npz into
tests/golden/.  (float64, compressed; 1007 x 265 doubles ~ 2 MB.)"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import qcpkg  # noqa: E402

pkg = qcpkg.load()


def main():
    n = int(sys.argv[1]); iters = int(sys.argv[2]); out = Path(sys.argv[3])
    bs = pkg.BasisSet.load(ROOT / "data" / "basis" / "6-31G_st.json")
    system = pkg.MolecularSystem.from_atoms(pkg.molecules.water_cluster(n), bs)
    seen = {}
    with pkg.engine.FockEngine(system, tau=1e-12) as eng:
        ints = eng.one_electron()

        class Tap:
            def rhf(self, P):
                seen["P"] = np.array(P, copy=True)
                return eng.rhf(P)
        pkg.hf.restricted_hartree_fock(system, pkg.hf.HartreeFockConfig(iters, 1e-14), ints, Tap())
    P = seen["P"]
    nocc = system.n_electrons() // 2
    w, v = np.linalg.eigh(P)
    keep = np.argsort(w)[::-1][:nocc]
    L = v[:, keep] * np.sqrt(np.maximum(w[keep], 0.0))
    err = float(np.max(np.abs(L @ L.T - P)))
    out.mkdir(parents=True, exist_ok=True)
    f = out / f"waters{n}_scf{iters}_density_factor.npz"
    np.savez_compressed(f, L=L)
    print(f"{f}: L {L.shape}, max|LL^T - P| = {err:.2e}, {f.stat().st_size / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
