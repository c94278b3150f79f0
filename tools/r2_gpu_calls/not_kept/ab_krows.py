"""A/B of the shared-memory exchange rows (KROWS instantiation of the block kernel) in ONE process: for every
configuration (a set of QCF_* environment knobs read by qcf_create) the whole-build device time at N = 1007, the
difference of G to the first configuration's G, and optionally the serialised per-class times (QCF_PROFILE=1).
  python tools/ab_krows.py [n_waters=53] [reps=5]

Record of GPU call 19 (profiles/r2_ab_call19_smem_exchange_rows.log).  The QCF_KROWS_* knobs exist only in the library
of commit ccb1038; the variant was measured slower and removed (profiles/README.md, "Experiments that were not kept")."""
import collections
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[3]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))
import qcpkg  # noqa: E402
import ab  # noqa: E402

pkg = qcpkg.load()

BIG = "1000000"
CONFIGS = [
    ("default", {}),
    ("krows_p1_b128", {"QCF_KROWS_MAX_PRIM": "1"}),
    ("krows_p9_b128", {"QCF_KROWS_MAX_PRIM": "9"}),
    ("krows_p36_b128", {"QCF_KROWS_MAX_PRIM": "36"}),
    ("krows_all_b128", {"QCF_KROWS_MAX_PRIM": BIG}),
    ("krows_p9_b64", {"QCF_KROWS_MAX_PRIM": "9", "QCF_KROWS_BLOCK": "64"}),
    ("krows_p36_b64", {"QCF_KROWS_MAX_PRIM": "36", "QCF_KROWS_BLOCK": "64"}),
    ("krows_p9_b128_smem34k", {"QCF_KROWS_MAX_PRIM": "9", "QCF_KROWS_SMEM": "40000"}),
    ("krows_p36_b128_smem34k", {"QCF_KROWS_MAX_PRIM": "36", "QCF_KROWS_SMEM": "40000"}),
    ("krows_p36_b128_kpt128", {"QCF_KROWS_MAX_PRIM": "36", "QCF_KETS_PER_THREAD": "128"}),
    ("profile_default", {"QCF_PROFILE": "1"}),
    ("profile_krows_p36_b128", {"QCF_PROFILE": "1", "QCF_KROWS_MAX_PRIM": "36"}),
]


def class_table(recs):
    agg = collections.defaultdict(lambda: [0.0, 0])
    for r in recs:
        agg[(r["la"], r["lb"], r["lc"], r["ld"])][0] += r["ms"]
        agg[(r["la"], r["lb"], r["lc"], r["ld"])][1] += r["quartets"]
    return agg


def main():
    import torch
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 53
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    bs = pkg.BasisSet.load(ROOT / "data" / "basis" / "6-31G_st.json")
    system = pkg.MolecularSystem.from_atoms(pkg.molecules.water_cluster(n), bs)
    flush = torch.empty(512 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")
    P = None
    G0 = None
    tables = {}
    for tag, env in CONFIGS:
        for k, v in env.items():
            os.environ[k] = v
        try:
            with pkg.engine.FockEngine(system, tau=1e-12) as eng:
                if P is None:
                    P = ab.density(system, eng, n)
                ms = []
                for r in range(reps + 2):
                    flush.zero_(); torch.cuda.synchronize()
                    G = eng.rhf(P)
                    if r >= 2:
                        ms.append(eng.stats()["kernel_ms"])
                st = eng.stats()
                if G0 is None:
                    G0 = G.copy()
                print(f"AB {tag} n={n} N={eng.n} kernel_ms min={min(ms):.3f} med={np.median(ms):.3f} quartets={st['quartets']:.4e} "
                      f"launches={st['launches']} max|G-G_default|={np.max(np.abs(G - G0)):.2e} symmetric={np.array_equal(G, G.T)}", flush=True)
                if env.get("QCF_PROFILE") == "1":
                    tables[tag] = class_table(eng.launch_profile())
        finally:
            for k in env:
                del os.environ[k]
    if len(tables) == 2:
        (ta, a), (tb, b) = tables.items()
        print(f"serialised per-class ms: {ta} | {tb}")
        for k in sorted(a, key=lambda k: -a[k][0]):
            print(f"{k}  {a[k][0]:8.2f}  {b[k][0]:8.2f}   quartets {a[k][1]:.3e}")
        print(f"total  {sum(v[0] for v in a.values()):8.2f}  {sum(v[0] for v in b.values()):8.2f}")


if __name__ == "__main__":
    main()
