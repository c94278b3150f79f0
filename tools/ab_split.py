"""The bra split of an 8-rank run, rank by rank on ONE GPU: per-rank device time of every rank's share (the slowest rank
is what an 8-GPU build takes), the modelled imbalance, and the sum of the ranks' partial matrices against the
single-rank build -- for the per-group split and the per-launch split (QCF_SPLIT_MIN_BRAS).
  python tools/ab_split.py [world=8] [reps=4]"""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))
import qcpkg  # noqa: E402
import ab  # noqa: E402

pkg = qcpkg.load()

CONFIGS = [
    ("default", {}),
    ("per_group", {"QCF_SPLIT_MIN_BRAS": "0"}),
    ("per_launch_148", {"QCF_SPLIT_MIN_BRAS": "148"}),
    ("per_launch_296", {"QCF_SPLIT_MIN_BRAS": "296"}),
    ("per_launch_74", {"QCF_SPLIT_MIN_BRAS": "74"}),
    ("per_launch_592", {"QCF_SPLIT_MIN_BRAS": "592"}),
]


def main():
    import torch
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    n = 53
    bs = pkg.BasisSet.load(ROOT / "data" / "basis" / "6-31G_st.json")
    system = pkg.MolecularSystem.from_atoms(pkg.molecules.water_cluster(n), bs)
    flush = torch.empty(512 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")
    with pkg.engine.FockEngine(system, tau=1e-12) as eng:
        P = ab.density(system, eng, n)
        G_full = eng.rhf(P)
        q_full = eng.stats()["quartets"]
    only = os.environ.get("AB_SPLIT_CONFIGS")
    for tag, env in CONFIGS:
        if only and tag not in only.split(","):
            continue
        os.environ.update(env)
        try:
            acc = np.zeros_like(G_full)
            q = 0
            meds, imb = [], 1.0
            for rank in range(world):
                with pkg.engine.FockEngine(system, tau=1e-12, rank=rank, world_size=world) as eng:
                    ms = []
                    for r in range(reps + 2):
                        flush.zero_(); torch.cuda.synchronize()
                        G = eng.rhf(P)
                        if r >= 2:
                            ms.append(eng.stats()["kernel_ms"])
                    st = eng.stats()
                    acc += G
                    q += st["quartets"]
                    imb = st["rank_imbalance"]
                    meds.append(float(np.median(ms)))
            print(f"SPLIT {tag} world={world} rank_ms={[round(m, 2) for m in meds]} max={max(meds):.3f} mean={np.mean(meds):.3f} "
                  f"imbalance_model={imb:.4f} max|sum_r G_r - G|={np.max(np.abs(acc - G_full)):.2e} quartets_equal={q == q_full}", flush=True)
        finally:
            for k in env:
                del os.environ[k]


if __name__ == "__main__":
    main()
