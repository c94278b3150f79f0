"""What one rank of an 8-rank run does, timed on ONE GPU (qcf_opts.rank / world_size; the other ranks' shares are not
run): whole-share device time of several launch-shape settings in one process.
  python tools/ab_rankshare.py [world=8] [reps=6]"""
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
    ("block32", {"QCF_BLOCK": "32"}),
    ("block32_c148", {"QCF_BLOCK": "32", "QCF_TARGET_CTAS": "148"}),
    ("block128", {"QCF_BLOCK": "128"}),
    ("psmin16", {"QCF_PS_MIN": "16"}),
    ("serialcap1e6", {"QCF_SERIAL_CAP": "1e6"}),
    ("kpt16", {"QCF_KETS_PER_THREAD": "16"}),
    ("kpt64", {"QCF_KETS_PER_THREAD": "64"}),
    ("c37", {"QCF_TARGET_CTAS": "37"}),
    ("c148", {"QCF_TARGET_CTAS": "148"}),
    ("order1", {"QCF_ORDER": "1"}),
    ("nograph", {"QCF_NO_GRAPH": "1"}),
    ("block32_kpt64", {"QCF_BLOCK": "32", "QCF_KETS_PER_THREAD": "64"}),
    ("psmin16_serialcap1e6", {"QCF_PS_MIN": "16", "QCF_SERIAL_CAP": "1e6"}),
    ("streams16", {"QCF_STREAMS": "16"}),
    ("kpt8", {"QCF_KETS_PER_THREAD": "8"}),
    ("serialcap1.6e7", {"QCF_SERIAL_CAP": "1.6e7"}),
]


def main():
    import torch
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    n = 53
    bs = pkg.BasisSet.load(ROOT / "data" / "basis" / "6-31G_st.json")
    system = pkg.MolecularSystem.from_atoms(pkg.molecules.water_cluster(n), bs)
    flush = torch.empty(512 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")
    P = None
    for tag, env in CONFIGS:
        os.environ.update(env)
        try:
            line = []
            for rank in (0, 5 % world):
                with pkg.engine.FockEngine(system, tau=1e-12, rank=rank, world_size=world) as eng:
                    if P is None:
                        P = ab.density(system, eng, n)
                    ms = []
                    for r in range(reps + 2):
                        flush.zero_(); torch.cuda.synchronize()
                        eng.rhf(P)
                        if r >= 2:
                            ms.append(eng.stats()["kernel_ms"])
                    line.append(f"rank{rank} min={min(ms):.3f} med={np.median(ms):.3f}")
            print(f"RS {tag} world={world} " + " | ".join(line), flush=True)
        finally:
            for k in env:
                del os.environ[k]


if __name__ == "__main__":
    main()
