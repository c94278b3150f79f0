"""A/B timing of the Fock build on one GPU: (H2O)_n / 6-31G*, SCF-iteration density, whole-build device time (CUDA
events, L2 flushed between builds) and optionally the serialised per-launch profile (QCF_PROFILE=1).
  python tools/ab.py [n_waters=53] [reps=5] [profile_topn=0]
Environment knobs read by the engine: QCF_LIB (alternative libqcfock.so), QCF_PS_MIN, QCF_KETS_PER_THREAD,
QCF_STREAMS, QCF_TARGET_CTAS, QCF_SERIAL_CAP, QCF_NO_GRAPH, QCF_DETERMINISTIC."""
import collections
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import qcpkg  # noqa: E402

pkg = qcpkg.load()


def density(system, eng, n, iters=6):
    if os.environ.get("AB_MOL"):
        n = -1
    f = ROOT / "tests" / "golden" / f"waters{n}_scf{iters}_density_factor.npz"
    if f.exists():
        L = np.load(f)["L"]
        return L @ L.T
    cache = ROOT / "gpurun_out" / f"waters{n}_scf{iters}_density_factor.npz"
    if cache.exists():
        L = np.load(cache)["L"]
        return L @ L.T
    seen = {}
    ints = eng.one_electron()

    class Tap:
        def rhf(self, P):
            seen["P"] = np.array(P, copy=True)
            return eng.rhf(P)
    pkg.hf.restricted_hartree_fock(system, pkg.hf.HartreeFockConfig(iters, 1e-14), ints, Tap())
    return seen["P"]


def main():
    import torch
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 53
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    topn = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    tag = os.environ.get("AB_TAG", "")
    bs = pkg.BasisSet.load(ROOT / "data" / "basis" / "6-31G_st.json")
    if os.environ.get("AB_MOL"):            # a molecule file of data/mol instead of a water cluster (n is ignored)
        system = pkg.MolecularSystem.load(ROOT / "data" / "mol" / (os.environ["AB_MOL"] + ".json"), bs)
    else:
        system = pkg.MolecularSystem.from_atoms(pkg.molecules.water_cluster(n), bs)
    flush = torch.empty(512 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")
    ngpus = int(os.environ.get("AB_NGPUS", "1"))
    rank, world = int(os.environ.get("AB_RANK", "0")), int(os.environ.get("AB_WORLD", "1"))   # one rank's share of a multi-process run
    with pkg.engine.FockEngine(system, tau=1e-12, n_gpus=ngpus, rank=rank, world_size=world) as eng:
        P = density(system, eng, n)
        ms, e2e, host = [], [], []
        for r in range(reps + 2):
            flush.zero_(); torch.cuda.synchronize()
            eng.rhf(P)
            st = eng.stats()
            if r >= 2:
                ms.append(st["kernel_ms"]); e2e.append(st["total_ms"]); host.append(st["host_ms"])
        peak = eng.fp64_peak_tflops()
        tf = st["model_flops"] / (np.median(ms) * 1e-3) / 1e12
        print(f"AB {tag} n={n} N={eng.n} kernel_ms min={min(ms):.3f} med={np.median(ms):.3f} e2e_med={np.median(e2e):.3f} "
              f"host_ms={np.median(host):.3f} quartets={st['quartets']:.4e} modelTF={tf:.2f} frac={tf / peak:.3f} peak={peak:.2f} "
              f"launches={st['launches']} graph={st['graph_launches']} create_ms={st['create_ms']:.0f} "
              f"ngpus={ngpus} device_ms={[round(x, 2) for x in eng.device_times()]} imbalance_model={st['rank_imbalance']:.4f}", flush=True)
        if topn and os.environ.get("QCF_PROFILE") == "1":
            recs = eng.launch_profile()
            tot = sum(r["ms"] for r in recs)
            print(f"total serialized ms {tot:.2f}")
            agg = collections.defaultdict(lambda: [0.0, 0, 0.0])
            for r in recs:
                k = (r["la"], r["lb"], r["lc"], r["ld"])
                fl = r["quartets"] * (r["kab"] * r["kcd"] * r["flops_per_prim_quartet"])
                agg[k][0] += r["ms"]; agg[k][1] += r["quartets"]; agg[k][2] += fl
            print("class        ms     %   quartets   modelTF")
            for k, v in sorted(agg.items(), key=lambda x: -x[1][0]):
                print(f"{k}  {v[0]:8.2f} {100 * v[0] / tot:5.1f} {v[1]:.3e} {v[2] / max(v[0], 1e-9) / 1e9:7.3f}")
            print("top launches")
            for r in sorted(recs, key=lambda r: -r["ms"])[:topn]:
                fl = r["quartets"] * (r["kab"] * r["kcd"] * r["flops_per_prim_quartet"])
                print(f"({r['la']}{r['lb']}|{r['lc']}{r['ld']}) K={r['kab']:2d}x{r['kcd']:2d} nbra={r['nbra']:6d} nket={r['nket']:6d} "
                      f"q={r['quartets']:.3e} ms={r['ms']:8.3f} TF={fl / max(r['ms'], 1e-9) / 1e9:7.3f} ns/q={1e6 * r['ms'] / max(r['quartets'], 1):.2f}")


if __name__ == "__main__":
    main()
