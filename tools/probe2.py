import time
import os
os.environ["QCF_PROFILE"] = "1"
import sys, os
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
import qcpkg
pkg = qcpkg.load()
from qchem_rs_b200 import hf, engine, molecules
from qchem_rs_b200.basis import BasisSet, MolecularSystem


def water_cluster(n, basis="6-31G_st"):
    bs = BasisSet.load(ROOT / "data" / "basis" / f"{basis}.json")
    return MolecularSystem.from_atoms(molecules.water_cluster(n), bs)

n = int(sys.argv[1]); iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
system = water_cluster(n)
fb = system.flat()
with engine.FockEngine(system, tau=1e-12) as eng:
    ints = eng.one_electron()
    last = {}
    class B:
        def rhf(self, P):
            G = eng.rhf(P); st = eng.stats()
            print(f"  build kernel_ms={st['kernel_ms']:.2f} quartets={st['quartets']:.3e} model_TF={st['model_flops']/st['kernel_ms']/1e9:.3f}", flush=True)
            last['p'] = eng.launch_profile()
            return G
    hf.restricted_hartree_fock(system, hf.HartreeFockConfig(iters, 1e-8), ints, B())
    recs = last['p']
    import collections
    tot = sum(r['ms'] for r in recs)
    print(f"total serialized ms {tot:.2f}")
    # by class
    agg = collections.defaultdict(lambda: [0.0, 0, 0.0])
    for r in recs:
        k = (r['la'], r['lb'], r['lc'], r['ld'])
        fl = r['quartets'] * (r['kab'] * r['kcd'] * r['flops_per_prim_quartet'])
        agg[k][0] += r['ms']; agg[k][1] += r['quartets']; agg[k][2] += fl
    print("class        ms     %   quartets   modelTF")
    for k, v in sorted(agg.items(), key=lambda x: -x[1][0]):
        print(f"{k}  {v[0]:8.2f} {100*v[0]/tot:5.1f} {v[1]:.3e} {v[2]/max(v[0],1e-9)/1e9:7.3f}")
    print("top launches")
    for r in sorted(recs, key=lambda r: -r["ms"])[:int(os.environ.get("TOPN", "40"))]:
        fl = r['quartets'] * (r['kab'] * r['kcd'] * r['flops_per_prim_quartet'])
        print(f"({r['la']}{r['lb']}|{r['lc']}{r['ld']}) K={r['kab']:2d}x{r['kcd']:2d} nbra={r['nbra']:6d} nket={r['nket']:6d} q={r['quartets']:.3e} ms={r['ms']:8.3f} TF={fl/max(r['ms'],1e-9)/1e9:7.3f} ns/q={1e6*r['ms']/max(r['quartets'],1):.2f}")
