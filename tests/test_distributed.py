"""World-size-2 `gloo` tests (CPU) of the N>1 host logic: the rank partition of the bra-pair list plus one
all_reduce reproduces the single-rank Fock matrix and the same SCF.  The rank-local partial builder here is
the oracle's direct-SCF loop restricted to bra pairs ip % world == rank -- the same split rule the CUDA
engine applies per pair group (qcf_opts.rank / world_size); the CUDA partials are covered on the GPU by
test_bra_partition_sums_to_full_build."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
    import torch.distributed as dist
    from helpers import load_system, oracle_lib, random_symmetric_density
    from qchem_rs_b200 import hf, distributed
    os.environ["OMP_NUM_THREADS"] = "2"
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        system = load_system("water", "STO-3G")
        fb = system.flat()
        direct = oracle_lib.DirectFock(fb, tau=1e-12)

        class Partial:
            def partial_rhf(self, P):
                (J,), (K,) = direct.jk([P], stride=world, offset=rank)
                return J - 0.5 * K

            def partial_uhf(self, Pa, Pb):
                (Ja, Jb), (Ka, Kb) = direct.jk([Pa, Pb], stride=world, offset=rank)
                return Ja + Jb - Ka, Ja + Jb - Kb

        red = distributed.ReducedFock(Partial())
        P = random_symmetric_density(fb.n_basis, 21)
        Pb = random_symmetric_density(fb.n_basis, 22)
        G = red.rhf(P)
        Ga, Gb = red.uhf(P, Pb)
        out = hf.restricted_hartree_fock(system, hf.HartreeFockConfig(100, 1e-8), oracle_lib.one_electron(fb), red)
        np.savez(Path(out_dir) / f"r{rank}.npz", G=G, Ga=Ga, Gb=Gb, e=out.total_energy(), it=out.iterations)
    finally:
        dist.destroy_process_group()


def test_partition_plus_allreduce_matches_single_rank(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    from helpers import load_system, oracle_lib, random_symmetric_density
    from qchem_rs_b200 import hf
    system = load_system("water", "STO-3G")
    fb = system.flat()
    dense = oracle_lib.DenseFock(fb)
    P = random_symmetric_density(fb.n_basis, 21)
    Pb = random_symmetric_density(fb.n_basis, 22)
    ref = hf.restricted_hartree_fock(system, hf.HartreeFockConfig(100, 1e-8), oracle_lib.one_electron(fb), dense)
    res = [np.load(tmp_path / f"r{r}.npz") for r in range(world)]
    for r in res:
        np.testing.assert_allclose(r["G"], dense.rhf(P), atol=1e-11)
        np.testing.assert_allclose(r["Ga"], dense.uhf_one(P, Pb), atol=1e-11)
        np.testing.assert_allclose(r["Gb"], dense.uhf_one(Pb, P), atol=1e-11)
        assert int(r["it"]) == ref.iterations
        assert abs(float(r["e"]) - ref.total_energy()) < 1e-8
    # every rank ends with bit-identical matrices (all_reduce), so the replicated SCF stays in lock step
    np.testing.assert_array_equal(res[0]["G"], res[1]["G"])
