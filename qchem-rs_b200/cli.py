"""Command-line driver mirroring qchem-cli (qchem-cli/src/main.rs:10-62): subcommands `rhf` and `uhf` with
`--basis-set/-b`, `--molecule/-m`, `--max-iterations` (100), `--epsilon` (1e-6); `uhf` adds `--charge/-c` and
`--spin-multiplicity/-s`.  The reference parses the last two and ignores them (main.rs:111-116, uhf.rs:43-45); here they
are HONOURED (labelled extension): n_alpha - n_beta = multiplicity - 1, n_alpha + n_beta = sum(Z) - charge;
multiplicity 0 (the reference's default) keeps the reference semantics n_alpha = n_beta = n_electrons / 2.

Extra switches of this engine: --backend host|device (SCF loop in numpy around the engine, or the device-resident
qcf_scf_step loop), --gpus N (qcf_opts.n_gpus), --deterministic, --incremental K (full rebuild every K-th iteration).
Energies are printed to full precision (the reference prints 3 decimals, main.rs:102-104).

    python -m qcpkg_cli rhf -b data/basis/6-31G.json -m data/mol/benzene.json          (see qchem_cli.py at the repo root)
"""
from __future__ import annotations

import argparse
import sys
import time

from . import engine, hf
from .basis import BasisSet, MolecularSystem


def build_parser():
    ap = argparse.ArgumentParser(prog="qchem-cli", description="Hartree-Fock on the B200 Fock-build engine")
    ap.add_argument("--verbose", "-v", action="store_true")
    sub = ap.add_subparsers(dest="command", required=True)
    for name in ("rhf", "uhf"):
        p = sub.add_parser(name)
        p.add_argument("--basis-set", "-b", required=True)
        p.add_argument("--molecule", "-m", required=True)
        p.add_argument("--max-iterations", type=int, default=100)
        p.add_argument("--epsilon", type=float, default=1e-6)
        p.add_argument("--backend", choices=["host", "device"], default="device")
        p.add_argument("--gpus", type=int, default=1)
        p.add_argument("--deterministic", action="store_true")
        p.add_argument("--incremental", type=int, default=0, metavar="K")
        p.add_argument("--tau", type=float, default=1e-12)
        if name == "uhf":
            p.add_argument("--charge", "-c", type=int, default=0)
            p.add_argument("--spin-multiplicity", "-s", type=int, default=0)
    return ap


def occupations(n_electrons: int, charge: int, multiplicity: int):
    """(n_alpha, n_beta).  multiplicity 0 = reference semantics (uhf.rs:43-45)."""
    if multiplicity == 0 and charge == 0:
        return n_electrons // 2, n_electrons // 2
    n = n_electrons - charge
    unpaired = max(multiplicity, 1) - 1
    if n < 0 or (n - unpaired) % 2 or n < unpaired:
        raise SystemExit(f"charge {charge} and multiplicity {multiplicity} are inconsistent with {n_electrons} electrons")
    return (n + unpaired) // 2, (n - unpaired) // 2


def main(argv=None) -> int:
    args = build_parser().parse_args(argv)
    basis = BasisSet.load(args.basis_set)
    system = MolecularSystem.load(args.molecule, basis)
    cfg = hf.HartreeFockConfig(args.max_iterations, args.epsilon)
    t0 = time.perf_counter()
    with engine.FockEngine(system, tau=args.tau, n_gpus=args.gpus, deterministic=args.deterministic) as eng:
        ints = eng.one_electron()
        if args.command == "rhf":
            if args.backend == "device":
                out = hf.restricted_hartree_fock_device(system, cfg, ints, eng, full_rebuild_every=args.incremental)
            else:
                builder = hf.IncrementalFock(eng, args.incremental) if args.incremental else eng
                out = hf.restricted_hartree_fock(system, cfg, ints, builder)
        else:
            na, nb = occupations(system.n_electrons(), args.charge, args.spin_multiplicity)
            if args.backend == "device":
                out = hf.unrestricted_hartree_fock_device(system, cfg, ints, eng, na, nb, full_rebuild_every=args.incremental)
            else:
                builder = hf.IncrementalFock(eng, args.incremental) if args.incremental else eng
                out = hf.unrestricted_hartree_fock(system, cfg, ints, builder, n_alpha=na, n_beta=nb)
        st = eng.stats()
    dt = time.perf_counter() - t0
    if out is None:
        print("hartree fock did not converge", file=sys.stderr)      # main.rs:107 panics here
        return 1
    print(f"hartree fock converged after {out.iterations} iterations, took {dt:.3f} s "
          f"(N = {st['n_basis']}, {st['n_devices']} GPU(s), last build {st['kernel_ms']:.3f} ms)")
    print(f"electronic energy: {out.electronic_energy:.10f}")
    print(f"nuclear repulsion: {out.nuclear_repulsion:.10f}")
    print(f"hartree fock energy: {out.total_energy():.10f}")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
