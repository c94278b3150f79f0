#!/usr/bin/env python
"""bench.py -- Fock-build benchmark of the B200 engine (BASELINE.json metric: Fock-build s/iter and
shell quartets/s at 1/2/4/8 B200, fraction of FP64 peak, host-CPU reference beside it).

One "step" = one Fock build G(P) (one SCF iteration's two-electron work) on the (H2O)_n / 6-31G*
cluster of SURVEY.md 8d (default n = 53, N = 1007 basis functions), with P the density of SCF
iteration `--scf-iters` of that system (not random: density-weighted screening is realistic).

  python bench.py --gpus N --steps K --warmup W            own arm (CUDA engine; torchrun for N > 1)
  python bench.py --impl reference ...                     the reference algorithm on the host cores

`value`  : unique shell quartets evaluated per second, whole job, P resident in HBM.
`e2e`    : same metric through the host API (P in pinned host memory -> H2D -> build -> allreduce -> D2H G, pinned).
`roofline`: SURVEY.md 8d model flops of the evaluated quartets / CUDA-event time / measured FP64 FMA peak.
The oracle is used ONLY in the cpu_baseline / --impl reference legs (it is the CPU restatement of the
reference; the reference itself needs cargo + the absent `molint` crate and cannot be built).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
import qcpkg  # noqa: E402

METRIC = "fock_build_shell_quartets_per_s"
UNIT = "quartets/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--waters", type=int, default=53, help="(H2O)_n cluster size; 53 -> N = 1007")
    ap.add_argument("--scf-iters", type=int, default=6, help="SCF iterations that produce the timed density")
    ap.add_argument("--tau", type=float, default=1e-12)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload_name(n, nbf):
    return f"(H2O)_{n} 6-31G* RHF, N={nbf}, one Fock build per step"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.stop = threading.Event()
        self.th = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.1)

    def __enter__(self):
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.th.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def hbm_note():
    """DRAM side of the roofline for the largest launch of the top class, (ps|ss) K = 3x3, from the committed
    ncu --set full capture (profiles/r1_final_block_kernel_1000.txt): 6.39 MB read+written in 0.588 ms."""
    peak = 6524.3
    try:
        peak = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
    except Exception:
        pass
    achieved = 6.39e6 / 0.588e-3 / 1e9
    return {"achieved_gbs": achieved, "peak_gbs": peak, "frac": achieved / peak,
            "source": "ncu capture in profiles/ (not measured live); the path is FP64-pipe bound, not HBM bound"}


def scf_density(pkg, system, ints, builder, iters):
    """Density that enters the Fock build of SCF iteration `iters` (0 = the Hueckel guess), produced by
    the reference's own RHF loop (qchem-rs_b200/hf.py restates rhf.rs:32-108) driving `builder`."""
    seen = {}

    class Tap:
        def rhf(self, P):
            seen["P"] = np.array(P, copy=True)
            return builder.rhf(P)
    out = pkg.hf.restricted_hartree_fock(system, pkg.hf.HartreeFockConfig(iters, 1e-14), ints, Tap())
    return seen["P"], out


def cpu_sample(pkg, fb, P, tau, budget_s):
    """Oracle direct-SCF Fock build (OpenMP) on a bounded sample: every `stride`-th bra pair."""
    from oracle import oracle_lib
    d = oracle_lib.DirectFock(fb, tau=tau)
    ncores = oracle_lib.num_threads()
    npair = len(fb.shell_l) * (len(fb.shell_l) + 1) // 2
    stride = max(1, npair // 64)
    t0 = time.perf_counter(); d.jk([P], stride=stride, offset=stride // 2); t1 = time.perf_counter()
    rate = d.last_quartets / max(t1 - t0, 1e-9)
    full = rate and (d.last_quartets * stride) / rate
    stride = max(1, int(np.ceil(full / budget_s))) if full else 1
    t0 = time.perf_counter(); d.jk([P], stride=stride, offset=stride // 2); t1 = time.perf_counter()
    q = d.last_quartets
    return {"value": q / (t1 - t0), "unit": UNIT, "cores": ncores, "kind": "port",
            "sample": f"every {stride}-th bra shell pair of the same build ({q} quartets, {t1 - t0:.1f} s); "
                      f"extrapolated full build {(t1 - t0) * stride:.1f} s/iter",
            "seconds_per_iter_extrapolated": (t1 - t0) * stride}


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port, direct-SCF form, all host threads)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pkg = qcpkg.load()
    from oracle import oracle_lib
    bs = pkg.BasisSet.load(ROOT / "data" / "basis" / "6-31G_st.json")
    system = pkg.MolecularSystem.from_atoms(pkg.molecules.water_cluster(args.waters), bs)
    fb = system.flat()
    n = fb.n_basis
    # density: superposition-free cheap stand-in is not allowed to differ from the own arm's workload, so
    # use the same SCF-iteration density when the CUDA engine is available; otherwise the Hueckel guess.
    ints = oracle_lib.one_electron(fb)
    P = None
    try:
        import torch
        if torch.cuda.is_available():
            with pkg.engine.FockEngine(system, tau=args.tau) as eng:
                P, _ = scf_density(pkg, system, ints, eng, args.scf_iters)
    except Exception:
        P = None
    dens_note = f"SCF iteration {args.scf_iters}"
    if P is None:
        S, T, V = ints
        x = pkg.hf.compute_transformation_matrix(S)
        P = pkg.hf.compute_hueckel_density(T + V, S, x, system.n_electrons() // 2, 2.0)
        dens_note = "Hueckel guess (no CUDA device for the SCF pre-iterations)"
    per_step = max(2.0, min(args.cpu_seconds, 120.0 / max(1, args.steps + args.warmup)))
    res = None
    times, quartets = [], []
    d = oracle_lib.DirectFock(fb, tau=args.tau)
    probe = cpu_sample(pkg, fb, P, args.tau, per_step)
    stride = max(1, int(round(probe["seconds_per_iter_extrapolated"] / per_step)))
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter(); d.jk([P], stride=stride, offset=stride // 2); t1 = time.perf_counter()
        if it >= args.warmup:
            times.append(t1 - t0); quartets.append(d.last_quartets)
    tot_t, tot_q = sum(times), sum(quartets)
    value = tot_q / tot_t
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": f"synthetic ({dens_note} density of a generated water cluster)",
            "config": {"workload": workload_name(args.waters, n), "tau": args.tau,
                       "sample": f"every {stride}-th bra shell pair per step"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": oracle_lib.num_threads(), "kind": "port",
                             "sample": f"every {stride}-th bra shell pair per step, {args.steps} steps, "
                                       f"extrapolated full build {stride * tot_t / args.steps:.1f} s/iter"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_b200(args):
    import torch
    import torch.distributed as dist
    pkg = qcpkg.load()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the engine has no CPU fallback); use --impl reference for the CPU arm")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    if world != args.gpus and rank == 0:
        print(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE", file=sys.stderr)

    bs = pkg.BasisSet.load(ROOT / "data" / "basis" / "6-31G_st.json")
    system = pkg.MolecularSystem.from_atoms(pkg.molecules.water_cluster(args.waters), bs)
    fb = system.flat()
    n = fb.n_basis
    eng = pkg.engine.FockEngine(system, tau=args.tau, device=local, rank=rank, world_size=world)
    fock = pkg.distributed.DeviceFock(eng, dev)
    ints = eng.one_electron()
    P, _ = scf_density(pkg, system, ints, fock, args.scf_iters)        # same on every rank (allreduced G)
    fock.dP[0].copy_(torch.from_numpy(P))
    peak = eng.fp64_peak_tflops()
    flush = torch.empty(512 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)   # 512 MB > 126 MB L2

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        q = fl = 0
        launches = 0
        barrier()
        for e0, e1 in ev:
            flush.zero_()
            e0.record()
            fn()
            e1.record()
            st = eng.stats()          # waits for the build; counters of this rank
            q += st["quartets"]; fl += st["model_flops"]; launches += st["launches"]
        barrier()
        ms = sum(e0.elapsed_time(e1) for e0, e1 in ev)
        t = torch.tensor([ms, float(q), fl, float(launches)], dtype=torch.float64, device=dev)
        if world > 1:
            tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            t[0] = tmax[0]
        ms, q, fl, launches = t.tolist()
        return ms, q, fl, int(launches)

    with ClockSampler(local) as clk:
        ms, q, fl, launches = timed(lambda: fock.rhf_device(), args.steps, args.warmup)
    clocks = clk.summary()
    # e2e: the density lives in pinned host memory (the engine's staging buffer); every step copies it to the
    # device, builds, all-reduces and reads the Fock matrix back into pinned host memory
    # (N = 1: the drop-in C-ABI call itself, qcf_build_rhf with caller-owned host buffers; N > 1: pinned buffers +
    # NCCL all-reduce, since the C ABI leaves the reduction to the caller)
    fock.hP[0].copy_(torch.from_numpy(P))
    e2e_call = (lambda: eng.rhf(P)) if world == 1 else (lambda: fock.rhf_pinned())
    ms_e, q_e, _, _ = timed(e2e_call, args.steps, max(1, args.warmup // 2))

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_sample(pkg, fb, P, args.tau, args.cpu_seconds)
    if rank == 0:
        value = q / (ms * 1e-3)
        achieved = fl / (ms * 1e-3) / 1e12
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": f"synthetic (SCF iteration {args.scf_iters} density of a generated water cluster)",
                "config": {"workload": workload_name(args.waters, n), "tau": args.tau, "parallelism": f"bra-pair split x{world}",
                           "l2": "flushed between timed steps (512 MB memset)", "quartets_per_step": q / args.steps,
                           "quartets_unscreened": eng.stats()["quartets_total"]},
                "e2e": {"value": q_e / (ms_e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 8 * n * n, "d2h_bytes_per_step": 8 * n * n,
                        "ms_per_step": ms_e / args.steps},
                "gpu_launches": launches,
                "roofline": {"bound": "fp64", "achieved": achieved, "peak": peak * world, "unit": "TFLOP/s",
                             "frac": achieved / (peak * world), "traffic": 6.39e6,
                             "hbm": hbm_note(),
                             "note": "achieved = SURVEY 8d model flops of the evaluated quartets / CUDA-event time, all eri_jk "
                                     "launches of the step; peak = FP64 FMA microbenchmark measured in this run (MEASURED_PEAKS.json "
                                     "has no FP64 figure; nominal 37.2); traffic = dram read+write bytes of the largest launch of the top class (ps|ss), "
                                     "ncu --set full capture in profiles/r1_final_block_kernel_1000.txt: the path is not HBM-bound"},
                "clocks": clocks}
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
