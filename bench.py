#!/usr/bin/env python
"""bench.py -- Fock-build benchmark of the B200 engine (BASELINE.json metric: Fock-build s/iter and
shell quartets/s at 1/2/4/8 B200, fraction of FP64 peak, host-CPU reference beside it).

One "step" = one Fock build G(P) (one SCF iteration's two-electron work) on the (H2O)_n / 6-31G*
cluster of SURVEY.md 8d (default n = 53, N = 1007 basis functions), with P the density of SCF
iteration `--scf-iters` of that system (not random: density-weighted screening is realistic).

  python bench.py --gpus N --steps K --warmup W            own arm (CUDA engine; torchrun for N > 1)
  python bench.py --impl reference ...                     the reference algorithm on the host cores

`value`  : unique shell quartets evaluated per second, whole job, P resident in HBM (one process per GPU, partial
           matrices summed by one NCCL all-reduce).
`e2e`    : same metric through the drop-in C-ABI call qcf_build_rhf with caller-owned HOST buffers (H2D of P and D2H of
           G inside the timed region).  N > 1: ONE context created with n_gpus = N drives all GPUs from rank 0's single
           host thread (in-library multi-GPU: peer copies of P, partial matrices summed over NVLink peer memory inside
           the finalize kernel); the other ranks wait on a CPU barrier.
`roofline`: SURVEY.md 8d model flops of the evaluated quartets / CUDA-event time / measured FP64 FMA peak.
The oracle is used ONLY in the cpu_baseline / --impl reference legs (it is the CPU restatement of the
reference; the reference itself needs cargo + the absent `molint` crate and cannot be built).  The reference arm
never imports the CUDA engine: its density comes from a committed fixture (tests/golden/) or the Hueckel guess.
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
GOLD = ROOT / "tests" / "golden"


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
    ap.add_argument("--deterministic", action="store_true", help="time the fixed-point (bitwise reproducible) mode")
    ap.add_argument("--scf", action="store_true", help="also time a whole device-resident SCF (reported in config.scf)")
    ap.add_argument("--scf-epsilon", type=float, default=1e-6, help="convergence threshold of the --scf runs (density rms, rhf.rs:87-91)")
    ap.add_argument("--scf-full-every", type=int, default=8, help="full rebuild period of the incremental --scf run")
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


def measured_hbm_peak():
    try:
        return float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def ncu_traffic():
    """dram read+write bytes per launch of the dominant kernel from this round's ncu --set full capture, if the summary
    was committed (profiles/r2_roofline_traffic.json); None otherwise -- never a typed-in constant."""
    f = ROOT / "profiles" / "r2_roofline_traffic.json"
    try:
        doc = json.loads(f.read_text())
        return float(doc["dram_bytes_per_launch"]), doc.get("kernel"), str(f.relative_to(ROOT))
    except Exception:
        return None, None, None


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


def fixture_path(waters, iters):
    return GOLD / f"waters{waters}_scf{iters}_density_factor.npz"


def load_fixture_density(waters, iters):
    """P = L L^T from the committed rank-n_occ factor of the SCF-iteration density (tools/make_bench_density.py)."""
    f = fixture_path(waters, iters)
    if not f.exists():
        return None
    L = np.load(f)["L"]
    return L @ L.T


def cpu_sample(fb, P, tau, budget_s, oracle_lib):
    """Oracle direct-SCF Fock build (OpenMP) on a bounded sample: every `stride`-th bra pair."""
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
            "note": "oracle/qc_oracle.cpp: a plain -O2 restatement (long-double Boys series, no primitive screening) -- a checker "
                    "first, a baseline second; the GPU/CPU ratio says nothing about kernel quality, the roofline fraction does",
            "sample": f"every {stride}-th bra shell pair of the same build ({q} quartets, {t1 - t0:.1f} s); "
                      f"extrapolated full build {(t1 - t0) * stride:.1f} s/iter",
            "seconds_per_iter_extrapolated": (t1 - t0) * stride}


def ref_faithful_leg(pkg, oracle_lib, gpu_engine_cls=None):
    """SURVEY.md 8d-(1) / BASELINE.md 3: the reference's own algorithm, single thread like the reference -- the one-off
    N^4 ERI tensor (molint::eri, rhf.rs:45), the repack (rhf.rs:58-62, inside the first contraction) and the per-iteration
    dense contraction (rhf.rs:152-167) -- on benzene / 6-31G (N = 66, the largest BASELINE config the reference's 8 N^4
    bytes allow in seconds), with the GPU build of the same molecule beside it."""
    bs = pkg.BasisSet.load(ROOT / "data" / "basis" / "6-31G.json")
    system = pkg.MolecularSystem.load(ROOT / "data" / "mol" / "benzene.json", bs)
    fb = system.flat()
    n = fb.n_basis
    nthreads = oracle_lib.num_threads()
    oracle_lib.set_num_threads(1)
    try:
        t0 = time.perf_counter(); dense = oracle_lib.DenseFock(fb); t1 = time.perf_counter()
        rng = np.random.default_rng(0)
        a = rng.normal(size=(n, n)); P = 0.5 * (a + a.T)
        dense.rhf(P)                                   # builds ET (rhf.rs:58-62) on the first call
        t2 = time.perf_counter(); dense.rhf(P); t3 = time.perf_counter()
    finally:
        oracle_lib.set_num_threads(nthreads)
    out = {"config": f"benzene 6-31G RHF, N={n}", "threads": 1, "eri_tensor_build_s": t1 - t0,
           "contraction_s_per_iter": t3 - t2, "tensor_bytes": 8 * n ** 4}
    if gpu_engine_cls is not None:
        with gpu_engine_cls(system, tau=1e-12) as eng:
            for _ in range(3):
                eng.rhf(P)
            st = eng.stats()
            out["gpu_build_ms"] = st["kernel_ms"]
            out["gpu_e2e_ms"] = st["total_ms"]
            out["gpu_create_ms"] = st["create_ms"]
    return out


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port, direct-SCF form, all host threads).  This arm never
    imports the CUDA engine."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pkg = qcpkg.load()
    from oracle import oracle_lib
    # torchrun exports OMP_NUM_THREADS=1 to its workers: use the host's cores explicitly so that the CPU arm is the same
    # at every --gpus N
    ncores = os.cpu_count() or oracle_lib.num_procs()
    oracle_lib.set_num_threads(ncores)
    bs = pkg.BasisSet.load(ROOT / "data" / "basis" / "6-31G_st.json")
    system = pkg.MolecularSystem.from_atoms(pkg.molecules.water_cluster(args.waters), bs)
    fb = system.flat()
    n = fb.n_basis
    P = load_fixture_density(args.waters, args.scf_iters)
    dens_note = f"SCF iteration {args.scf_iters} density, committed fixture {fixture_path(args.waters, args.scf_iters).name}"
    if P is None or P.shape != (n, n):
        S, T, V = oracle_lib.one_electron(fb)
        x = pkg.hf.compute_transformation_matrix(S)
        P = pkg.hf.compute_hueckel_density(T + V, S, x, system.n_electrons() // 2, 2.0)
        dens_note = "Hueckel guess density (no committed SCF-iteration fixture for this cluster size)"
    per_step = max(2.0, min(args.cpu_seconds, 120.0 / max(1, args.steps + args.warmup)))
    times, quartets = [], []
    d = oracle_lib.DirectFock(fb, tau=args.tau)
    probe = cpu_sample(fb, P, args.tau, per_step, oracle_lib)
    stride = max(1, int(round(probe["seconds_per_iter_extrapolated"] / per_step)))
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter(); d.jk([P], stride=stride, offset=stride // 2); t1 = time.perf_counter()
        if it >= args.warmup:
            times.append(t1 - t0); quartets.append(d.last_quartets)
    tot_t, tot_q = sum(times), sum(quartets)
    value = tot_q / tot_t
    sample = (f"every {stride}-th bra shell pair per step (rate measured on a 1/{stride} sample of the same build and "
              f"compared rate-to-rate; extrapolated full build {stride * tot_t / args.steps:.1f} s/iter)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": f"synthetic ({dens_note}; generated water cluster)",
            "config": {"workload": workload_name(args.waters, n), "tau": args.tau, "sample": sample,
                       "sampled": stride > 1, "seconds_per_full_build_extrapolated": stride * tot_t / args.steps},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": oracle_lib.num_threads(), "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "engine_library_loaded": pkg.engine._LIB is not None}
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
    cpu_group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        cpu_group = dist.new_group(backend="gloo")     # CPU barriers while one rank drives all GPUs (e2e)
    if world != args.gpus and rank == 0:
        print(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE", file=sys.stderr)

    bs = pkg.BasisSet.load(ROOT / "data" / "basis" / "6-31G_st.json")
    system = pkg.MolecularSystem.from_atoms(pkg.molecules.water_cluster(args.waters), bs)
    fb = system.flat()
    n = fb.n_basis
    eng = pkg.engine.FockEngine(system, tau=args.tau, device=local, rank=rank, world_size=world, deterministic=args.deterministic)
    fock = pkg.distributed.DeviceFock(eng, dev)
    ints = eng.one_electron()
    P, _ = scf_density(pkg, system, ints, fock, args.scf_iters)        # same on every rank (allreduced G)
    fixture = load_fixture_density(args.waters, args.scf_iters)
    fixture_diff = float(np.max(np.abs(fixture - P))) if fixture is not None and fixture.shape == P.shape else None
    fock.dP[0].copy_(torch.from_numpy(P))
    peak = eng.fp64_peak_tflops()
    flush = torch.empty(512 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)   # 512 MB > 126 MB L2

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps, warmup, stats_of, sync=True):
        for _ in range(warmup):
            fn()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        q = fl = host = 0.0
        launches = 0
        if sync:
            barrier()
        else:
            torch.cuda.synchronize(dev)
        for e0, e1 in ev:
            flush.zero_()
            e0.record()
            fn()
            e1.record()
            st = stats_of.stats()          # waits for the build; counters of this rank / context
            q += st["quartets"]; fl += st["model_flops"]; launches += st["launches"]; host += st["host_ms"]
        if sync:
            barrier()
        else:
            torch.cuda.synchronize(dev)
        ms = sum(e0.elapsed_time(e1) for e0, e1 in ev)
        return ms, q, fl, launches, host

    def reduce_over_ranks(ms, q, fl, launches):
        t = torch.tensor([ms, float(q), fl, float(launches)], dtype=torch.float64, device=dev)
        if world > 1:
            tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            t[0] = tmax[0]
        ms, q, fl, launches = t.tolist()
        return ms, q, fl, int(launches)

    with ClockSampler(local) as clk:
        ms, q, fl, launches, host_ms = timed(lambda: fock.rhf_device(), args.steps, args.warmup, eng)
    kernel_ms_rank = eng.stats()["kernel_ms"]
    ms, q, fl, launches = reduce_over_ranks(ms, q, fl, launches)
    clocks = clk.summary()
    rank_ms = [kernel_ms_rank]
    if world > 1:
        t = torch.zeros(world, dtype=torch.float64, device=dev); t[rank] = kernel_ms_rank
        dist.all_reduce(t)
        rank_ms = t.tolist()

    # ---- e2e: the drop-in C-ABI call with HOST buffers ------------------------------------------------------------
    st0 = eng.stats()
    e2e = None
    multi_check = None
    g_nccl = fock.rhf(P) if world > 1 else None       # all ranks take part in the all-reduce
    if world == 1:
        ms_e, q_e, _, _, _ = timed(lambda: eng.rhf(P), args.steps, max(1, args.warmup // 2), eng)
        e2e = {"value": q_e / (ms_e * 1e-3), "ms_per_step": ms_e / args.steps, "path": "qcf_build_rhf, one context, one GPU"}
    else:
        torch.cuda.synchronize(dev)
        dist.barrier(group=cpu_group)
        if rank == 0:
            # one context, n_gpus = world, driven from this host thread; the other ranks idle on a CPU barrier
            with pkg.engine.FockEngine(system, tau=args.tau, device=0, n_gpus=world, deterministic=args.deterministic) as eng_all:
                ms_e, q_e, _, _, _ = timed(lambda: eng_all.rhf(P), args.steps, max(2, args.warmup // 2), eng_all, sync=False)
                dev_ms = eng_all.device_times()
                st_all = eng_all.stats()
                g_all = eng_all.rhf(P)
            # the N-GPU results against a plain single-GPU build of the same density (outside every timed region)
            with pkg.engine.FockEngine(system, tau=args.tau, device=0) as eng_one:
                g_one = eng_one.rhf(P)
            multi_check = {"max_abs_diff_nccl_ranks_vs_1gpu": float(np.max(np.abs(g_nccl - g_one))),
                           "max_abs_diff_inlibrary_ngpu_vs_1gpu": float(np.max(np.abs(g_all - g_one))),
                           "symmetric": bool(np.array_equal(g_all, g_all.T) and np.array_equal(g_nccl, g_nccl.T))}
            e2e = {"value": q_e / (ms_e * 1e-3), "ms_per_step": ms_e / args.steps,
                   "path": f"qcf_build_rhf, ONE context with n_gpus={world} in rank 0's process (single host thread)",
                   "device_ms": dev_ms, "host_enqueue_ms": st_all["host_ms"], "create_ms": st_all["create_ms"],
                   "rank_imbalance_model": st_all["rank_imbalance"]}
        dist.barrier(group=cpu_group)

    scf = None
    if args.scf and rank == 0 and world == 1:
        cfg = pkg.hf.HartreeFockConfig(100, args.scf_epsilon)
        t0 = time.perf_counter(); out_dev = pkg.hf.restricted_hartree_fock_device(system, cfg, ints, eng); t1 = time.perf_counter()
        t2 = time.perf_counter(); out_inc = pkg.hf.restricted_hartree_fock_device(system, cfg, ints, eng, full_rebuild_every=args.scf_full_every); t3 = time.perf_counter()
        scf = {"epsilon": args.scf_epsilon}
        for name, o, dt in (("full_builds", out_dev, t1 - t0), (f"incremental_every_{args.scf_full_every}", out_inc, t3 - t2)):
            if o is not None:
                scf[name] = {"iterations": o.iterations, "wall_s": dt, "init_s": o.init_s,
                             "steps_wall_s": sum(s["wall_ms"] for s in o.steps) * 1e-3, "e_total": o.total_energy(),
                             "build_ms": [round(s["build_ms"], 2) for s in o.steps],
                             "linalg_ms_mean": float(np.mean([s["linalg_ms"] for s in o.steps]))}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle_lib
        oracle_lib.set_num_threads(os.cpu_count() or 1)
        cpu = cpu_sample(fb, P, args.tau, args.cpu_seconds, oracle_lib)
        try:
            cpu["ref_faithful"] = ref_faithful_leg(pkg, oracle_lib, pkg.engine.FockEngine)
        except Exception as exc:       # the N^4 leg is a side report; never lose the headline line to it
            cpu["ref_faithful"] = {"error": repr(exc)}
    if rank == 0:
        value = q / (ms * 1e-3)
        achieved = fl / (ms * 1e-3) / 1e12
        hbm_peak, hbm_kind = measured_hbm_peak()
        # algorithmic bytes of one build: pair data (8 doubles per kept primitive pair + ~64 B per pair), P in, G out, three accumulators
        alg_bytes = 64.0 * st0["prim_pairs_kept"] + 64.0 * st0["n_pairs"] + 5 * 8.0 * n * n
        traffic, traffic_kernel, traffic_src = ncu_traffic()
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": f"synthetic (SCF iteration {args.scf_iters} density of a generated water cluster)",
                "config": {"workload": workload_name(args.waters, n), "tau": args.tau,
                           "parallelism": f"cost-balanced bra-pair split x{world}",
                           "l2": "flushed between timed steps (512 MB memset)", "quartets_per_step": q / args.steps,
                           "quartets_unscreened": st0["quartets_total"], "prim_pairs": st0["prim_pairs"],
                           "prim_pairs_kept": st0["prim_pairs_kept"], "shell_pairs": st0["n_pairs"],
                           "create_s": st0["create_ms"] * 1e-3, "host_enqueue_ms_per_build": host_ms / args.steps,
                           "graph_launches_per_build": st0["graph_launches"], "deterministic": bool(args.deterministic),
                           "rank_kernel_ms": [round(x, 3) for x in rank_ms], "rank_imbalance_model": st0["rank_imbalance"],
                           "density_fixture_maxdiff": fixture_diff},
                "gpu_launches": launches,
                "roofline": {"bound": "fp64", "achieved": achieved, "peak": peak * world, "unit": "TFLOP/s",
                             "frac": achieved / (peak * world), "traffic": traffic,
                             "traffic_kernel": traffic_kernel, "traffic_source": traffic_src,
                             "hbm": {"algorithmic_bytes_per_build": alg_bytes, "achieved_gbs": alg_bytes / (ms / args.steps * 1e-3) / 1e9,
                                     "peak_gbs": hbm_peak, "peak_kind": hbm_kind,
                                     "frac": alg_bytes / (ms / args.steps * 1e-3) / 1e9 / hbm_peak},
                             "note": "achieved = SURVEY 8d model flops of the evaluated quartets / CUDA-event time of the whole step "
                                     "(all eri_jk launches); peak = FP64 FMA microbenchmark measured in this run (MEASURED_PEAKS.json "
                                     "has no FP64 figure; nominal 37.2 TFLOP/s); hbm = algorithmic bytes of one build (pair data + P + G + "
                                     "accumulators) / step time against the measured copy bandwidth: the path is FP64-pipe bound, not HBM bound; "
                                     "traffic = dram bytes per launch of the dominant kernel from this round's ncu capture (null if none committed)"},
                "clocks": clocks}
        if e2e is not None:
            e2e.update({"unit": UNIT, "h2d_bytes_per_step": 8 * n * n, "d2h_bytes_per_step": 8 * n * n})
            line["e2e"] = e2e
        if multi_check is not None:
            line["config"]["multi_gpu_check"] = multi_check
        if scf is not None:
            line["config"]["scf"] = scf
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
