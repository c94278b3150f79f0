"""Multi-GPU Fock build: one process per GPU, a static split of the bra-pair list, one allreduce.

SURVEY.md 8e: shell quartets are independent, only the N x N accumulators are shared.  Every rank holds
the full pair data and a full replica of P, evaluates its share of every (class, primitive-count) group's
bra-pair list -- a static split balanced by modelled cost (primitive quartets x class op count x length of
the Schwarz prefix; qcf_opts.rank / world_size, engine.cu make_plan) -- and produces a PARTIAL G.  The
partials are summed with a single `all_reduce(SUM)` (NCCL over NVLink on the GPU box; gloo in the CPU
tests).  There is no other communication on the path.

This module is the multi-PROCESS route (one process per GPU, torchrun).  The single-process route -- one
context created with qcf_opts.n_gpus = N driving all GPUs from one host thread, partial matrices summed over
NVLink peer memory inside the finalize kernel -- lives entirely inside libqcfock.so (`FockEngine(n_gpus=N)`).

The reduce step is written against `torch.distributed` only, so the host-side logic (partition +
reduction) is testable on CPU with world_size 2 and any object that has `partial_rhf` / `partial_uhf`.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


class HostPartialAdapter:
    """Wraps a builder whose `rhf(P)` / `uhf(Pa, Pb)` return this rank's partial matrices on the host."""

    def __init__(self, builder):
        self.b = builder

    def partial_rhf(self, P):
        return self.b.rhf(P)

    def partial_uhf(self, Pa, Pb):
        return self.b.uhf(Pa, Pb)


class ReducedFock:
    """Fock builder for the SCF drivers (`rhf(P)`, `uhf(Pa, Pb)`) over host-resident partial builders:
    calls the rank-local partial build and sums the partial matrices over the process group."""

    def __init__(self, partial, device: str = "cpu"):
        self.partial = partial
        self.device = device

    def _reduce(self, mats):
        _, ws = world()
        if ws == 1:
            return mats
        t = torch.from_numpy(np.stack(mats)).to(self.device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        out = t.cpu().numpy()
        return [out[i] for i in range(len(mats))]

    def rhf(self, P):
        return self._reduce([self.partial.partial_rhf(P)])[0]

    def uhf(self, Pa, Pb):
        ga, gb = self.partial.partial_uhf(Pa, Pb)
        ga, gb = self._reduce([ga, gb])
        return ga, gb


class DeviceFock:
    """The product path at N GPUs: density replicated in HBM, `qcf_build_*_dev` on torch's current
    stream, `all_reduce` (NCCL) on the same stream, result stays on the device unless asked for.

    `engine` is a `FockEngine` created with (device=local_rank, rank=rank, world_size=world_size)."""

    def __init__(self, engine, device: torch.device):
        self.eng = engine
        self.dev = device
        n = engine.n
        self.n = n
        self.dP = torch.zeros((2, n, n), dtype=torch.float64, device=device)
        self.dG = torch.zeros((2, n, n), dtype=torch.float64, device=device)
        self.hP = torch.zeros((2, n, n), dtype=torch.float64).pin_memory()
        self.hG = torch.zeros((2, n, n), dtype=torch.float64).pin_memory()

    # -- device-resident ---------------------------------------------------------------------------
    def rhf_device(self):
        """G[0] = sum over ranks of partial G(P = dP[0]); asynchronous on the current stream."""
        s = torch.cuda.current_stream(self.dev).cuda_stream
        self.eng.rhf_dev(self.dP[0].data_ptr(), self.dG[0].data_ptr(), s)
        if world()[1] > 1:
            dist.all_reduce(self.dG[0], op=dist.ReduceOp.SUM)
            self._symmetrize(self.dG[0])
        return self.dG[0]

    @staticmethod
    def _symmetrize(g):
        """Every rank's partial matrix is exactly symmetric, but a ring / tree all-reduce sums element (i, j) and
        element (j, i) in different rank orders, so the reduced matrix is symmetric only to rounding (1e-16).  The
        reference's contract is an exactly symmetric G (utils.rs:7-13): average the two triangles."""
        g.copy_(0.5 * (g + g.transpose(-1, -2)))

    def uhf_device(self):
        s = torch.cuda.current_stream(self.dev).cuda_stream
        self.eng.uhf_dev(self.dP[0].data_ptr(), self.dP[1].data_ptr(), self.dG[0].data_ptr(), self.dG[1].data_ptr(), s)
        if world()[1] > 1:
            dist.all_reduce(self.dG, op=dist.ReduceOp.SUM)
            self._symmetrize(self.dG)
        return self.dG

    # -- host API (what the SCF drivers call): H2D, build, allreduce, D2H --------------------------
    def rhf(self, P: np.ndarray) -> np.ndarray:
        self.hP[0].copy_(torch.from_numpy(np.ascontiguousarray(P, dtype=np.float64)))
        self.dP[0].copy_(self.hP[0], non_blocking=True)
        self.rhf_device()
        self.hG[0].copy_(self.dG[0], non_blocking=True)
        torch.cuda.current_stream(self.dev).synchronize()
        return self.hG[0].numpy().copy()

    def rhf_pinned(self) -> torch.Tensor:
        """Same as `rhf`, for callers that keep the density in the engine's pinned staging buffer `hP[0]`
        (a `DMatrix` allocated there on the Rust side): H2D from pinned memory, build, all-reduce, D2H into
        the pinned `hG[0]`, which is returned as a view (valid until the next call)."""
        self.dP[0].copy_(self.hP[0], non_blocking=True)
        self.rhf_device()
        self.hG[0].copy_(self.dG[0], non_blocking=True)
        torch.cuda.current_stream(self.dev).synchronize()
        return self.hG[0]

    def uhf(self, Pa: np.ndarray, Pb: np.ndarray):
        self.hP[0].copy_(torch.from_numpy(np.ascontiguousarray(Pa, dtype=np.float64)))
        self.hP[1].copy_(torch.from_numpy(np.ascontiguousarray(Pb, dtype=np.float64)))
        self.dP.copy_(self.hP, non_blocking=True)
        self.uhf_device()
        self.hG.copy_(self.dG, non_blocking=True)
        torch.cuda.current_stream(self.dev).synchronize()
        g = self.hG.numpy().copy()
        return g[0], g[1]
