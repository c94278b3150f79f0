"""ctypes binding of oracle/liboracle.so -- TEST INFRASTRUCTURE (see the header of qc_oracle.cpp).

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs.  Matrices cross this boundary as column-major N x N float64, the nalgebra `DMatrix` layout the
reference uses (they are symmetric, so numpy's row-major view is the same matrix).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None

_dp = ctypes.POINTER(ctypes.c_double)


def build(force: bool = False) -> Path:
    so = _HERE / "liboracle.so"
    src = _HERE / "qc_oracle.cpp"
    if force or not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.check_call(["make", "-C", str(_HERE), "-s"], env={**os.environ})
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(str(build()))
        L.orc_nbasis.restype = ctypes.c_int
        L.orc_nuclear_repulsion.restype = ctypes.c_double
        L.orc_jk_direct.restype = ctypes.c_longlong
        L.orc_num_threads.restype = ctypes.c_int
        L.orc_num_procs.restype = ctypes.c_int
        _LIB = L
    return _LIB


def _mat(n):
    return np.zeros((n, n), dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(_dp)


def boys(mmax: int, T: float) -> np.ndarray:
    F = np.zeros(mmax + 1)
    lib().orc_boys(ctypes.c_int(mmax), ctypes.c_double(T), _p(F))
    return F


def one_electron(fb):
    """(S, T, V) of molint::overlap/kinetic/nuclear (rhf.rs:41-43)."""
    n = fb.n_basis
    S, T, V = _mat(n), _mat(n), _mat(n)
    L = lib()
    L.orc_overlap(fb.ref(), _p(S)); L.orc_kinetic(fb.ref(), _p(T)); L.orc_nuclear(fb.ref(), _p(V))
    return S, T, V


def nuclear_repulsion(fb) -> float:
    return lib().orc_nuclear_repulsion(fb.ref())


def eri_tensor(fb) -> np.ndarray:
    n = fb.n_basis
    eri = np.zeros((n, n, n, n), dtype=np.float64)
    lib().orc_eri_tensor(fb.ref(), _p(eri))
    return eri


def eri_shell_quartet(fb, a, b, c, d) -> np.ndarray:
    nc = lambda l: (l + 1) * (l + 2) // 2
    shp = tuple(nc(int(fb.shell_l[s])) for s in (a, b, c, d))
    out = np.zeros(shp, dtype=np.float64)
    lib().orc_eri_shell_quartet(fb.ref(), int(a), int(b), int(c), int(d), _p(out))
    return out


def schwarz(fb) -> np.ndarray:
    ns = len(fb.shell_l)
    Q = np.zeros((ns, ns))
    lib().orc_schwarz(fb.ref(), _p(Q))
    return Q


class DenseFock:
    """Reference-faithful stored-integral Fock builder (rhf.rs:45, 58-62, 152-167; uhf.rs:210-227)."""

    def __init__(self, fb):
        self.n = fb.n_basis
        self.eri = eri_tensor(fb)
        self.et = np.zeros(self.n ** 4)
        self.ready = ctypes.c_int(0)

    def rhf(self, P: np.ndarray) -> np.ndarray:
        G = _mat(self.n)
        P = np.ascontiguousarray(P, dtype=np.float64)
        lib().orc_fock_rhf_dense(ctypes.c_int(self.n), _p(P), _p(self.eri), _p(self.et),
                                 ctypes.byref(self.ready), _p(G))
        return G

    def uhf(self, Pa: np.ndarray, Pb: np.ndarray):
        """(G_alpha, G_beta): the two calls of uhf.rs:90-91 (both see the old densities)."""
        return self.uhf_one(Pa, Pb), self.uhf_one(Pb, Pa)

    def uhf_one(self, P1: np.ndarray, P2: np.ndarray) -> np.ndarray:
        """uhf.rs:210-227 for one spin: density_one = P1, density_two = P2."""
        G = _mat(self.n)
        P1 = np.ascontiguousarray(P1, dtype=np.float64); P2 = np.ascontiguousarray(P2, dtype=np.float64)
        lib().orc_fock_uhf_dense(ctypes.c_int(self.n), _p(P1), _p(P2), _p(self.eri), _p(G))
        return G


class DirectFock:
    """Direct-SCF oracle (OpenMP): J and K per density from unique shell quartets."""

    def __init__(self, fb, tau: float = 0.0):
        self.fb = fb
        self.n = fb.n_basis
        self.tau = tau
        self.Q = schwarz(fb)
        self.last_quartets = 0

    def jk(self, dens, stride: int = 1, offset: int = 0):
        nd = len(dens)
        dens = [np.ascontiguousarray(P, dtype=np.float64) for P in dens]
        J = [_mat(self.n) for _ in range(nd)]
        K = [_mat(self.n) for _ in range(nd)]
        arr = lambda xs: (_dp * nd)(*[_p(x) for x in xs])
        self.last_quartets = lib().orc_jk_direct(
            self.fb.ref(), ctypes.c_double(self.tau), ctypes.c_int(nd), arr(dens), arr(J), arr(K),
            _p(self.Q), ctypes.c_int(stride), ctypes.c_int(offset))
        return J, K

    def rhf(self, P):
        (J,), (K,) = self.jk([P])
        return J - 0.5 * K

    def uhf(self, Pa, Pb):
        (Ja, Jb), (Ka, Kb) = self.jk([Pa, Pb])
        return Ja + Jb - Ka, Ja + Jb - Kb


def jk_blocks_exact(fb, P, pairs):
    """Exact, UNSCREENED J and K blocks for the shell pairs `pairs` = [(sa, sb), ...]:
    J_ab = sum_cd P_cd (ab|cd), K_ab = sum_cd P_cd (ac|bd) over ALL c, d -- the reference's dense contraction
    (rhf.rs:58-62, 152-167; uhf.rs:210-227) restricted to one shell block.  Returns two lists of (na, nb) arrays."""
    nc = lambda l: (l + 1) * (l + 2) // 2
    P = np.ascontiguousarray(P, dtype=np.float64)
    sa = np.ascontiguousarray([p[0] for p in pairs], dtype=np.int32)
    sb = np.ascontiguousarray([p[1] for p in pairs], dtype=np.int32)
    shapes = [(nc(int(fb.shell_l[a])), nc(int(fb.shell_l[b]))) for a, b in pairs]
    tot = sum(x * y for x, y in shapes)
    J = np.zeros(tot); K = np.zeros(tot)
    ip = ctypes.POINTER(ctypes.c_int)
    lib().orc_jk_blocks_exact(fb.ref(), _p(P), ctypes.c_int(len(pairs)), sa.ctypes.data_as(ip), sb.ctypes.data_as(ip), _p(J), _p(K))
    outJ, outK, o = [], [], 0
    for x, y in shapes:
        outJ.append(J[o:o + x * y].reshape(x, y).copy()); outK.append(K[o:o + x * y].reshape(x, y).copy()); o += x * y
    return outJ, outK


def num_threads() -> int:
    return lib().orc_num_threads()


def num_procs() -> int:
    return lib().orc_num_procs()


def set_num_threads(n: int) -> None:
    """OpenMP thread count of the oracle.  torchrun exports OMP_NUM_THREADS=1 to its workers; bench.py's reference
    arm calls this with the host's core count so that the CPU arm uses the same cores at every --gpus N."""
    lib().orc_set_num_threads(ctypes.c_int(int(n)))
