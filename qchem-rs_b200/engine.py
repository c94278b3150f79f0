"""ctypes binding of libqcfock.so (include/qcfock.h) and the `FockEngine` wrapper the SCF drivers use.

This is the Python image of the Rust shim SURVEY.md 8b describes (`qcfock-sys` + a safe `FockEngine`
with Drop and Result): same entry points, same error convention.  There is no CPU fallback: if the
library is missing or no CUDA device is usable, construction raises.
"""
from __future__ import annotations

import ctypes
import subprocess
from pathlib import Path

import numpy as np

from .basis import CBasis, FlatBasis

_HERE = Path(__file__).resolve().parent
_SO = _HERE / "libqcfock.so"
_LIB = None
_dp = ctypes.POINTER(ctypes.c_double)

QCF_TAU_NONE = -1.0

EXPORTS = ["qcf_create", "qcf_nbasis", "qcf_build_rhf", "qcf_build_uhf", "qcf_build_jk", "qcf_build_rhf_dev",
           "qcf_build_uhf_dev", "qcf_build_rhf_incremental", "qcf_build_uhf_incremental", "qcf_eri_quartet",
           "qcf_schwarz", "qcf_boys", "qcf_fp64_peak", "qcf_stats", "qcf_device_times", "qcf_launch_profile",
           "qcf_one_electron", "qcf_last_error", "qcf_destroy", "qcf_scf_init", "qcf_scf_step", "qcf_scf_get",
           "qcf_system_load", "qcf_system_basis", "qcf_system_n_electrons", "qcf_system_n_basis",
           "qcf_system_nuclear_repulsion", "qcf_system_error", "qcf_system_free"]


class FockError(RuntimeError):
    pass


class COpts(ctypes.Structure):
    _fields_ = [("screen_tau", ctypes.c_double), ("device", ctypes.c_int), ("rank", ctypes.c_int),
                ("world_size", ctypes.c_int), ("block_threads", ctypes.c_int), ("n_gpus", ctypes.c_int),
                ("deterministic", ctypes.c_int)]


class CStats(ctypes.Structure):
    _fields_ = [("n_basis", ctypes.c_int), ("n_shells", ctypes.c_int), ("n_pairs", ctypes.c_int),
                ("n_groups", ctypes.c_int), ("quartets", ctypes.c_longlong), ("quartets_total", ctypes.c_longlong),
                ("model_flops", ctypes.c_double), ("kernel_ms", ctypes.c_double), ("total_ms", ctypes.c_double),
                ("launches", ctypes.c_int), ("prim_pairs", ctypes.c_longlong), ("prim_pairs_kept", ctypes.c_longlong),
                ("create_ms", ctypes.c_double), ("host_ms", ctypes.c_double), ("n_devices", ctypes.c_int),
                ("graph_launches", ctypes.c_int), ("rank_imbalance", ctypes.c_double)]


class CScfInfo(ctypes.Structure):
    _fields_ = [("iteration", ctypes.c_int), ("converged", ctypes.c_int), ("electronic_energy", ctypes.c_double),
                ("density_rms", ctypes.c_double), ("build_ms", ctypes.c_double), ("linalg_ms", ctypes.c_double),
                ("wall_ms", ctypes.c_double)]


class CLaunchRec(ctypes.Structure):
    _fields_ = [("la", ctypes.c_int), ("lb", ctypes.c_int), ("kab", ctypes.c_int), ("lc", ctypes.c_int),
                ("ld", ctypes.c_int), ("kcd", ctypes.c_int), ("nbra", ctypes.c_int), ("nket", ctypes.c_int),
                ("quartets", ctypes.c_longlong), ("flops_per_prim_quartet", ctypes.c_double), ("ms", ctypes.c_float)]


def build_library(force: bool = False, jobs: int = 8) -> Path:
    """Compile libqcfock.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    subprocess.check_call(["make", "-C", str(_HERE / "csrc"), f"-j{jobs}", "-s"] + (["-B"] if force else []))
    return _SO


def lib():
    global _LIB, _SO
    if _LIB is None:
        import os
        if os.environ.get("QCF_LIB"):          # an alternative build of the same library (kernel A/B runs)
            _SO = Path(os.environ["QCF_LIB"]).resolve()
        if not _SO.exists():
            raise FockError(f"{_SO} is missing: build it with `make -C qchem-rs_b200/csrc` "
                            "(there is no CPU fallback)")
        L = ctypes.CDLL(str(_SO))
        L.qcf_last_error.restype = ctypes.c_char_p
        L.qcf_last_error.argtypes = [ctypes.c_void_p]
        L.qcf_destroy.argtypes = [ctypes.c_void_p]
        L.qcf_destroy.restype = None
        L.qcf_create.argtypes = [ctypes.POINTER(CBasis), ctypes.POINTER(COpts), ctypes.POINTER(ctypes.c_void_p)]
        L.qcf_nbasis.argtypes = [ctypes.c_void_p]
        L.qcf_build_rhf.argtypes = [ctypes.c_void_p, _dp, _dp]
        L.qcf_build_uhf.argtypes = [ctypes.c_void_p, _dp, _dp, _dp, _dp]
        L.qcf_build_jk.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(_dp), ctypes.POINTER(_dp), ctypes.POINTER(_dp)]
        L.qcf_build_rhf_incremental.argtypes = [ctypes.c_void_p, _dp, _dp, ctypes.c_int]
        L.qcf_build_uhf_incremental.argtypes = [ctypes.c_void_p, _dp, _dp, _dp, _dp, ctypes.c_int]
        L.qcf_device_times.argtypes = [ctypes.c_void_p, ctypes.c_int, _dp]
        L.qcf_scf_init.argtypes = [ctypes.c_void_p, _dp, _dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.qcf_scf_step.argtypes = [ctypes.c_void_p, ctypes.c_double, ctypes.POINTER(CScfInfo)]
        L.qcf_scf_get.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, _dp]
        L.qcf_system_load.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p)]
        L.qcf_system_basis.argtypes = [ctypes.c_void_p]
        L.qcf_system_basis.restype = ctypes.POINTER(CBasis)
        L.qcf_system_n_electrons.argtypes = [ctypes.c_void_p]
        L.qcf_system_n_basis.argtypes = [ctypes.c_void_p]
        L.qcf_system_nuclear_repulsion.argtypes = [ctypes.c_void_p]
        L.qcf_system_nuclear_repulsion.restype = ctypes.c_double
        L.qcf_system_error.argtypes = [ctypes.c_void_p]
        L.qcf_system_error.restype = ctypes.c_char_p
        L.qcf_system_free.argtypes = [ctypes.c_void_p]
        L.qcf_system_free.restype = None
        L.qcf_build_rhf_dev.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.qcf_build_uhf_dev.argtypes = [ctypes.c_void_p] + [ctypes.c_void_p] * 5
        L.qcf_eri_quartet.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 4 + [_dp]
        L.qcf_schwarz.argtypes = [ctypes.c_void_p, _dp]
        L.qcf_one_electron.argtypes = [ctypes.c_void_p, _dp, _dp, _dp]
        L.qcf_boys.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, _dp, _dp]
        L.qcf_fp64_peak.argtypes = [ctypes.c_void_p, _dp]
        L.qcf_stats.argtypes = [ctypes.c_void_p, ctypes.POINTER(CStats)]
        L.qcf_launch_profile.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(CLaunchRec)]
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(_dp)


class NativeSystem:
    """`qcf_system`: basis + molecule loaded by the library's own JSON loaders (the native stand-in for
    BasisSet::load / MolecularSystem::load, main.rs:76-77).  Usable wherever a FlatBasis is (`ref()`, `n_basis`)."""

    def __init__(self, basis_json, molecule_json):
        self._lib = lib()
        self._sys = ctypes.c_void_p()
        rc = self._lib.qcf_system_load(str(basis_json).encode(), str(molecule_json).encode(), ctypes.byref(self._sys))
        if rc != 0:
            msg = self._lib.qcf_system_error(self._sys).decode() if self._sys else "qcf_system_load failed"
            if self._sys:
                self._lib.qcf_system_free(self._sys)
                self._sys = ctypes.c_void_p()
            raise FockError(f"qcf_system_load: {msg} (code {rc})")
        self.c = self._lib.qcf_system_basis(self._sys).contents
        self.n_basis = self._lib.qcf_system_n_basis(self._sys)
        self.n_electrons = self._lib.qcf_system_n_electrons(self._sys)
        self.nuclear_repulsion = self._lib.qcf_system_nuclear_repulsion(self._sys)
        ns, na = self.c.n_shells, self.c.n_atoms
        self.shell_l = np.ctypeslib.as_array(self.c.shell_l, shape=(ns,)).copy()
        self.shell_atom = np.ctypeslib.as_array(self.c.shell_atom, shape=(ns,)).copy()
        self.shell_nprim = np.ctypeslib.as_array(self.c.shell_nprim, shape=(ns,)).copy()
        self.shell_prim_off = np.ctypeslib.as_array(self.c.shell_prim_off, shape=(ns,)).copy()
        nprim = int(self.shell_prim_off[-1] + self.shell_nprim[-1])
        self.exps = np.ctypeslib.as_array(self.c.exps, shape=(nprim,)).copy()
        self.coefs = np.ctypeslib.as_array(self.c.coefs, shape=(nprim,)).copy()
        self.Z = np.ctypeslib.as_array(self.c.Z, shape=(na,)).copy()
        self.xyz = np.ctypeslib.as_array(self.c.xyz, shape=(3 * na,)).copy()

    def ref(self):
        return ctypes.byref(self.c)

    def close(self):
        if getattr(self, "_sys", None):
            self._lib.qcf_system_free(self._sys)
            self._sys = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class FockEngine:
    """Owns a `qcf_ctx`.  `rhf(P)` / `uhf(Pa, Pb)` are the per-iteration calls that stand in for
    rhf.rs:67-68 and uhf.rs:90-91."""

    def __init__(self, system_or_flat, tau: float = 0.0, device: int = 0, rank: int = 0, world_size: int = 1,
                 block_threads: int = 0, n_gpus: int = 1, deterministic: bool = False):
        if isinstance(system_or_flat, (FlatBasis, NativeSystem)):
            self.fb = system_or_flat
        else:
            self.fb = system_or_flat.flat()
        self.n = self.fb.n_basis
        self._lib = lib()
        self._ctx = ctypes.c_void_p()
        opts = COpts(tau, device, rank, world_size, block_threads, n_gpus, 1 if deterministic else 0)
        rc = self._lib.qcf_create(self.fb.ref(), ctypes.byref(opts), ctypes.byref(self._ctx))
        if rc != 0:
            msg = self._lib.qcf_last_error(self._ctx).decode() if self._ctx else "qcf_create failed"
            if self._ctx:
                self._lib.qcf_destroy(self._ctx)
                self._ctx = ctypes.c_void_p()
            raise FockError(f"qcf_create: {msg} (code {rc})")

    # -- helpers ------------------------------------------------------------------------------------
    def _check(self, rc, what):
        if rc != 0:
            raise FockError(f"{what}: {self._lib.qcf_last_error(self._ctx).decode()} (code {rc})")

    def close(self):
        if getattr(self, "_ctx", None):
            self._lib.qcf_destroy(self._ctx)
            self._ctx = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def handle(self):
        return self._ctx

    # -- Fock builds --------------------------------------------------------------------------------
    def rhf(self, P: np.ndarray) -> np.ndarray:
        P = np.ascontiguousarray(P, dtype=np.float64)
        G = np.empty((self.n, self.n), dtype=np.float64)
        self._check(self._lib.qcf_build_rhf(self._ctx, _p(P), _p(G)), "qcf_build_rhf")
        return G

    def uhf(self, Pa: np.ndarray, Pb: np.ndarray):
        Pa = np.ascontiguousarray(Pa, dtype=np.float64)
        Pb = np.ascontiguousarray(Pb, dtype=np.float64)
        Ga = np.empty((self.n, self.n)); Gb = np.empty((self.n, self.n))
        self._check(self._lib.qcf_build_uhf(self._ctx, _p(Pa), _p(Pb), _p(Ga), _p(Gb)), "qcf_build_uhf")
        return Ga, Gb

    def rhf_incremental(self, P: np.ndarray, reset: bool = False) -> np.ndarray:
        """G = G_prev + G(P - P_prev) with difference-density screening (qcf_build_rhf_incremental)."""
        P = np.ascontiguousarray(P, dtype=np.float64)
        G = np.empty((self.n, self.n), dtype=np.float64)
        self._check(self._lib.qcf_build_rhf_incremental(self._ctx, _p(P), _p(G), 1 if reset else 0), "qcf_build_rhf_incremental")
        return G

    def uhf_incremental(self, Pa: np.ndarray, Pb: np.ndarray, reset: bool = False):
        Pa = np.ascontiguousarray(Pa, dtype=np.float64)
        Pb = np.ascontiguousarray(Pb, dtype=np.float64)
        Ga = np.empty((self.n, self.n)); Gb = np.empty((self.n, self.n))
        self._check(self._lib.qcf_build_uhf_incremental(self._ctx, _p(Pa), _p(Pb), _p(Ga), _p(Gb), 1 if reset else 0),
                    "qcf_build_uhf_incremental")
        return Ga, Gb

    def jk(self, dens):
        nd = len(dens)
        dens = [np.ascontiguousarray(P, dtype=np.float64) for P in dens]
        J = [np.empty((self.n, self.n)) for _ in range(nd)]
        K = [np.empty((self.n, self.n)) for _ in range(nd)]
        arr = lambda xs: (_dp * nd)(*[_p(x) for x in xs])
        self._check(self._lib.qcf_build_jk(self._ctx, nd, arr(dens), arr(J), arr(K)), "qcf_build_jk")
        return J, K

    def rhf_dev(self, dP_ptr: int, dG_ptr: int, stream: int = 0):
        self._check(self._lib.qcf_build_rhf_dev(self._ctx, dP_ptr, dG_ptr, stream), "qcf_build_rhf_dev")

    def uhf_dev(self, dPa: int, dPb: int, dGa: int, dGb: int, stream: int = 0):
        self._check(self._lib.qcf_build_uhf_dev(self._ctx, dPa, dPb, dGa, dGb, stream), "qcf_build_uhf_dev")

    # -- parity / diagnostics -----------------------------------------------------------------------
    def eri_quartet(self, a, b, c, d) -> np.ndarray:
        nc = lambda l: (l + 1) * (l + 2) // 2
        shp = tuple(nc(int(self.fb.shell_l[s])) for s in (a, b, c, d))
        out = np.zeros(shp)
        self._check(self._lib.qcf_eri_quartet(self._ctx, int(a), int(b), int(c), int(d), _p(out)), "qcf_eri_quartet")
        return out

    def one_electron(self):
        """(S, T, V): molint::overlap / kinetic / nuclear (rhf.rs:41-43) evaluated on the device."""
        S, T, V = (np.zeros((self.n, self.n)) for _ in range(3))
        self._check(self._lib.qcf_one_electron(self._ctx, _p(S), _p(T), _p(V)), "qcf_one_electron")
        return S, T, V

    def schwarz(self) -> np.ndarray:
        ns = len(self.fb.shell_l)
        Q = np.zeros((ns, ns))
        self._check(self._lib.qcf_schwarz(self._ctx, _p(Q)), "qcf_schwarz")
        return Q

    def boys(self, mmax: int, T) -> np.ndarray:
        T = np.ascontiguousarray(np.atleast_1d(T), dtype=np.float64)
        F = np.zeros((len(T), mmax + 1))
        self._check(self._lib.qcf_boys(self._ctx, mmax, len(T), _p(T), _p(F)), "qcf_boys")
        return F

    def fp64_peak_tflops(self) -> float:
        v = ctypes.c_double()
        self._check(self._lib.qcf_fp64_peak(self._ctx, ctypes.byref(v)), "qcf_fp64_peak")
        return v.value

    def launch_profile(self) -> list:
        n = self._lib.qcf_launch_profile(self._ctx, 0, None)
        if n < 0:
            self._check(n, "qcf_launch_profile")
        recs = (CLaunchRec * max(n, 1))()
        self._check(min(self._lib.qcf_launch_profile(self._ctx, n, recs), 0), "qcf_launch_profile")
        return [{k: getattr(recs[i], k) for k, _ in CLaunchRec._fields_} for i in range(n)]

    # -- device-resident SCF (qcf_scf_*) ------------------------------------------------------------
    def scf_init(self, S, H, n_alpha: int, n_beta: int = 0, unrestricted: bool = False, full_rebuild_every: int = 0):
        S = np.ascontiguousarray(S, dtype=np.float64); H = np.ascontiguousarray(H, dtype=np.float64)
        self._check(self._lib.qcf_scf_init(self._ctx, _p(S), _p(H), 1 if unrestricted else 0, int(n_alpha), int(n_beta),
                                           int(full_rebuild_every)), "qcf_scf_init")

    def scf_step(self, epsilon: float) -> dict:
        info = CScfInfo()
        self._check(self._lib.qcf_scf_step(self._ctx, float(epsilon), ctypes.byref(info)), "qcf_scf_step")
        return {k: getattr(info, k) for k, _ in CScfInfo._fields_}

    def scf_get(self, what: str, spin: int = 0) -> np.ndarray:
        code = {"density": 0, "fock": 1, "g": 2, "orbital_energies": 3}[what]
        out = np.empty(self.n if code == 3 else (self.n, self.n), dtype=np.float64)
        self._check(self._lib.qcf_scf_get(self._ctx, code, int(spin), _p(out)), "qcf_scf_get")
        return out

    def device_times(self) -> list:
        """Device time (ms) each GPU of the context spent on its share of the last build."""
        buf = (ctypes.c_double * 16)()
        n = self._lib.qcf_device_times(self._ctx, 16, buf)
        if n < 0:
            self._check(n, "qcf_device_times")
        return [buf[i] for i in range(n)]

    def stats(self) -> dict:
        st = CStats()
        self._check(self._lib.qcf_stats(self._ctx, ctypes.byref(st)), "qcf_stats")
        return {k: getattr(st, k) for k, _ in CStats._fields_}
