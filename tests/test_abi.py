"""CPU-side checks of the drop-in boundary: libqcfock.so loads and exports every symbol that
include/qcfock.h declares; argument errors are reported through return codes (no compute calls)."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def declared_symbols():
    text = (ROOT / "include" / "qcfock.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qcf_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_boundary():
    syms = declared_symbols()
    for s in ("qcf_create", "qcf_build_rhf", "qcf_build_uhf", "qcf_build_jk", "qcf_stats", "qcf_last_error", "qcf_destroy"):
        assert s in syms


def test_library_exports_every_declared_symbol():
    from qchem_rs_b200 import engine
    so = ROOT / "qchem-rs_b200" / "libqcfock.so"
    if not so.exists():
        engine.build_library()
    lib = ctypes.CDLL(str(so))
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in qcfock.h but not exported"
    assert sorted(engine.EXPORTS) == declared_symbols()


def test_null_arguments_are_rejected_without_a_gpu():
    from qchem_rs_b200 import engine
    L = engine.lib()
    assert L.qcf_create(None, None, None) == -1
    assert L.qcf_nbasis(None) == -1
    assert L.qcf_build_rhf(None, None, None) == -1
    assert L.qcf_last_error(None) == b"null context"


def test_create_fails_loudly_without_cuda_device():
    """No CPU fallback: on a box without a GPU, qcf_create must fail with QCF_ERR_CUDA."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from helpers import load_system
    from qchem_rs_b200 import engine
    with pytest.raises(engine.FockError, match="CUDA|cuda"):
        engine.FockEngine(load_system("water", "STO-3G"))


def test_ctypes_mirrors_match_the_header_layout(tmp_path):
    """sizeof / offsetof of every struct in include/qcfock.h, as compiled by gcc, against the ctypes mirrors
    the Python binding (and the Rust #[repr(C)] structs of INTEGRATION.md) are written from."""
    import ctypes
    import subprocess
    from qchem_rs_b200 import engine
    from qchem_rs_b200.basis import CBasis
    src = tmp_path / "layout.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "qcfock.h"\n'
        'int main(void) {\n'
        '  printf("%zu %zu %zu %zu\\n", sizeof(qcf_basis), sizeof(qcf_opts), sizeof(qcf_stats_t), sizeof(qcf_launch_rec));\n'
        '  printf("%zu %zu %zu %zu\\n", offsetof(qcf_basis, cartesian), offsetof(qcf_opts, block_threads),\n'
        '         offsetof(qcf_stats_t, prim_pairs_kept), offsetof(qcf_launch_rec, ms));\n'
        '  return 0; }\n')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", str(ROOT / "include"), str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)], text=True).split()
    sizes = [int(x) for x in out[:4]]
    offs = [int(x) for x in out[4:]]
    assert sizes == [ctypes.sizeof(CBasis), ctypes.sizeof(engine.COpts), ctypes.sizeof(engine.CStats), ctypes.sizeof(engine.CLaunchRec)]
    assert offs == [CBasis.cartesian.offset, engine.COpts.block_threads.offset, engine.CStats.prim_pairs_kept.offset,
                    engine.CLaunchRec.ms.offset]
