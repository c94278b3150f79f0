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


def test_scf_info_layout_matches_header(tmp_path):
    import subprocess
    from qchem_rs_b200 import engine
    src = tmp_path / "layout2.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "qcfock.h"\n'
                   'int main(void) { printf("%zu %zu %zu %zu\\n", sizeof(qcf_scf_info), offsetof(qcf_scf_info, wall_ms),\n'
                   '  offsetof(qcf_opts, deterministic), offsetof(qcf_stats_t, rank_imbalance)); return 0; }\n')
    exe = tmp_path / "layout2"
    subprocess.check_call(["gcc", "-I", str(ROOT / "include"), str(src), "-o", str(exe)])
    out = [int(x) for x in subprocess.check_output([str(exe)], text=True).split()]
    assert out == [ctypes.sizeof(engine.CScfInfo), engine.CScfInfo.wall_ms.offset, engine.COpts.deterministic.offset,
                   engine.CStats.rank_imbalance.offset]


def _flat_water():
    from helpers import load_system
    return load_system("water", "STO-3G").flat()


def test_create_rejects_bad_basis_arrays_before_touching_the_gpu():
    """Argument validation of qcf_create: more than 255 primitives in a shell (the primitive-pair index is
    16 bit), negative primitive offsets, non-positive exponents and l > 2 are QCF_ERR_ARG with a message."""
    import numpy as np
    from qchem_rs_b200 import engine
    L = engine.lib()

    def create(fb):
        ctx = ctypes.c_void_p()
        rc = L.qcf_create(fb.ref(), None, ctypes.byref(ctx))
        msg = L.qcf_last_error(ctx).decode() if ctx else ""
        if ctx:
            L.qcf_destroy(ctx)
        return rc, msg

    fb = _flat_water()
    fb.shell_nprim[0] = 300
    rc, msg = create(fb)
    assert rc == -1 and "255" in msg
    fb = _flat_water()
    fb.shell_prim_off[1] = -3
    rc, msg = create(fb)
    assert rc == -1 and "shell_prim_off" in msg
    fb = _flat_water()
    fb.exps[2] = -1.0
    rc, msg = create(fb)
    assert rc == -1 and "exponent" in msg
    fb = _flat_water()
    fb.shell_l[0] = 3
    rc, msg = create(fb)
    assert rc == -1 and "angular" in msg
    fb = _flat_water()
    fb.c.cartesian = 0
    rc, msg = create(fb)
    assert rc == -1 and "Cartesian" in msg


@pytest.mark.parametrize("mol,basis", [("water", "STO-3G"), ("benzene", "6-31G"), ("caffeine", "6-31G_st"), ("hydrogen", "STO-3G"),
                                       ("oxygen", "6-31G"), ("ethylene", "6-31G")])
def test_native_loaders_match_python_loaders(mol, basis):
    """qcf_system_load (C++ JSON loaders behind the C ABI, stand-in for BasisSet::load / MolecularSystem::load,
    main.rs:76-77) on the reference's own data files against the Python loaders the other tests use."""
    import numpy as np
    from helpers import load_system, DATA
    from qchem_rs_b200 import engine, hf
    system = load_system(mol, basis)
    fb = system.flat()
    nat = engine.NativeSystem(DATA / "basis" / f"{basis}.json", DATA / "mol" / f"{mol}.json")
    assert nat.n_basis == fb.n_basis and nat.n_electrons == system.n_electrons()
    np.testing.assert_array_equal(nat.Z, fb.Z)
    np.testing.assert_array_equal(nat.shell_l, fb.shell_l)
    np.testing.assert_array_equal(nat.shell_atom, fb.shell_atom)
    np.testing.assert_array_equal(nat.shell_nprim, fb.shell_nprim)
    np.testing.assert_array_equal(nat.shell_prim_off, fb.shell_prim_off)
    np.testing.assert_array_equal(nat.xyz, fb.xyz)
    np.testing.assert_array_equal(nat.exps, fb.exps)
    np.testing.assert_allclose(nat.coefs, fb.coefs, rtol=1e-15)
    assert nat.nuclear_repulsion == pytest.approx(hf.compute_nuclear_repulsion(system.atoms), rel=1e-15)


def test_native_loader_errors_are_reported(tmp_path):
    from helpers import DATA
    from qchem_rs_b200 import engine
    with pytest.raises(engine.FockError, match="cannot open"):
        engine.NativeSystem(tmp_path / "missing.json", DATA / "mol" / "water.json")
    bad = tmp_path / "bad.json"
    bad.write_text('[{"element": "8", "position": [0, 0]}]')
    with pytest.raises(engine.FockError, match="malformed atom"):
        engine.NativeSystem(DATA / "basis" / "STO-3G.json", bad)
    heavy = tmp_path / "u.json"
    heavy.write_text('[{"element": "92", "position": [0, 0, 0]}]')
    with pytest.raises(engine.FockError, match="no element"):
        engine.NativeSystem(DATA / "basis" / "STO-3G.json", heavy)


def test_cli_occupations_and_parser():
    """qchem_cli.py (stand-in for qchem-cli, main.rs:10-62): flags parse like the reference's; --charge and
    --spin-multiplicity are honoured, multiplicity 0 keeps the reference semantics (uhf.rs:43-45)."""
    from qchem_rs_b200 import cli
    assert cli.occupations(16, 0, 0) == (8, 8)        # O2, reference semantics
    assert cli.occupations(16, 0, 3) == (9, 7)        # triplet O2
    assert cli.occupations(10, 1, 2) == (5, 4)        # water cation doublet
    assert cli.occupations(10, -1, 2) == (6, 5)
    with pytest.raises(SystemExit):
        cli.occupations(10, 0, 2)                     # even electron count cannot be a doublet
    args = cli.build_parser().parse_args(["uhf", "-b", "b.json", "-m", "m.json", "-c", "1", "-s", "2", "--max-iterations", "50"])
    assert (args.command, args.charge, args.spin_multiplicity, args.max_iterations, args.epsilon) == ("uhf", 1, 2, 50, 1e-6)
    args = cli.build_parser().parse_args(["rhf", "--basis-set", "b.json", "--molecule", "m.json"])
    assert args.max_iterations == 100 and args.backend == "device" and args.gpus == 1


def test_every_entry_point_is_bound_by_the_rust_shim_and_documented():
    """The reference-side binding (rust/qcfock-sys/src/lib.rs, uncompiled here: no Rust toolchain) and INTEGRATION.md
    must name every function include/qcfock.h declares, so that neither drifts from the C ABI."""
    import re
    root = Path(__file__).resolve().parents[1]
    header = (root / "include" / "qcfock.h").read_text()
    declared = sorted(set(re.findall(r"\b(qcf_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 25
    shim = (root / "rust" / "qcfock-sys" / "src" / "lib.rs").read_text()
    doc = (root / "INTEGRATION.md").read_text()
    assert [f for f in declared if not re.search(rf"\bfn {f}\s*\(", shim)] == []
    assert [f for f in declared if f not in doc] == []
