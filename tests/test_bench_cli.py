"""CPU-side checks of bench.py: the reference arm runs without a GPU and prints the contract's JSON line;
the own arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import subprocess
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]


def test_reference_arm_prints_the_contract_line():
    import os
    # torchrun exports OMP_NUM_THREADS=1 to its workers; the reference arm must use the host's cores regardless, and it
    # must never map the CUDA engine
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--waters", "2", "--steps", "1",
                          "--warmup", "1", "--cpu-seconds", "2", "--scf-iters", "1"], capture_output=True, text=True, timeout=600,
                         env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["cpu_baseline"]["cores"] == os.cpu_count()
    assert line["engine_library_loaded"] is False
    assert line["impl"] == "reference" and line["metric"] == "fock_build_shell_quartets_per_s"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["gpu_launches"] == 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0


def test_own_arm_needs_a_gpu():
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--waters", "2", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode != 0
    assert "no CUDA device" in (out.stderr + out.stdout)
