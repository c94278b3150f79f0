#!/usr/bin/env python
"""qchem-cli stand-in:  python qchem_cli.py rhf -b data/basis/STO-3G.json -m data/mol/water.json   (qchem-rs_b200/cli.py)"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent))
import qcpkg  # noqa: E402

if __name__ == "__main__":
    pkg = qcpkg.load()
    from qchem_rs_b200 import cli
    raise SystemExit(cli.main())
