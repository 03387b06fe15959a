"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/cmdlmc_b200.h declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "cmdlmc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cmd_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from cmdlmc_b200 import build, _abi
    build.build()
    return _abi.lib()


def test_exports_every_declared_symbol(lib):
    names = declared_symbols()
    assert len(names) > 25
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_binding_covers_header(lib):
    from cmdlmc_b200 import _abi
    assert set(declared_symbols()) <= set(_abi.SIGNATURES), \
        sorted(set(declared_symbols()) - set(_abi.SIGNATURES))
    assert lib.cmd_abi_version() == 1


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import numpy as np
    import cmdlmc_b200 as cm
    from cmdlmc_b200._abi import CmdError
    with pytest.raises(CmdError) as e:
        cm.AtomBoxCubic([10.0, 10, 10])
    assert e.value.code == -2
    with pytest.raises(CmdError):
        cm.Fermi(0.06, 2.3, 0.1)(np.array([2.5]))


def test_product_never_imports_oracle():
    """The product package must never import, load or link the oracle (test infrastructure)."""
    pkg = os.path.join(ROOT, "cmdlmc_b200")
    bad = re.compile(r"(^\s*(from|import)\s+(\.*)oracle)|libcmdlmc_oracle|#include\s*[\"<].*oracle|"
                     r"oracle\.(oracle|ref_import)|orc_[a-z_]+\s*\(", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not bad.search(text), os.path.join(dirpath, f)
