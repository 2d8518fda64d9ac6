"""The C-ABI shared library loads on a CPU-only box and exports exactly what include/iago_b200.h declares."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "iago_b200.h")


def header_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"^IAGO_API[^;(]*?\b(iago_\w+)\s*\(", src, flags=re.M)))


@pytest.fixture(scope="module")
def so_path():
    from iago_b200 import build
    return build.build()


def test_header_declares_symbols():
    syms = header_symbols()
    assert len(syms) >= 15 and "iago_rollout" in syms and "iago_rollout_host" in syms


def test_library_exports_every_declared_symbol(so_path):
    lib = ctypes.CDLL(so_path)
    for name in header_symbols():
        assert hasattr(lib, name), f"{name} declared in the header but not exported"


def test_ctypes_table_matches_header(so_path):
    from iago_b200 import _lib
    assert sorted(_lib.SIGNATURES) == header_symbols()
    lib = _lib.load_library()
    assert lib.iago_abi_version() == 1


def test_no_unexpected_exports(so_path):
    out = subprocess.run(["nm", "-D", "--defined-only", so_path], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    extra = {s for s in exported if not s.startswith("iago_") and not s.startswith("_")}
    assert not extra, extra
    assert set(header_symbols()) <= exported


def test_sass_is_sm100a(so_path):
    out = subprocess.run(["cuobjdump", "-lelf", so_path], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_fails_loudly_without_gpu(so_path):
    """No CPU fallback: creating a context without a usable B200 must raise, not degrade."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import iago_b200
    with pytest.raises(iago_b200.IagoError):
        iago_b200.Engine(0)
    from iago_b200.mcts_self_play import Simulate
    from iago_b200 import boards
    with pytest.raises(iago_b200.IagoError):
        Simulate(boards.start_state())(1)
