"""The product package never touches the oracle, the reference tree, or a CPU compute fallback."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "iago_b200")


def product_sources():
    for dp, dn, fn in os.walk(PKG):
        dn[:] = [d for d in dn if d not in ("build", "__pycache__")]
        for f in fn:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                yield os.path.join(dp, f)


def test_product_does_not_import_oracle_or_reference():
    bad = []
    for p in product_sources():
        src = open(p).read()
        if re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M) or "/root/reference" in src.replace(
                "/root/reference/", "REFDOC/"):
            bad.append(p)
        if "libothello_oracle" in src or "chainer_shim" in src:
            bad.append(p)
    assert not bad, bad


def test_no_triton_or_compile_in_product():
    for p in product_sources():
        src = open(p).read()
        assert "import triton" not in src and "torch.compile" not in src and "tilelang" not in src, p


def test_boards_roundtrip():
    import numpy as np
    from iago_b200 import boards
    rng = np.random.default_rng(0)
    s = rng.integers(0, 3, size=(100, 8, 8)).astype(np.float32)
    p1, p2 = boards.to_bitboards(s)
    assert (boards.from_bitboards(p1, p2) == s).all()
    q1, q2 = boards.to_bitboards(boards.start_state())
    assert int(q1[0]) == boards.START_P1 and int(q2[0]) == boards.START_P2
    assert boards.mask_to_actions((1 << 19) | (1 << 26) | (1 << 37) | (1 << 44)) == [19, 26, 37, 44]
