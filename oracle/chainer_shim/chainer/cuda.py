"""Importable but unused on the hot path.  `cupy` is numpy so that `from chainer.cuda import cupy as cp` (rl_env.py:6)
imports and `cp.zeros` (rl_env.py:14) allocates a host array — the stand-in has no GPU path."""
import numpy as cupy  # noqa: F401


def to_gpu(x, *a, **k):
    raise RuntimeError("chainer stand-in has no GPU path")


def to_cpu(x):
    return x
