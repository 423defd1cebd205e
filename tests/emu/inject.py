"""TEST INFRASTRUCTURE ONLY: run the product's host code against the host emulation of its kernels.

The product (fumi_b200/) has no CPU path and no emulation switch: every compute front-end calls
`_lib.require_cuda`.  The CPU tests that exercise kernel *logic* without a GPU bind tests/emu/libfumi_emu.so (the
kernel sources compiled against tests/emu/cuda_emu.h) in place of libfumi_b200.so and lift that check from here.
"""
import contextlib

import build_emu
from fumi_b200 import _lib, engine


def enable():
    saved = (_lib._LIB, _lib.require_cuda, engine.DEFAULT_PRECISION)
    _lib.load(build_emu.build())
    _lib.require_cuda = lambda device, what: None
    engine.DEFAULT_PRECISION = 0          # the emulation has no tcgen05 / TMA
    return saved


def disable(saved):
    _lib._LIB, _lib.require_cuda, engine.DEFAULT_PRECISION = saved


@contextlib.contextmanager
def emulation():
    saved = enable()
    try:
        yield
    finally:
        disable(saved)
