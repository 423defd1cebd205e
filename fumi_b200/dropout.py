"""Host mirror of the counter-based dropout mask used inside the episode kernels
(fumi_mask_hash in csrc/common.cuh).  The reference draws Bernoulli masks from the torch global
generator (nn.Dropout inside im_net, fumi.py:93-99), which cannot be reproduced by a batched
device kernel; masks here are a pure function of (seed, task, pass, layer, row, col)
so that forward and backward regenerate them and tests can inject the same masks into the oracle."""
import numpy as np

_M = np.uint64(0xFFFFFFFFFFFFFFFF)


_U32 = np.uint32


def _lowbias32(x):
    x = np.asarray(x, np.uint32)
    with np.errstate(over="ignore"):
        x = x ^ (x >> _U32(16)); x = x * _U32(0x7FEB352D)
        x = x ^ (x >> _U32(15)); x = x * _U32(0x846CA68B)
        x = x ^ (x >> _U32(16))
    return x


def mask_base(seed, task, pas, layer):
    """fumi_mask_base of csrc/common.cuh."""
    seed, task = int(seed), int(task)
    with np.errstate(over="ignore"):
        x = _lowbias32(_U32(seed & 0xFFFFFFFF) ^ _U32(0x9E3779B9))
        x = _lowbias32(x ^ _U32((seed >> 32) & 0xFFFFFFFF))
        x = _lowbias32(x + _U32(task & 0xFFFFFFFF) * _U32(0x85EBCA6B))
        x = _lowbias32(x ^ _U32((task >> 32) & 0xFFFFFFFF))
        x = _lowbias32(x + _U32(pas) * _U32(0x9E3779B1) + _U32(layer) * _U32(0x61C88647))
    return x


def mask_array(seed, task, pas, layer, rows, cols, p):
    """[rows, cols] float32 mask with entries 0 or 1/(1-p): one 32-bit hash per (row, column pair); the even
    column uses its low 16 bits, the odd one the high 16 bits; kept iff field >= floor(p * 65536)."""
    r, c = np.meshgrid(np.arange(rows, dtype=np.uint32), np.arange(cols, dtype=np.uint32), indexing="ij")
    base = mask_base(seed, task, pas, layer)
    with np.errstate(over="ignore"):
        h = _lowbias32(base + r * _U32(0xC2B2AE35) + (c >> _U32(1)) * _U32(0x27D4EB2F))
    field = np.where((c & _U32(1)) == 1, h >> _U32(16), h & _U32(0xFFFF))
    thr = _U32(int(np.float32(p) * np.float32(65536.0)))
    keep = field >= thr
    return keep.astype(np.float32) * np.float32(1.0 / (1.0 - np.float32(p)))
