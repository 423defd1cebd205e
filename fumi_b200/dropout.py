"""Host mirror of the counter-based dropout mask used inside the episode kernels
(fumi_mask_hash in csrc/common.cuh).  The reference draws Bernoulli masks from the torch global
generator (nn.Dropout inside im_net, fumi.py:93-99), which cannot be reproduced by a batched
device kernel; masks here are a pure function of (seed, task, pass, layer, row, col)
so that forward and backward regenerate them and tests can inject the same masks into the oracle."""
import numpy as np

_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def mask_hash64(seed, task, pas, layer, row, col_group):
    """fumi_mask_hash64 of csrc/common.cuh: one 64-bit hash per (row, 4-column group)."""
    with np.errstate(over="ignore"):
        x = np.uint64(seed) ^ (np.uint64(task) * np.uint64(0x9E3779B97F4A7C15))
        x = x + ((np.uint64(pas) << np.uint64(40)) ^ (np.uint64(layer) << np.uint64(32))
                 ^ (np.asarray(row, np.uint64) << np.uint64(12)) ^ np.asarray(col_group, np.uint64))
        x ^= x >> np.uint64(30)
        x = x * np.uint64(0xBF58476D1CE4E5B9)
        x ^= x >> np.uint64(27)
        x = x * np.uint64(0x94D049BB133111EB)
        x ^= x >> np.uint64(31)
    return x


def mask_array(seed, task, pas, layer, rows, cols, p):
    """[rows, cols] float32 mask with entries 0 or 1/(1-p): column c uses the 16-bit field (c & 3) of the hash
    of its (row, 4-column group); kept iff field >= floor(p * 65536)."""
    r, c = np.meshgrid(np.arange(rows, dtype=np.uint64), np.arange(cols, dtype=np.uint64), indexing="ij")
    bits = mask_hash64(seed, task, pas, layer, r, c >> np.uint64(2))
    field = (bits >> (np.uint64(16) * (c & np.uint64(3)))) & np.uint64(0xFFFF)
    thr = np.uint64(int(np.float32(p) * np.float32(65536.0)))
    keep = field >= thr
    return keep.astype(np.float32) * np.float32(1.0 / (1.0 - np.float32(p)))
