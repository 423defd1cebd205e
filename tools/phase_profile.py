"""Per-phase SM-cycle breakdown of the episode kernels (diagnostics; needs a GPU).
   python tools/phase_profile.py [tasks]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from fumi_b200 import _lib, utils  # noqa: E402
from fumi_b200.data.bank import FeatureBank  # noqa: E402
from fumi_b200.data.loader import EpisodeLoader  # noqa: E402
from fumi_b200.data.synth import class_split, make_bank  # noqa: E402
from fumi_b200.sampler import EpisodeSampler  # noqa: E402

NAMES = {0: "bwd prologue (W1/G planes, maxes)", 1: "bwd q: tile data -> smem, wait, Q0", 2: "bwd q: dZ1q row-per-warp, Q1",
         3: "bwd q: head sums, aW1/dZ0q/a_S gemms, atomics, Q2", 4: "bwd aW1 planes after query pass",
         5: "bwd s: records landed (U0)", 6: "bwd s: undo W1 gemm + planes", 7: "bwd s: r_dH0 planes, T1",
         8: "bwd s: r_dZ1 partials + r_W1/r_H0 (dZ1 parts), T2", 9: "bwd s: row chain, T3",
         11: "bwd s: head sums, r_H0/r_W1 gemms, atomics, a_S, T4", 12: "bwd s: fold a_W1 + next records issue",
         14: "bwd epilogue",
         20: "fwd prologue", 21: "fwd s: wait B1", 22: "fwd s: Z1 gemm, B2", 23: "fwd s: row-per-warp c, stash H0, B3",
         24: "fwd s: dhp sums, dZ0 gemm, S update", 25: "fwd s: W1 update gemm + planes",
         26: "fwd s: records, head/b1 update, next H0", 28: "fwd q: wait QB1",
         29: "fwd q: wait QB2", 30: "fwd q: row-per-warp scoring", 33: "fwd epilogue"}


def main():
    tasks = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
    test = len(sys.argv) > 2 and sys.argv[2] == "test"          # meta-test: 100 steps, no stash, no dropout
    K = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    dev = torch.device("cuda", 0)
    args = bench.make_args("fumi", 5, K, 20 if test else 32, 100 if test else 5, not test, dev, 2048, 768, tasks, 0.0 if test else 0.25)
    bank = make_bank(num_images=673 * 62, num_classes=673)
    cats = class_split(673)[0]
    sampler = EpisodeSampler(bank.cat_of, cats, 5, K, 20 if test else 32)
    fb = FeatureBank(feats=torch.from_numpy(bank.feats[sampler.ids]).to(dev), text=torch.from_numpy(bank.text[cats]).to(dev),
                     ids=sampler.ids, categories=cats)
    loader = EpisodeLoader(fb, sampler, tasks)
    torch.manual_seed(123)
    model = utils.init_model(args, {})
    opt = utils.init_optim(args, model)
    eng = model._get_engine(dev)
    eng.precision = 2
    sampler.new_iterator()
    b = loader.next_batch().to(dev)
    nst = 100 if test else 5
    for _ in range(2):
        eng.fumi_batch(model, b, steps=nst, step_size=0.01, train=not test)
    torch.cuda.synchronize()
    L = _lib.lib()
    L.fumi_debug_phase_profile(1)
    eng.fumi_batch(model, b, steps=nst, step_size=0.01, train=not test)
    torch.cuda.synchronize()
    out = np.zeros(64, np.uint64)
    L.fumi_debug_read_phases(_lib.ptr(out))
    L.fumi_debug_phase_profile(0)
    for lo, hi, name in ((20, 60, "forward"), (0, 15, "backward")):
        tot = float(out[lo:hi].sum())
        print(f"== {name}: {tot / 1e6:.1f} Mcycles summed over CTAs")
        for i in range(lo, hi):
            if out[i]:
                print(f"  {i:2d} {NAMES.get(i, '?'):38s} {100 * float(out[i]) / tot:5.1f}%")


if __name__ == "__main__":
    main()
