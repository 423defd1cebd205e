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

NAMES = {0: "bwd prologue", 1: "bwd q: tile loads", 2: "bwd q: a_head,dZ1q", 3: "bwd q: aW1/dZ0q gemms+atomics",
         4: "bwd q: a_S gemm", 5: "bwd s: loads+undo W1", 6: "bwd s: tile loads", 7: "bwd s: r_dH1/r_W1/r_H0 gemms",
         8: "bwd s: r_dL,r_head", 9: "bwd s: jacobian,r_H1", 10: "bwd s: r_Z1,r_head", 11: "bwd s: r_H0/r_W1 gemms+atomics",
         12: "bwd s: fold+reload bZ", 13: "bwd s: a_S gemm", 14: "bwd epilogue",
         20: "fwd prologue", 21: "fwd s: H0", 22: "fwd s: H1", 23: "fwd s: logits+softmax", 24: "fwd s: (unused)",
         25: "fwd s: dhp,dZ1", 26: "fwd s: dZ0 gemm,S,stash", 27: "fwd s: W1 update gemm", 28: "fwd q: loads",
         29: "fwd q: H0", 30: "fwd q: H1", 31: "fwd q: logits+softmax", 32: "fwd q: stash, tile end", 33: "fwd epilogue"}


def main():
    tasks = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
    dev = torch.device("cuda", 0)
    args = bench.make_args("fumi", 5, 5, 32, 5, True, dev, 2048, 768, tasks, 0.25)
    bank = make_bank(num_images=673 * 62, num_classes=673)
    cats = class_split(673)[0]
    sampler = EpisodeSampler(bank.cat_of, cats, 5, 5, 32)
    fb = FeatureBank(feats=torch.from_numpy(bank.feats[sampler.ids]).to(dev), text=torch.from_numpy(bank.text[cats]).to(dev),
                     ids=sampler.ids, categories=cats)
    loader = EpisodeLoader(fb, sampler, tasks)
    torch.manual_seed(123)
    model = utils.init_model(args, {})
    opt = utils.init_optim(args, model)
    eng = model._get_engine(dev)
    eng.precision = 1
    sampler.new_iterator()
    b = loader.next_batch().to(dev)
    for _ in range(2):
        eng.fumi_batch(model, b, steps=5, step_size=0.01, train=True)
    torch.cuda.synchronize()
    L = _lib.lib()
    L.fumi_debug_phase_profile(1)
    eng.fumi_batch(model, b, steps=5, step_size=0.01, train=True)
    torch.cuda.synchronize()
    out = np.zeros(64, np.uint64)
    L.fumi_debug_read_phases(_lib.ptr(out))
    L.fumi_debug_phase_profile(0)
    for lo, hi, name in ((20, 34, "forward"), (0, 15, "backward")):
        tot = float(out[lo:hi].sum())
        print(f"== {name}: {tot / 1e6:.1f} Mcycles summed over CTAs")
        for i in range(lo, hi):
            if out[i]:
                print(f"  {i:2d} {NAMES.get(i, '?'):38s} {100 * float(out[i]) / tot:5.1f}%")


if __name__ == "__main__":
    main()
