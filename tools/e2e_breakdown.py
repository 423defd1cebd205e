"""Host-side time per step of the end-to-end path (diagnostics; needs a GPU): where the host spends its time between
the device sync of one step and the first big kernel of the next."""
import os, random, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from fumi_b200 import utils
from fumi_b200.data.bank import FeatureBank
from fumi_b200.data.loader import EpisodeLoader
from fumi_b200.data.synth import class_split, make_bank
from fumi_b200.sampler import EpisodeSampler

dev = torch.device("cuda", 0)
tasks = 4096
args = bench.make_args("fumi", 5, 5, 32, 5, True, dev, 2048, 768, tasks, 0.25)
bank = make_bank(num_images=195605, num_classes=673)
cats = class_split(673)[0]
sampler = EpisodeSampler(bank.cat_of, cats, 5, 5, 32)
fb = FeatureBank(feats=torch.from_numpy(bank.feats[sampler.ids]).to(dev), text=torch.from_numpy(bank.text[cats]).to(dev),
                 ids=sampler.ids, categories=cats)
torch.manual_seed(123); random.seed(123)
model = utils.init_model(args, {}); opt = utils.init_optim(args, model)
eng = model._get_engine(dev)
it = iter(EpisodeLoader(fb, sampler, tasks, prefetch=2))
for _ in range(3):
    model.evaluate(args, next(it), opt, task="train")
# host time of every C-ABI call inside fumi_batch
calls = {}
orig_call = eng._call
def timed_call(name, fn, *a):
    t = time.perf_counter()
    r = orig_call(name, fn, *a)
    calls[name] = calls.get(name, 0.0) + time.perf_counter() - t
    return r
eng._call = timed_call
T = {k: 0.0 for k in ("next", "zero", "batch", "adam", "sync")}
n = 10
for _ in range(n):
    t0 = time.perf_counter(); b = next(it)
    t1 = time.perf_counter(); model.train(); opt.zero_grad()
    t2 = time.perf_counter(); res = eng.fumi_batch(model, b, steps=5, step_size=args.step_size, train=True)
    t3 = time.perf_counter(); opt.step()
    t4 = time.perf_counter(); la = res["loss_acc"].cpu().numpy()
    t5 = time.perf_counter()
    del res
    for k, v in zip(T, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4)):
        T[k] += v
print({k: round(v / n * 1e3, 3) for k, v in T.items()}, "ms per step (host)")
print({k: round(v / n * 1e3, 3) for k, v in sorted(calls.items(), key=lambda kv: -kv[1])}, "ms per step inside the C calls")
