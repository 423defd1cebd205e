import os, random, sys, time
import numpy as np, torch
sys.path.insert(0, "/root/repo")
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
fb = FeatureBank(feats=torch.from_numpy(bank.feats[sampler.ids]).to(dev), text=torch.from_numpy(bank.text[cats]).to(dev), ids=sampler.ids, categories=cats)
torch.manual_seed(123); random.seed(123)
model = utils.init_model(args, {}); opt = utils.init_optim(args, model)
import gc
for mode in ("1",) * 10:
    os.environ["FUMI_EARLY_LOSS"] = mode
    ld = EpisodeLoader(fb, sampler, tasks, prefetch=2)
    it = iter(ld)
    for _ in range(3): model.evaluate(args, next(it), opt, task="train")
    torch.cuda.synchronize()
    tn = te = 0.0
    per = []
    ms0 = torch.cuda.memory_stats()
    t0 = time.perf_counter()
    for _ in range(40):
        a = time.perf_counter(); b = next(it); c = time.perf_counter(); model.evaluate(args, b, opt, task="train"); d = time.perf_counter()
        tn += c - a; te += d - c; per.append((round((c - a) * 1e3, 1), round((d - c) * 1e3, 1)))
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    ms1 = torch.cuda.memory_stats()
    print({k: ms1[k] - ms0[k] for k in ('num_device_alloc', 'num_device_free', 'num_alloc_retries', 'num_sync_all_streams')}, ms1['reserved_bytes.all.current'] >> 20)
    print([p for p in per if p[0] + p[1] > 12], gc.get_count(), gc.get_stats()[2])
    print("early" if mode == "1" else "late ", "ms/step", round((t1 - t0) / 40 * 1e3, 3), "next", round(tn / 40 * 1e3, 3), "evaluate", round(te / 40 * 1e3, 3))
    ld.close()
