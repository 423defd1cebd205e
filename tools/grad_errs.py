"""Print the tensor-normalised relative error of every meta-gradient of a golden train case (GPU)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import kernel_cases as kc
from helpers import load_golden, params_of, relerr
from fumi_b200.optim import FusedAdam

name = sys.argv[1] if len(sys.argv) > 1 else "fumi_train_n5k5_d512_tanh"
dev = "cuda:0"
g, bank = load_golden(name)
model = kc.make_fumi(g, bank, params_of(g), dev)
model._get_engine(dev).precision = 2
opt = FusedAdam(model.parameters(), lr=float(g["lr"]), weight_decay=float(g["wd"]))
eng = model._get_engine(dev)
res = eng.fumi_batch(model, kc._torchmeta_batch(g, bank), steps=int(g["steps"]), step_size=float(g["alpha"]), train=True)
print("loss", res["loss_acc"].cpu().numpy(), float(g["loss"]), float(g["acc"]))
print("logits relerr", relerr(res["logits"].cpu().numpy(), g["logits"]))
for k, p in model.named_parameters():
    print(f"{k:28s} {relerr(p.grad.cpu().numpy(), g['grad:' + k]):.3e}  max|ref| {np.abs(g['grad:' + k]).max():.3e}")
