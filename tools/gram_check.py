import os, sys, time, torch, numpy as np
sys.path.insert(0, "/root/repo")
from fumi_b200 import _lib
import ctypes as C
L = _lib.lib()
dev = torch.device("cuda:0")
torch.manual_seed(0)
def run(R, D, B, NK, NQ, reps=5):
    feats = torch.relu(torch.randn(R, D, device=dev) + 0.5)
    sup = torch.randint(0, R, (B, NK), device=dev)
    qry = torch.randint(0, R, (B, NQ), device=dev)
    out = torch.full((B, NK + NQ, NK), float("nan"), device=dev)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    def call():
        _lib.check(L.fumi_gram(_lib.ptr(feats), R, D, _lib.ptr(sup), _lib.ptr(qry), B, NK, NQ, _lib.ptr(out), st), "gram")
    call(); torch.cuda.synchronize()
    nb = min(B, 64)
    X = feats[torch.cat([sup[:nb], qry[:nb]], 1)].double()          # [nb, NK+NQ, D]
    ref = X @ X[:, :NK].transpose(1, 2)
    err = ((out[:nb].double() - ref).abs().max() / ref.abs().max()).item()
    # last tasks too
    X2 = feats[torch.cat([sup[-8:], qry[-8:]], 1)].double()
    ref2 = X2 @ X2[:, :NK].transpose(1, 2)
    err2 = ((out[-8:].double() - ref2).abs().max() / ref2.abs().max()).item()
    ref32 = (feats[torch.cat([sup[:nb], qry[:nb]], 1)] @ feats[sup[:nb]].transpose(1, 2)).double()
    e32 = ((ref32 - ref).abs().max() / ref.abs().max()).item()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): call()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    gb = B * (NK + NQ) * D * 4 / 1e9
    print(f"R={R} D={D} B={B} NK={NK} NQ={NQ}: relerr {err:.2e} / {err2:.2e} (torch fp32 {e32:.2e})  {ms:.3f} ms  {gb/ms*1e3:.0f} GB/s  nan={torch.isnan(out).any().item()}")
run(5000, 2048, 300, 25, 160)
run(5000, 512, 37, 5, 100)
run(5000, 2048, 149, 25, 100)
run(116127, 2048, 4096, 25, 160)
run(116127, 2048, 4096, 5, 100)
