"""Time fumi_gram at the bench shape (diagnostics; FUMI_GRAM_DBG / FUMI_GRAM_TC select variants)."""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fumi_b200 import _lib
L = _lib.lib(); dev = torch.device("cuda:0"); torch.manual_seed(0)
R, D, B, NK, NQ = 116127, 2048, 4096, 25, 160
feats = torch.relu(torch.randn(R, D, device=dev) + 0.5)
sup = torch.randint(0, R, (B, NK), device=dev); qry = torch.randint(0, R, (B, NQ), device=dev)
out = torch.empty(B, NK + NQ, NK, device=dev)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
if os.environ.get("FUMI_GRAM_F16") == "1":
    am = torch.empty(1, device=dev); hi = torch.empty(R, D, dtype=torch.float16, device=dev); lo = torch.empty_like(hi)
    _lib.check(L.fumi_absmax(_lib.ptr(feats), feats.numel(), _lib.ptr(am), st), "absmax")
    _lib.check(L.fumi_split_f16(_lib.ptr(feats), _lib.ptr(am), _lib.ptr(hi), _lib.ptr(lo), feats.numel(), st), "split")
    call = lambda: _lib.check(L.fumi_gram_f16(_lib.ptr(hi), _lib.ptr(lo), _lib.ptr(am), R, D, _lib.ptr(sup), _lib.ptr(qry), B, NK, NQ, _lib.ptr(out), st), "gram16")
else:
    call = lambda: _lib.check(L.fumi_gram(_lib.ptr(feats), R, D, _lib.ptr(sup), _lib.ptr(qry), B, NK, NQ, _lib.ptr(out), st), "gram")
call(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): call()
e1.record(); torch.cuda.synchronize()
print(os.environ.get("FUMI_GRAM_DBG", "0"), f"{e0.elapsed_time(e1) / 5:.3f} ms")
