"""Accuracy / time of the tcgen05 3xTF32 GEMM vs the number of k stages per TMEM drain (FUMI_GEMM_DRAIN)."""
import os, sys, subprocess
if len(sys.argv) == 1:
    for d in (1, 2, 4):
        subprocess.run([sys.executable, __file__, str(d)], env=dict(os.environ, FUMI_GEMM_DRAIN=str(d)))
    sys.exit(0)
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fumi_b200.engine import EpisodeEngine
eng = EpisodeEngine("cuda:0", precision=1)
rs = np.random.RandomState(0)
M, N, K = 116127, 256, 2048
a = torch.from_numpy(np.maximum(rs.randn(4096, K), 0).astype(np.float32)).cuda()
b = torch.from_numpy((rs.randn(N, K) / np.sqrt(K)).astype(np.float32)).cuda()
want = a.double() @ b.double().T
got = eng.gemm_tc(eng.split_tf32(a), eng.split_tf32(b))
f32 = (a @ b.T).double()
rel = lambda x: float((x.double() - want).abs().max() / want.abs().max())
A = torch.randn(M, K, device="cuda"); ap = eng.split_tf32(A); bp = eng.split_tf32(b)
out = torch.empty(M, N, device="cuda")
for _ in range(3): eng.gemm_tc(ap, bp, out=out)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): eng.gemm_tc(ap, bp, out=out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"drain={sys.argv[1]}: err {rel(got):.2e} (fp32 matmul {rel(f32):.2e})  proj GEMM {ms:.3f} ms = {3*2*M*N*K/ms/1e9:.0f} TFLOP/s tf32-equivalent, {2*M*N*K/ms/1e9:.0f} useful")
