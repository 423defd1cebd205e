"""Parity cases shared by the host-emulation tests (CPU, here) and the GPU tests (B200 box).

Each function takes a torch device and drives the product's public host API (fumi_b200.fumi /
maml / am3 / engine -> C ABI), then checks against the golden vectors produced by the reference
and against the CPU oracle.  Tolerances: integers / predictions bit-exact; fp32 within 1e-4
relative (north_star), tensor-normalised: max|a-b| / max|b|.
"""
import types

import numpy as np
import torch

from fumi_b200 import engine as engine_mod
from fumi_b200 import fumi as fumi_mod
from fumi_b200 import maml as maml_mod
from fumi_b200 import am3 as am3_mod
from fumi_b200.data.bank import EpisodeBatch, FeatureBank
from fumi_b200.optim import FusedAdam
from oracle import episode_np

from helpers import argv_int, flat_batch, load_golden, params_of, relerr

TOL = 1e-4


def _args(g, device, **kw):
    a = types.SimpleNamespace(device=torch.device(device), step_size=float(g["alpha"]),
                              num_train_adapt_steps=int(g["steps"]), num_test_adapt_steps=int(g["steps"]),
                              first_order=bool(g["first_order"]) if "first_order" in g else False)
    a.__dict__.update(kw)
    return a


def _torchmeta_batch(g, bank):
    """The reference's batch dict rebuilt from golden ids: [[ids, text, im], targets]."""
    t = torch.from_numpy
    sup_text = bank.text[bank.cat_of[g["sup_ids"]]]
    qry_text = bank.text[bank.cat_of[g["qry_ids"]]]
    return {"train": [[t(g["sup_ids"]), t(sup_text), t(bank.feats[g["sup_ids"]])], t(g["sup_y"])],
            "test": [[t(g["qry_ids"]), t(qry_text), t(bank.feats[g["qry_ids"]])], t(g["qry_y"])]}


def _bank_batch(g, bank, device, N):
    """Index form: whole synthetic bank resident on the device, rows = image ids."""
    fb = FeatureBank(feats=torch.from_numpy(bank.feats).to(device), text=torch.from_numpy(bank.text).to(device),
                     ids=np.arange(bank.feats.shape[0]), categories=np.arange(bank.text.shape[0]))
    from helpers import class_text_rows
    cats = class_text_rows(bank, g["sup_ids"], g["sup_y"], N)
    return EpisodeBatch(bank=fb, sup_rows=torch.from_numpy(g["sup_ids"]), qry_rows=torch.from_numpy(g["qry_ids"]),
                        sup_y=torch.from_numpy(g["sup_y"]), qry_y=torch.from_numpy(g["qry_y"]),
                        sup_ids=g["sup_ids"], qry_ids=g["qry_ids"], head_class=torch.from_numpy(cats))


def _load(model, params, device):
    model.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()})
    return model.to(device)


def make_fumi(g, bank, params, device, dropout=0.0):
    argv = str(g["argv"])
    m = fumi_mod.FUMI(n_way=argv_int(g, "--num_ways", 5), im_emb_dim=bank.feats.shape[1], im_hid_dim=[256, 64],
                      text_encoder="BERT", text_emb_dim=bank.text.shape[1], text_hid_dim=256, dropout_rate=dropout,
                      norm_hypernet="--norm_hypernet" in argv)
    return _load(m, params, device)


def step_gates(rec, lay, NK):
    """(H0 > 0 [NK,256], H1 > 0 [NK,64]) of one step record of the stash (fumi_stash_layout): fp32 activations
    (format 0) or, on the tensor-core path, H0 as the forward's fp16 hi/lo operand planes (format 1)."""
    h1 = rec[lay.rec_h1:lay.rec_h1 + NK * 64].reshape(NK, 64)
    if lay.format == 0:
        h0 = rec[lay.rec_h0_hi:lay.rec_h0_hi + NK * 256].reshape(NK, 256)
    else:
        hi = rec[lay.rec_h0_hi:lay.rec_h0_hi + NK * 128].view(np.float16).reshape(NK, 256).astype(np.float32)
        lo = rec[lay.rec_h0_lo:lay.rec_h0_lo + NK * 128].view(np.float16).reshape(NK, 256).astype(np.float32)
        h0 = hi + lo                       # times 2^-e (e = rec.view(int32)[lay.rec_exp]): irrelevant for the sign
    return h0 > 0, h1 > 0


def check_adapted(res, g, eng, bank, N):
    """Adapted head / W1 / biases and the materialised W0 rows against the reference's."""
    lay = eng.stash_layout(res["cfg"])
    B = g["sup_ids"].shape[0]
    st = res["stash"].cpu().numpy().reshape(B, lay.per_task)
    NK = g["sup_ids"].shape[1]
    for b in range(B):
        head = st[b, lay.head:lay.head + N * 65].reshape(N, 65)
        w1 = st[b, lay.w1t:lay.w1t + 256 * 64].reshape(256, 64).T
        assert relerr(head, g["hp_adapted"][b]) < TOL
        assert relerr(w1, g["W1_adapted"][b]) < TOL
        assert relerr(st[b, lay.b1:lay.b1 + 64], g["b1_adapted"][b]) < TOL
        assert relerr(st[b, lay.b0:lay.b0 + 256], g["b0_adapted"][b]) < TOL
    # W0_task = W0 - alpha * S^T X   (Gram form <-> the reference's materialised linear0.weight)
    S = st[0, lay.S:lay.S + NK * 256].reshape(NK, 256)
    X = bank.feats[g["sup_ids"][0]]
    return S, X


def gemm_tc_case(device):
    """tcgen05 3xTF32 GEMM: fp32-level accuracy against fp64, ragged M / N / K, split-K, fused epilogue."""
    eng = engine_mod.EpisodeEngine(device, precision=1)
    rs = np.random.RandomState(8)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
    for (M, N, K, split) in [(128, 256, 32, 1), (128, 256, 64, 1), (403, 256, 768, 1), (1000, 64, 512, 1),
                             (300, 65, 100, 1), (256, 2048, 5000, 0), (256, 512, 4100, 7), (37, 300, 36, 1)]:
        a = (rs.randn(M, K) * rs.uniform(0.1, 3, size=(M, 1))).astype(np.float32)
        b = (rs.randn(N, K) / np.sqrt(K)).astype(np.float32)
        bias = rs.randn(N).astype(np.float32)
        want = a.astype(np.float64) @ b.astype(np.float64).T
        ap, bp = eng.split_tf32(t(a)), eng.split_tf32(t(b))
        assert np.array_equal((ap[0] + ap[1]).cpu().numpy(), a), "hi + lo must reconstruct x exactly"
        got = eng.gemm_tc(ap, bp, split_k=split).cpu().numpy()
        err = relerr(got, want)
        f32 = relerr(a @ b.T, want)
        assert err < max(4 * f32, 2e-6), (M, N, K, split, err, f32)
        if split == 1:
            got = eng.gemm_tc(ap, bp, bias=t(bias), act=1).cpu().numpy()
            assert relerr(got, np.maximum(want + bias, 0)) < max(4 * f32, 2e-6), (M, N, K, "bias+relu")
        acc = eng.gemm_tc(ap, bp, split_k=split, out=torch.ones(M, N, device=device), accumulate=True).cpu().numpy()
        assert relerr(acc, want + 1) < max(4 * f32, 2e-6), (M, N, K, "accumulate")
    # weight-gradient form: C = X1^T X2 through transposed planes
    R, C1, C2 = 1234, 256, 512
    x1, x2 = rs.randn(R, C1).astype(np.float32), rs.randn(R, C2).astype(np.float32)
    got = eng.gemm_tc(eng.transpose_split_tf32(t(x1)), eng.transpose_split_tf32(t(x2)), K=R).cpu().numpy()
    assert relerr(got, x1.astype(np.float64).T @ x2.astype(np.float64)) < 2e-6


def gemm_f16_case(device):
    """tcgen05 3 x fp16 GEMM: fp32-level accuracy against fp64 on operands of very different magnitudes (the
    power-of-two plane scales), ragged shapes, split-K, fused epilogue, transposed (weight-gradient) form."""
    eng = engine_mod.EpisodeEngine(device, precision=2)
    rs = np.random.RandomState(9)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
    for (M, N, K, split, sa, sb) in [(128, 256, 64, 1, 1.0, 1.0), (403, 256, 768, 1, 3.0, 0.02), (1000, 64, 512, 1, 1e-6, 1e3),
                                      (300, 65, 104, 1, 1.0, 1.0), (256, 2048, 5000, 0, 1e-5, 2.0), (256, 512, 4104, 7, 40.0, 1e-3),
                                      (37, 300, 40, 1, 1.0, 1.0)]:
        a = (rs.randn(M, K) * rs.uniform(0.1, 3, size=(M, 1)) * sa).astype(np.float32)
        b = (rs.randn(N, K) / np.sqrt(K) * sb).astype(np.float32)
        bias = (rs.randn(N) * sa * sb).astype(np.float32)
        want = a.astype(np.float64) @ b.astype(np.float64).T
        ap, bp = eng.split_f16(t(a)), eng.split_f16(t(b))
        assert abs(float(ap[2].item()) - np.abs(a).max()) == 0.0
        got = eng.gemm_f16(ap, bp, split_k=split).cpu().numpy()
        err = relerr(got, want)
        f32 = relerr(a @ b.T, want)
        assert err < max(4 * f32, 2e-6), (M, N, K, split, err, f32)
        if split == 1:
            got = eng.gemm_f16(ap, bp, bias=t(bias), act=1).cpu().numpy()
            assert relerr(got, np.maximum(want + bias, 0)) < max(4 * f32, 2e-6), (M, N, K, "bias+relu")
        acc = eng.gemm_f16(ap, bp, split_k=split, out=torch.ones(M, N, device=device), accumulate=True).cpu().numpy()
        assert relerr(acc - 1, want) < max(4 * f32, 1e-5) or relerr(acc, want + 1) < 2e-6, (M, N, K, "accumulate")
    R, C1, C2 = 1234, 256, 512
    x1, x2 = (rs.randn(R, C1) * 1e-4).astype(np.float32), np.maximum(rs.randn(R, C2), 0).astype(np.float32)
    x1[rs.rand(R) < 0.5] = 0.0                                    # like d_proj: many untouched rows
    got = eng.gemm_f16(eng.transpose_split_f16(t(x1)), eng.transpose_split_f16(t(x2)), K=R).cpu().numpy()
    assert relerr(got, x1.astype(np.float64).T @ x2.astype(np.float64)) < 2e-6
    # more M tiles than SMs (a CTA walks several tiles: ring / TMEM ping-pong run on across them; odd tile count -> idle
    # partner in the last multicast pair), and split-K summed in slice order: two runs are bit-identical
    a = rs.randn(148 * 128 * 2 + 77, 192).astype(np.float32)
    b = (rs.randn(256, 192) / np.sqrt(192)).astype(np.float32)
    got = eng.gemm_f16(eng.split_f16(t(a)), eng.split_f16(t(b))).cpu().numpy()
    assert relerr(got, a.astype(np.float64) @ b.astype(np.float64).T) < 2e-6
    xa, xb = eng.transpose_split_f16(t(x1)), eng.transpose_split_f16(t(x2))
    r1 = eng.gemm_f16(xa, xb, K=R, split_k=5).cpu().numpy()
    r2 = eng.gemm_f16(xa, xb, K=R, split_k=5).cpu().numpy()
    assert np.array_equal(r1, r2), "split-K must be reproducible run to run"
    z = eng.split_f16(torch.zeros(8, 64, device=device))          # all-zero operand: scale 1, no NaN
    assert float(eng.gemm_f16(z, eng.split_f16(t(rs.randn(16, 64).astype(np.float32)))).abs().max()) == 0.0


def warp_gemm_f16_case(device):
    """The fp16 hi/lo plane warp GEMM (ldmatrix addressing, fragment order, plane scales) in every layout variant."""
    from fumi_b200 import _lib
    rs = np.random.RandomState(12)
    if torch.device(device).type == "cuda":
        import build_test_kernels            # test-only sm_100a kernels (tests/csrc), not part of libfumi_b200.so
        L = build_test_kernels.load()
    else:
        import ctypes
        import build_emu                     # the emulation build carries the same test kernel
        L = ctypes.CDLL(build_emu.build())
        L.fumi_debug_gemm_f16.restype = ctypes.c_int
        L.fumi_debug_gemm_f16.argtypes = [ctypes.c_void_p] * 2 + [ctypes.c_int32] * 4 + [ctypes.c_void_p] * 2
    stream = _lib.stream_ptr(torch.device(device)) if torch.device(device).type == "cuda" else None
    for (M, N) in [(16, 8), (32, 16), (16, 64), (32, 8)]:
        for K in (16, 48, 64):
            for variant in range(4):
                a = (rs.randn(M, K) * rs.uniform(1e-3, 30)).astype(np.float32)
                b = (rs.randn(K, N) * rs.uniform(1e-4, 5)).astype(np.float32)
                out = torch.full((M, N), float("nan"), device=device)
                ta, tb = torch.from_numpy(a).to(device), torch.from_numpy(b).to(device)
                _lib.check(L.fumi_debug_gemm_f16(_lib.ptr(ta), _lib.ptr(tb), variant, M, N, K, _lib.ptr(out), stream), "dbg")
                want = a.astype(np.float64) @ b.astype(np.float64)
                assert relerr(out.cpu().numpy(), want) < 2e-6, (M, N, K, variant, relerr(out.cpu().numpy(), want))


def fumi_train_case(device, name, via="dict", precision=0):
    g, bank = load_golden(name)
    N = argv_int(g, "--num_ways", 5)
    params = params_of(g)
    model = make_fumi(g, bank, params, device)
    model._get_engine(device).precision = precision
    opt = FusedAdam(model.parameters(), lr=float(g["lr"]), weight_decay=float(g["wd"]))
    args = _args(g, device)
    batch = _torchmeta_batch(g, bank) if via == "dict" else _bank_batch(g, bank, device, N)
    eng = model._get_engine(device)
    # forward only first: the adapted state (the backward reuses the S slots for its adjoint)
    fw = eng.fumi_batch(model, batch, steps=int(g["steps"]), step_size=float(g["alpha"]), train=False,
                        return_state=True)
    S, X = check_adapted(fw, g, eng, bank, N)
    W0_task = params["im_net.linear0.weight"][:16] - float(g["alpha"]) * (S.T @ X)[:16]
    assert relerr(W0_task, g["W0_adapted_t0_rows16"]) < TOL
    res = eng.fumi_batch(model, batch, steps=int(g["steps"]), step_size=float(g["alpha"]), train=True)
    assert np.array_equal(res["preds"].cpu().numpy(), g["preds"]), "query predictions must be bit-exact"
    assert relerr(res["logits"].cpu().numpy(), g["logits"]) < TOL
    la = res["loss_acc"].cpu().numpy()
    assert abs(la[0] - float(g["loss"])) < TOL * abs(float(g["loss"]))
    assert abs(la[1] - float(g["acc"])) < 1e-6
    gmax = max(np.abs(g["grad:" + k]).max() for k in params)
    for k, p in model.named_parameters():
        got, ref = p.grad.cpu().numpy(), g["grad:" + k]
        if k == "hyper_net.2.bias":       # true gradient is exactly 0: absolute check
            assert np.abs(got - ref).max() < 1e-4 * gmax
        else:
            assert relerr(got, ref) < 2e-4, (k, relerr(got, ref))
    # Adam normalises noise-level gradient entries (|g| ~ eps) into +-lr steps, so the optimizer kernel is
    # checked on the reference's own gradient; the gradient itself was checked just above.
    for k, p in model.named_parameters():
        p.grad.copy_(torch.from_numpy(g["grad:" + k]).to(device))
    opt.step()
    for k, p in model.named_parameters():
        ref = g["post:" + k]
        assert np.abs(p.detach().cpu().numpy() - ref).max() <= 1e-7 + 2e-6 * np.abs(ref).max(), k
    return res


def fumi_evaluate_api_case(device, name):
    """Through FUMI.evaluate exactly as the reference loop calls it (fumi.py:242,316)."""
    g, bank = load_golden(name)
    model = make_fumi(g, bank, params_of(g), device)
    opt = FusedAdam(model.parameters(), lr=float(g["lr"]), weight_decay=float(g["wd"]))
    loss, acc, preds, targets = model.evaluate(_args(g, device), _torchmeta_batch(g, bank), opt, task="train")
    assert loss.dtype == np.float32 and loss.shape == () and acc.shape == ()
    assert preds.dtype == torch.float32 and preds.shape == tuple(g["preds"].shape)       # fumi.py:140,180
    assert np.array_equal(preds.cpu().numpy().astype(np.int64), g["preds"])
    assert np.array_equal(targets.cpu().numpy(), g["qry_y"])
    assert abs(float(loss) - float(g["loss"])) < TOL * abs(float(g["loss"]))
    for k, p in model.named_parameters():          # one optimizer step was taken (fumi.py:193)
        ref, pre = g["post:" + k], g["param:" + k]
        big = np.abs(g["grad:" + k] + float(g["wd"]) * pre) > 1e-5
        assert np.abs(p.detach().cpu().numpy() - ref)[big].max() <= 1e-6, k
        assert np.abs(p.detach().cpu().numpy() - pre).max() <= 1.01 * float(g["lr"]) + 1e-7


def fumi_test_case(device, name, via="dict"):
    g, bank = load_golden(name)
    gp, _ = load_golden("fumi_test_n5k1_full")
    params = params_of(gp)
    model = make_fumi(g, bank, params, device, dropout=0.25)          # eval mode: dropout layers inert
    batch = _torchmeta_batch(g, bank) if via == "dict" else _bank_batch(g, bank, device, 5)
    eng = model._get_engine(device)
    model.eval()
    res = eng.fumi_batch(model, batch, steps=int(g["steps"]), step_size=float(g["alpha"]), train=False,
                         return_state=True)
    assert np.array_equal(res["preds"].cpu().numpy(), g["preds"])
    la = res["loss_acc"].cpu().numpy()
    assert abs(la[1] - float(g["acc"])) < 1e-6
    # A 100-step unroll is discontinuous in the ReLU gates: a pre-activation within rounding distance of 0
    # (|z| ~ 1e-6) can be gated either way by two correct fp32 implementations, after which the two
    # trajectories differ by a finite amount.  So: (1) the kernel must equal the fp64 oracle run with the
    # kernel's own gates to 1e-4, and every gate that differs from the oracle's must be such a tie;
    # (2) where no gate differs, the reference's golden logits must be met to 1e-4 as well.
    steps, NK = int(g["steps"]), g["sup_ids"].shape[1]
    lay = eng.stash_layout(res["cfg"])
    B = g["sup_ids"].shape[0]
    st = res["stash"].cpu().numpy().reshape(B, lay.per_task)
    gates = []
    for b in range(B):
        recs = [st[b, lay.steps + s * lay.per_step:lay.steps + (s + 1) * lay.per_step] for s in range(steps)]
        gates.append([step_gates(r, lay, NK) for r in recs])
    fb = flat_batch(g, bank, 5)
    o64 = episode_np.fumi_batch(params, fb, float(g["alpha"]), steps, dtype=np.float64, relu_gates=gates)
    logits = res["logits"].cpu().numpy()
    assert relerr(logits, o64["logits"]) < TOL
    n_ties = 0
    for b, t in enumerate(o64["tasks"]):
        assert all(z < 1e-5 for z in t["relu_ties"]), ("gate differs on a non-tie", b, max(t["relu_ties"]))
        n_ties += len(t["relu_ties"])
        lay_head = st[b, lay.head:lay.head + 5 * 65].reshape(5, 65)
        assert relerr(lay_head, t["adapted"][4]) < TOL
        assert relerr(st[b, lay.w1t:lay.w1t + 256 * 64].reshape(256, 64).T, t["adapted"][2]) < TOL
        if not t["relu_ties"]:
            assert relerr(logits[b], g["logits"][b]) < TOL
            assert relerr(lay_head, g["hp_adapted"][b]) < TOL
    if n_ties == 0:
        assert abs(la[0] - float(g["loss"])) < TOL * abs(float(g["loss"]))
        check_adapted(res, g, eng, bank, 5)
    else:
        assert abs(la[0] - float(g["loss"])) < 1e-2 * abs(float(g["loss"]))
    return n_ties


def maml_case(device, name):
    g, bank = load_golden(name)
    params = params_of(g)
    model = _load(maml_mod.PureImageNetwork(im_embed_dim=bank.feats.shape[1], n_way=5, hidden_dims=[256, 64]), params,
                  device)
    train = "train" in name
    opt = FusedAdam(model.parameters(), lr=float(g["lr"]), weight_decay=float(g["wd"])) if train else None
    loss, acc = maml_mod.evaluate(_args(g, device), model, _torchmeta_batch(g, bank), opt,
                                  task="train" if train else "test")
    res = maml_mod.evaluate.last
    assert np.array_equal(res["preds"].cpu().numpy(), g["preds"])
    assert relerr(res["logits"].cpu().numpy(), g["logits"]) < TOL
    assert abs(float(loss) - float(g["loss"])) < TOL * abs(float(g["loss"]))
    assert abs(float(acc) - float(g["acc"])) < 1e-6
    if train:
        for k, p in model.named_parameters():
            assert relerr(p.grad.cpu().numpy(), g["grad:" + k]) < 2e-4, k
            # parameters moved by about lr in the direction of -grad (exact Adam check: fumi_train_case)
            ref, pre = g["post:" + k], g["param:" + k]
            assert np.abs(p.detach().cpu().numpy() - pre).max() <= 1.01 * float(g["lr"]) + 1e-7
            big = np.abs(g["grad:" + k] + float(g["wd"]) * pre) > 1e-5
            assert np.abs(p.detach().cpu().numpy() - ref)[big].max() <= 1e-6, k


def maml_test_then_train_case(device, name="maml_train_n5k5_d512"):
    """A test-mode call with optimizer=None (test_loop, which training_run runs first and every eval_freq) must not
    detach the gradients from FusedAdam's flat buffer: the next train step still fills the flat buffer (the one the
    multi-GPU all-reduce sums) and steps with a single fused launch."""
    g, bank = load_golden(name)
    model = _load(maml_mod.PureImageNetwork(im_embed_dim=bank.feats.shape[1], n_way=5, hidden_dims=[256, 64]),
                  params_of(g), device)
    opt = FusedAdam(model.parameters(), lr=float(g["lr"]), weight_decay=float(g["wd"]))
    args, batch = _args(g, device), _torchmeta_batch(g, bank)
    maml_mod.evaluate(args, model, batch, None, task="test")
    f = opt._flat[0]
    assert opt._is_flat(f), "test-mode evaluate detached the gradients from the flat buffer"
    maml_mod.evaluate(args, model, batch, opt, task="train")
    assert opt._is_flat(f)
    eng = model._get_engine(device)
    assert eng._flat_grads(list(model.parameters())) is f["g_ext"]
    off = 0
    for k, p in model.named_parameters():                       # the flat buffer holds the real gradients
        got = f["g"][off:off + p.numel()].view(p.shape).cpu().numpy()
        assert relerr(got, g["grad:" + k]) < 2e-4, k
        off += p.numel()
    # a foreign zero_grad (Module.zero_grad sets p.grad = None) is repaired by the next FusedAdam.zero_grad
    model.zero_grad()
    assert not opt._is_flat(f)
    opt.zero_grad()
    assert opt._is_flat(f)


def oracle_train_case(device, N, K, Q, steps, D=2048, T=768, B=8, dropout=0.25, precision=2, tanh=False, seed=5):
    """Meta-train step at arbitrary (benched) dimensions against oracle/episode_np in fp64, fed the same counter-based
    dropout masks: predictions equal, logits within 1e-4, all meta-gradients within 2e-4.  No fixture: the oracle
    is evaluated here (dW0 alone is 2 MB at D = 2048).  This is bench.py's parity probe run as a test."""
    import random
    import bench
    from fumi_b200 import utils
    from fumi_b200.data.bank import FeatureBank
    from fumi_b200.data.loader import EpisodeLoader
    from fumi_b200.data.synth import class_split, make_bank
    from fumi_b200.sampler import EpisodeSampler
    C = max(60, 3 * N)
    bank = make_bank(num_images=C * (K + Q + 30), num_classes=C, im_dim=D, text_dim=T, min_per_class=K + Q + 8, seed=seed)
    cats = class_split(C)[0]
    sampler = EpisodeSampler(bank.cat_of, cats, N, K, Q)
    fb = FeatureBank(feats=torch.from_numpy(bank.feats[sampler.ids]).to(device), text=torch.from_numpy(bank.text[cats]).to(device),
                     ids=sampler.ids, categories=cats)
    argv = ["--model", "fumi", "--num_ways", str(N), "--num_shots", str(K), "--num_shots_test", str(Q),
            "--num_train_adapt_steps", str(steps), "--im_emb_dim", str(D), "--text_emb_dim", str(T),
            "--dropout", str(dropout), "--batch_size", str(B)] + (["--norm_hypernet"] if tanh else [])
    args = utils.parser().parse_args(argv)
    args.device = torch.device(device)
    args.precision = precision
    torch.manual_seed(123); np.random.seed(123); random.seed(123)
    model = utils.init_model(args, {})
    FusedAdam(model.parameters(), lr=3e-5, weight_decay=5e-4)
    eng = model._get_engine(device)
    loader = EpisodeLoader(fb, sampler, B)
    sampler.new_iterator()
    batch = loader.next_batch().to(device)
    par = bench.parity_probe(None, eng, model, args, batch, N, K, Q, steps, dropout, ntasks=B)
    assert par["preds_equal"], par
    assert par["logits_relerr"] < 1e-4, par
    assert par["grad_relerr_max"] < 2e-4, par
    return par


def am3_case(device, name="am3_test_n10k5_d512"):
    g, bank = load_golden(name)
    m = am3_mod.AM3(im_encoder="precomputed", im_emb_dim=bank.feats.shape[1], text_encoder="BERT",
                    text_emb_dim=bank.text.shape[1], text_hid_dim=256, prototype_dim=64, dropout=0.25)
    model = _load(m, params_of(g), device)
    out = model.evaluate(batch=_torchmeta_batch(g, bank), optimizer=None, scheduler=None, num_ways=10, device=device,
                         task="test")
    loss, acc, f1, prec, rec, lam, preds, trues, qidx, sidx, slam = out
    assert np.array_equal(preds, g["preds"])
    assert abs(float(loss) - float(g["loss"])) < TOL * abs(float(g["loss"]))
    for a, k in ((acc, "acc"), (f1, "f1"), (prec, "prec"), (rec, "rec")):
        assert abs(float(a) - float(g[k])) < 1e-9, k
    assert abs(float(lam) - float(g["avg_lamda"])) < 1e-6
    assert relerr(slam, g["sup_lamda"]) < TOL
    assert np.array_equal(qidx, g["qry_ids"]) and np.array_equal(sidx, g["sup_ids"])


def am3_train_case(device, name="am3_train_n10k5_d512"):
    """One AM3 meta-train step through AM3.evaluate(task='train') (am3.py:154-196) against the reference's gradients and
    post-Adam parameters (--dropout 0 golden)."""
    g, bank = load_golden(name)
    m = am3_mod.AM3(im_encoder="precomputed", im_emb_dim=bank.feats.shape[1], text_encoder="BERT",
                    text_emb_dim=bank.text.shape[1], text_hid_dim=256, prototype_dim=64, dropout=0.0)
    model = _load(m, params_of(g), device)
    opt = FusedAdam(model.parameters(), lr=float(g["lr"]), weight_decay=float(g["wd"]))
    out = model.evaluate(batch=_torchmeta_batch(g, bank), optimizer=opt, scheduler=None, num_ways=10, device=device,
                         task="train")
    loss, acc, f1, prec, rec, lam = out
    assert abs(float(loss) - float(g["loss"])) < TOL * abs(float(g["loss"]))
    for a, k in ((acc, "acc"), (f1, "f1"), (prec, "prec"), (rec, "rec")):
        assert abs(float(a) - float(g[k])) < 1e-9, k
    assert abs(float(lam) - float(g["avg_lamda"])) < 1e-6
    f = opt._flat[0]
    off = 0
    for k, p in model.named_parameters():
        if not p.requires_grad:
            continue
        got = f["g"][off:off + p.numel()].view(p.shape).cpu().numpy()
        assert relerr(got, g["grad:" + k]) < 2e-4, (k, relerr(got, g["grad:" + k]))
        off += p.numel()
        ref, pre = g["post:" + k], g["param:" + k]
        big = np.abs(g["grad:" + k] + float(g["wd"]) * pre) > 1e-5
        assert np.abs(p.detach().cpu().numpy() - ref)[big].max() <= 1e-6, k
        assert np.abs(p.detach().cpu().numpy() - pre).max() <= 1.01 * float(g["lr"]) + 1e-7


def dropout_case(device, name="fumi_train_n5k5_d512", p=0.25, seed=77):
    """Counter-based dropout masks: the device path must equal the oracle fed the same masks."""
    from fumi_b200.dropout import mask_array
    g, bank = load_golden(name)
    N = 5
    params = params_of(g)
    model = make_fumi(g, bank, params, device, dropout=p)
    model.train()
    model.dropout_base_seed = 0
    model.dropout_seed = seed - 1          # fumi_batch advances it by one per training batch
    eng = model._get_engine(device)
    steps, alpha = int(g["steps"]), float(g["alpha"])
    res = eng.fumi_batch(model, _torchmeta_batch(g, bank), steps=steps, step_size=alpha, train=True)
    fb = flat_batch(g, bank, N)
    B, NK = g["sup_ids"].shape
    NQ = g["qry_ids"].shape[1]
    masks = []
    for b in range(B):
        sup = [(mask_array(seed, b, s, 0, NK, 256, p), mask_array(seed, b, s, 1, NK, 64, p)) for s in range(steps)]
        qry = (mask_array(seed, b, steps, 0, NQ, 256, p), mask_array(seed, b, steps, 1, NQ, 64, p))
        masks.append(dict(sup=sup, qry=qry))
    ref = episode_np.fumi_batch(params, fb, alpha, steps, masks=masks, want_grad=True, dtype=np.float32)
    kept = np.mean([m["qry"][0] > 0 for m in masks])
    assert abs(kept - (1 - p)) < 0.02, "mask keep-rate is off"
    assert relerr(res["logits"].cpu().numpy(), ref["logits"]) < TOL
    assert np.array_equal(res["preds"].cpu().numpy(), ref["preds"])
    for k, q in model.named_parameters():
        if k != "hyper_net.2.bias":
            assert relerr(q.grad.cpu().numpy(), ref["grads"][k]) < 2e-4, k


def dense_case(device):
    eng = engine_mod.EpisodeEngine(device, precision=0)
    rs = np.random.RandomState(3)
    for (M, N, K) in [(37, 65, 24), (300, 256, 512), (5, 1, 256), (130, 129, 20)]:
        x, w, b = rs.randn(M, K).astype(np.float32), (rs.randn(N, K) / np.sqrt(K)).astype(np.float32), rs.randn(N).astype(np.float32)
        t = lambda a: torch.from_numpy(a).to(device)
        for act, f in ((0, lambda v: v), (1, lambda v: np.maximum(v, 0)), (2, np.tanh),
                       (3, lambda v: 1 / (1 + np.exp(-v)))):
            y = eng.linear_fwd(t(x), t(w), t(b), act=act, precision=0).cpu().numpy()
            want = f(x.astype(np.float64) @ w.T.astype(np.float64) + b)
            assert relerr(y, want) < 1e-5, (M, N, K, act)
        dy = rs.randn(M, N).astype(np.float32)
        dw, db = torch.empty(N, K, device=device), torch.empty(N, device=device)
        eng.linear_wgrad(t(dy), t(x), dw, db, precision=0)
        assert relerr(dw.cpu().numpy(), dy.astype(np.float64).T @ x) < 1e-5
        assert relerr(db.cpu().numpy(), dy.astype(np.float64).sum(0)) < 1e-5
        eng.linear_wgrad(t(dy), t(x), dw, None, accumulate=True, precision=0)
        assert relerr(dw.cpu().numpy(), 2 * (dy.astype(np.float64).T @ x)) < 1e-5
        gate = rs.randn(M, K).astype(np.float32)
        dx = eng.linear_dgrad(t(dy), t(w), t(gate)).cpu().numpy()
        assert relerr(dx, (dy.astype(np.float64) @ w) * (gate > 0)) < 1e-5


def gram_case(device, big=False):
    rs = np.random.RandomState(4)
    # tcgen05 path (NK <= 32, NK + NQ <= 192, D % 64 == 0): one and two row tiles, ragged tails, more tasks than
    # SMs (persistent loop + stage ring across task boundaries), single task; everything else: warp-level kernel
    shapes = [(300, 64, 3, 25, 40), (500, 132, 2, 100, 70), (64, 2048, 2, 5, 100)]
    if big:      # GPU only (minutes under the host emulation)
        shapes += [(700, 256, 331, 25, 160), (700, 128, 1, 32, 160), (700, 192, 5, 1, 7), (700, 64, 150, 7, 121),
                   (300, 96, 3, 25, 40), (300, 64, 3, 25, 200)]
    for precision in ((1, 2) if big else (None,)):       # 2: fp16 (hi, lo) bank planes, scaled by a power of two
        eng = engine_mod.EpisodeEngine(device, precision=precision)
        for (R, D, B, NK, NQ) in shapes:
            feats = (rs.randn(R, D) * (37.0 if precision == 2 else 1.0)).astype(np.float32)
            sup, qry = rs.randint(0, R, size=(B, NK)), rs.randint(0, R, size=(B, NQ))
            t = lambda a: torch.from_numpy(a).to(device)
            g = eng.gram(t(feats), t(sup), t(qry)).cpu().numpy()
            f64 = feats.astype(np.float64)
            for b in range(B):
                rows = np.concatenate([sup[b], qry[b]])
                assert relerr(g[b], f64[rows] @ f64[sup[b]].T) < 1e-5, (precision, R, D, B, NK, NQ, b)


def adam_case(device):
    rs = np.random.RandomState(5)
    shapes = [(7, 13), (64,), (33, 5)]
    ps = [torch.nn.Parameter(torch.from_numpy(rs.randn(*s).astype(np.float32)).to(device)) for s in shapes]
    ref = [p.detach().cpu().clone().requires_grad_(True) for p in ps]
    for decoupled in (False, True):
        mine = FusedAdam(ps, lr=1e-2, weight_decay=0.1, decoupled=decoupled)
        theirs = (torch.optim.AdamW if decoupled else torch.optim.Adam)(ref, lr=1e-2, weight_decay=0.1)
        for it in range(3):
            for p, r in zip(ps, ref):
                gr = torch.from_numpy(rs.randn(*p.shape).astype(np.float32))
                p.grad.copy_(gr.to(device))
                r.grad = gr.clone()
            mine.step()
            theirs.step()
            for p, r in zip(ps, ref):
                assert np.abs(p.detach().cpu().numpy() - r.detach().numpy()).max() < 2e-6


# ------------------------------------------------------------------------------------------------
# device sampler (fumi_sampler_plan + fumi_sampler_expand)
# ------------------------------------------------------------------------------------------------
def device_sampler_golden_case(device, name, N, K, Qtrain):
    """Image ids / labels produced in device memory == the reference loader's goldens, iterators interleaved."""
    import random
    from fumi_b200.data.synth import class_split
    from fumi_b200.sampler import EpisodeSampler
    g, bank = load_golden(name)
    C = bank.text.shape[0]
    B = g["b0_sup_ids"].shape[0]
    samplers = {}
    for split, cats in zip(("train", "val", "test"), class_split(C)):
        samplers[split] = EpisodeSampler(bank.cat_of, cats, N, K, Qtrain if split == "train" else int(100 / N))
    torch.manual_seed(123); np.random.seed(123); random.seed(123)          # main.py:51-53
    maml_mod.PureImageNetwork(im_embed_dim=16, n_way=N, hidden_dims=[256, 64])
    for i, split in enumerate(g["order"]):
        split = str(split)
        if i == 0:
            samplers["val"].new_iterator()
        if i == 1:
            samplers["train"].new_iterator()
            samplers["test"].new_iterator()
        d = samplers[split].expand(samplers[split].plan(B), device)
        for k in ("sup_ids", "qry_ids", "sup_y", "qry_y"):
            assert d[k].device.type == torch.device(device).type
            assert np.array_equal(d[k].cpu().numpy(), g[f"b{i}_{k}"]), (i, split, k)


def device_sampler_vs_host_case(device, sizes_hi=900, B=64):
    """Device expansion == the all-host native sampler (itself pinned to the oracle and the goldens) on ragged
    class sizes incl. one class needing a second 624-word generator block, and the generator states agree."""
    import random
    from fumi_b200.sampler import EpisodeSampler
    rs = np.random.RandomState(3)
    C = 40
    n_c = rs.randint(40, sizes_hi, size=C)
    n_c[3] = 1500
    n_c[7] = 40
    cat_of = np.repeat(np.arange(C), n_c)
    rs.shuffle(cat_of)
    for (N, K, Q) in ((5, 5, 32), (5, 1, 20), (10, 5, 10), (20, 5, 5)):
        s1 = EpisodeSampler(cat_of, np.arange(C), N, K, Q)
        s2 = EpisodeSampler(cat_of, np.arange(C), N, K, Q)
        random.seed(5); torch.manual_seed(5)
        s1.new_iterator()
        want = [s1.next_batch(B) for _ in range(2)]
        st1 = (random.getstate(), torch.get_rng_state())
        random.seed(5); torch.manual_seed(5)
        s2.new_iterator()
        got = [s2.expand(s2.plan(B), device) for _ in range(2)]
        assert random.getstate() == st1[0] and torch.equal(torch.get_rng_state(), st1[1])
        for w, d in zip(want, got):
            for k in w:
                assert np.array_equal(w[k], d[k].cpu().numpy()), (N, K, Q, k)


def device_loader_case(device):
    """EpisodeLoader(device_sampler=True, prefetch=2) hands out the same batches and leaves the same generator
    states as the synchronous all-host loader."""
    import random
    from fumi_b200.data.loader import EpisodeLoader
    from fumi_b200.sampler import EpisodeSampler
    rs = np.random.RandomState(3)
    C, N, K, Q, B = 40, 5, 2, 6, 9
    sizes = rs.randint(K + Q, K + Q + 30, size=C)
    cat_of = np.repeat(np.arange(C), sizes)
    rs.shuffle(cat_of)
    feats = torch.zeros(len(cat_of), 4, device=device)
    ref = []
    for mode in ("host", "device"):
        sampler = EpisodeSampler(cat_of, np.arange(C), N, K, Q, num_threads=2)
        bank = FeatureBank(feats=feats, text=torch.zeros(C, 4, device=device), ids=sampler.ids, categories=np.arange(C))
        loader = EpisodeLoader(bank, sampler, B, pin_memory=False, prefetch=0 if mode == "host" else 2,
                               device_sampler=(mode == "device"))
        random.seed(11); torch.manual_seed(12)
        it = iter(loader)
        for i in range(4):
            b = next(it)
            snap = (random.getstate(), torch.get_rng_state().clone())
            cur = {k: torch.as_tensor(getattr(b, k)).cpu().numpy().copy()
                   for k in ("sup_rows", "qry_rows", "sup_y", "qry_y", "sup_ids", "qry_ids", "head_class")}
            if mode == "host":
                ref.append((cur, snap))
            else:
                assert b.sup_rows.device.type == torch.device(device).type
                for k, v in cur.items():
                    assert np.array_equal(v, ref[i][0][k]), (i, k)
                assert snap[0] == ref[i][1][0] and torch.equal(snap[1], ref[i][1][1])
        loader.close()
    # device sampler + shard=(r, W): slices of the W * B single-stream batch
    W = 3
    sampler = EpisodeSampler(cat_of, np.arange(C), N, K, Q, num_threads=2)
    bank = FeatureBank(feats=feats, text=torch.zeros(C, 4, device=device), ids=sampler.ids, categories=np.arange(C))
    random.seed(31); torch.manual_seed(32)
    whole = next(iter(EpisodeLoader(bank, sampler, W * B, pin_memory=False, device_sampler=True)))
    for r in range(W):
        s2 = EpisodeSampler(cat_of, np.arange(C), N, K, Q, num_threads=2)
        random.seed(31); torch.manual_seed(32)
        part = next(iter(EpisodeLoader(bank, s2, B, pin_memory=False, device_sampler=True, shard=(r, W))))
        for k in ("sup_rows", "qry_rows", "sup_y", "qry_y", "sup_ids", "qry_ids", "head_class"):
            assert torch.equal(getattr(whole, k)[r * B:(r + 1) * B].cpu(), getattr(part, k).cpu()), (r, k)
