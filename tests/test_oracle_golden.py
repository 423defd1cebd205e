"""The CPU oracles (oracle/) against the golden vectors the reference itself produced."""
import random

import numpy as np
import pytest
import torch

from oracle import episode_np, episode_torch, sampler_np
from fumi_b200.data.synth import class_split

from helpers import argv_int, flat_batch, load_golden, params_of, relerr

TOL = 1e-4          # north_star: fp32 logits / adapted weights within 1e-4 relative error


@pytest.mark.parametrize("name", ["fumi_train_n5k5_d512", "fumi_train_n5k5_d512_tanh", "fumi_train_n20k5_d512"])
def test_fumi_train_numpy_oracle(name):
    g, bank = load_golden(name)
    N = argv_int(g, "--num_ways", 5)
    batch = flat_batch(g, bank, N)
    tanh = "--norm_hypernet" in str(g["argv"])
    r = episode_np.fumi_batch(params_of(g), batch, float(g["alpha"]), int(g["steps"]), tanh=tanh,
                              want_grad=True, dtype=np.float32)
    assert np.array_equal(r["preds"], g["preds"])
    assert relerr(r["logits"], g["logits"]) < TOL
    assert abs(float(r["loss"]) - float(g["loss"])) < TOL * abs(float(g["loss"]))
    assert abs(float(r["acc"]) - float(g["acc"])) < 1e-6
    for b, t in enumerate(r["tasks"]):
        assert relerr(t["hp0"], g["hp0"][b]) < TOL
        W0, b0, W1, b1, hp = t["adapted"]
        assert relerr(hp, g["hp_adapted"][b]) < TOL
        assert relerr(W1, g["W1_adapted"][b]) < TOL
        assert relerr(b1, g["b1_adapted"][b]) < TOL
        assert relerr(b0, g["b0_adapted"][b]) < TOL
    assert relerr(r["tasks"][0]["adapted"][0][:16], g["W0_adapted_t0_rows16"]) < TOL
    gmax = max(np.abs(g["grad:" + k]).max() for k in r["grads"])
    for k, v in r["grads"].items():
        ref = g["grad:" + k]
        if k == "hyper_net.2.bias":
            # true gradient is exactly zero (softmax shift invariance): absolute check
            assert np.abs(v - ref).max() < 1e-4 * gmax
        else:
            assert relerr(v, ref) < 2e-4, k
    # Adam(lr, L2 weight decay), first step
    lr, wd = float(g["lr"]), float(g["wd"])
    for k, gr in r["grads"].items():
        p = g["param:" + k]
        p1, _, _ = episode_np.adam_step(p, g["grad:" + k], np.zeros_like(p), np.zeros_like(p), 1, lr, wd)
        assert np.abs(p1 - g["post:" + k]).max() <= 1e-7 + 1e-6 * np.abs(p).max(), k


@pytest.mark.parametrize("name", ["fumi_test_n5k1_full", "fumi_test_n5k5_full"])
def test_fumi_test_numpy_oracle(name):
    g, bank = load_golden(name)
    gp, _ = load_golden("fumi_test_n5k1_full")          # the full-dims parameter set is stored once
    params = params_of(gp)
    batch = flat_batch(g, bank, 5)
    r = episode_np.fumi_batch(params, batch, float(g["alpha"]), int(g["steps"]), dtype=np.float32)
    assert int(g["steps"]) == 100
    assert np.array_equal(r["preds"], g["preds"])
    assert relerr(r["logits"], g["logits"]) < TOL
    for b, t in enumerate(r["tasks"]):
        assert relerr(t["adapted"][4], g["hp_adapted"][b]) < TOL
        assert relerr(t["adapted"][2], g["W1_adapted"][b]) < TOL


@pytest.mark.parametrize("name", ["maml_train_n5k5_d512", "maml_train_n5k5_d512_fo", "maml_test_n5k5_d512"])
def test_maml_numpy_oracle(name):
    g, bank = load_golden(name)
    batch = flat_batch(g, bank, 5)
    train = "train" in name
    r = episode_np.maml_batch(params_of(g), batch, float(g["alpha"]), int(g["steps"]),
                              first_order=bool(g["first_order"]), want_grad=train, dtype=np.float32)
    assert np.array_equal(r["preds"], g["preds"])
    assert relerr(r["logits"], g["logits"]) < TOL
    assert abs(float(r["loss"]) - float(g["loss"])) < TOL
    if train:
        for k, v in r["grads"].items():
            assert relerr(v, g["grad:" + k]) < 2e-4, k


def test_fumi_torch_port_matches_golden():
    g, bank = load_golden("fumi_train_n5k5_d512")
    batch = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in flat_batch(g, bank, 5).items()}
    params = {k: torch.from_numpy(v.copy()).requires_grad_(True) for k, v in params_of(g).items()}
    r = episode_torch.fumi_batch(params, batch, float(g["alpha"]), int(g["steps"]), train=True)
    assert np.array_equal(r["preds"].numpy(), g["preds"])
    assert relerr(r["logits"].numpy(), g["logits"]) < TOL
    for k, p in params.items():
        if k != "hyper_net.2.bias":
            assert relerr(p.grad.numpy(), g["grad:" + k]) < 2e-4, k


def test_am3_numpy_oracle():
    g, bank = load_golden("am3_test_n10k5_d512")
    batch = flat_batch(g, bank, 10)
    r = episode_np.am3_batch(params_of(g), batch, 10, dtype=np.float32)
    assert np.array_equal(r["preds"], g["preds"])
    assert abs(float(r["loss"]) - float(g["loss"])) < TOL * abs(float(g["loss"]))
    assert abs(r["acc"] - float(g["acc"])) < 1e-9
    assert relerr(r["lamda"], g["sup_lamda"]) < TOL
    assert abs(float(r["avg_lamda"]) - float(g["avg_lamda"])) < 1e-6


@pytest.mark.parametrize("name,N,K,Qtrain", [("sampler_n5k5b4", 5, 5, 32), ("sampler_n5k1b3", 5, 1, 32),
                                              ("sampler_n10k5b2", 10, 5, 20), ("sampler_n20k5b2", 20, 5, 16)])
def test_flat_sampler_matches_reference_loader(name, N, K, Qtrain):
    """Image ids and labels bit-exact, with val/train/test iterators interleaved (Appendix B.5)."""
    g, bank = load_golden(name)
    C = bank.text.shape[0]
    B = g["b0_sup_ids"].shape[0]
    splits = dict(zip(("train", "val", "test"), class_split(C)))
    samplers = {}
    for split, cats in splits.items():
        Q = Qtrain if split == "train" else int(100 / N)
        samplers[split] = sampler_np.FlatSampler(sampler_np.class_tables(bank.cat_of, cats), N, K, Q)
    # main.py:51-53 seeds after the loaders are built (seed flag default 123)
    torch.manual_seed(123); np.random.seed(123); random.seed(123)
    # model init consumes the torch stream before any iterator exists
    from fumi_b200.maml import PureImageNetwork
    PureImageNetwork(im_embed_dim=16, n_way=N, hidden_dims=[256, 64])
    # iterator creation order of oracle/make_golden.py: iter(val), 1 val batch, iter(train), iter(test), ...
    for i, split in enumerate(g["order"]):
        split = str(split)
        if i == 0:
            samplers["val"].new_iterator()
        if i == 1:
            samplers["train"].new_iterator()
            samplers["test"].new_iterator()
        b = samplers[split].next_batch(B)
        assert np.array_equal(b["sup_ids"], g[f"b{i}_sup_ids"]), (i, split)
        assert np.array_equal(b["qry_ids"], g[f"b{i}_qry_ids"]), (i, split)
        assert np.array_equal(b["sup_targets"], g[f"b{i}_sup_y"]), (i, split)
        assert np.array_equal(b["qry_targets"], g[f"b{i}_qry_y"]), (i, split)


def test_macro_scores_equal_sklearn():
    """AM3.evaluate's metrics come from device confusion counts (fumi_confusion_counts); the host formulas must equal
    sklearn's accuracy_score / precision_recall_fscore_support(average='macro') (utils/utils.py:323-326) exactly."""
    import warnings
    from sklearn.metrics import accuracy_score, precision_recall_fscore_support
    from fumi_b200.am3 import macro_scores
    rng = np.random.default_rng(5)
    for N, n, hi in ((10, 4096, 10), (5, 64, 3), (20, 500, 20)):
        y, p = rng.integers(0, N, n), rng.integers(0, hi, n)
        cm = np.bincount(y * N + p, minlength=N * N).reshape(N, N)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref = (accuracy_score(y, p),) + precision_recall_fscore_support(y, p, average="macro")[:3]
        assert macro_scores(cm) == ref
