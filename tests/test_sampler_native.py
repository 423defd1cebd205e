"""Native (C++) episodic sampler: bit-exact against the CPU oracle and the reference's goldens."""
import random

import numpy as np
import pytest
import torch

from fumi_b200 import _lib
from fumi_b200.data.synth import class_split, make_bank
from fumi_b200.sampler import EpisodeSampler
from oracle import sampler_np

from helpers import load_golden


def test_py_tuple_hash_matches_cpython():
    rs = np.random.RandomState(1)
    for n in (1, 2, 5, 10, 20):
        for _ in range(50):
            t = tuple(int(x) for x in rs.randint(0, 700, size=n))
            a = np.asarray(t, np.int64)
            assert _lib.lib().fumi_py_tuple_hash(_lib.ptr(a), n) == hash(t)


@pytest.mark.parametrize("C,N,K,Q,B", [(60, 5, 5, 32, 7), (60, 5, 1, 20, 3), (30, 10, 5, 10, 4), (403, 20, 5, 32, 5),
                                       (21, 5, 2, 3, 6), (26, 6, 2, 3, 6), (85, 6, 2, 3, 6), (86, 6, 2, 3, 6)])
def test_native_sampler_equals_oracle_streams(C, N, K, Q, B):
    """Same ids / labels AND the same generator states afterwards (pool / set-rejection paths of
    random.sample are both crossed: setsize = 21 (+ 4**ceil(log4(3k)) for k > 5))."""
    rs = np.random.RandomState(C)
    sizes = rs.randint(K + Q, K + Q + 40, size=C)
    cat_of = np.repeat(np.arange(C), sizes)
    rs.shuffle(cat_of)
    cats = rs.permutation(C)
    oracle = sampler_np.FlatSampler(sampler_np.class_tables(cat_of, cats), N, K, Q)
    native = EpisodeSampler(cat_of, cats, N, K, Q, num_threads=3)
    for rnd in range(3):
        random.seed(5 + rnd); torch.manual_seed(9 + rnd)
        torch.rand(617 * rnd + 3)                       # move the torch stream near/over a twist
        oracle.new_iterator()
        want = [oracle.next_batch(B) for _ in range(2)]
        end_py, end_t = random.getstate(), torch.get_rng_state()
        random.seed(5 + rnd); torch.manual_seed(9 + rnd)
        torch.rand(617 * rnd + 3)
        native.new_iterator()
        got = [native.next_batch(B) for _ in range(2)]
        assert random.getstate() == end_py
        assert torch.equal(torch.get_rng_state(), end_t)
        for w, g in zip(want, got):
            assert np.array_equal(g["classes"], w["classes"])
            assert np.array_equal(g["label_perm"], w["label_perm"])
            assert np.array_equal(g["sup_ids"], w["sup_ids"])
            assert np.array_equal(g["qry_ids"], w["qry_ids"])
            assert np.array_equal(g["sup_y"], w["sup_targets"])
            assert np.array_equal(g["qry_y"], w["qry_targets"])
            assert np.array_equal(native.ids[g["sup_rows"]], g["sup_ids"])
            assert np.array_equal(native.ids[g["qry_rows"]], g["qry_ids"])
            for b in range(B):
                for i in range(N):
                    assert g["head_class"][b, i] == g["classes"][b, list(g["label_perm"][b]).index(i)]


def test_class_too_small_raises_value_error():
    cat_of = np.repeat(np.arange(10), 5)
    s = EpisodeSampler(cat_of, np.arange(10), 3, 2, 4)
    random.seed(0); torch.manual_seed(0)
    with pytest.raises(ValueError, match="smaller than the minimum"):
        s.next_batch(2)


@pytest.mark.parametrize("name,N,K,Qtrain", [("sampler_n5k5b4", 5, 5, 32), ("sampler_n5k1b3", 5, 1, 32),
                                              ("sampler_n10k5b2", 10, 5, 20), ("sampler_n20k5b2", 20, 5, 16)])
def test_native_sampler_matches_reference_loader_golden(name, N, K, Qtrain):
    """Image ids and labels bit-exact vs the reference loader, iterators interleaved (B.5)."""
    from fumi_b200.maml import PureImageNetwork
    g, bank = load_golden(name)
    C = bank.text.shape[0]
    B = g["b0_sup_ids"].shape[0]
    samplers = {}
    for split, cats in zip(("train", "val", "test"), class_split(C)):
        Q = Qtrain if split == "train" else int(100 / N)
        samplers[split] = EpisodeSampler(bank.cat_of, cats, N, K, Q)
    torch.manual_seed(123); np.random.seed(123); random.seed(123)          # main.py:51-53
    PureImageNetwork(im_embed_dim=16, n_way=N, hidden_dims=[256, 64])       # init consumes the torch stream
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
        assert np.array_equal(b["sup_y"], g[f"b{i}_sup_y"]), (i, split)
        assert np.array_equal(b["qry_y"], g[f"b{i}_qry_y"]), (i, split)


def test_prefetching_loader_hands_out_the_same_stream_and_states():
    """EpisodeLoader(prefetch=2) == synchronous loader: same batches, and after each hand-out the global
    generator states equal the synchronous ones."""
    from fumi_b200.data.bank import FeatureBank
    from fumi_b200.data.loader import EpisodeLoader
    rs = np.random.RandomState(3)
    C, N, K, Q, B = 40, 5, 2, 6, 9
    sizes = rs.randint(K + Q, K + Q + 30, size=C)
    cat_of = np.repeat(np.arange(C), sizes)
    rs.shuffle(cat_of)
    feats = torch.zeros(len(cat_of), 4)
    ref_states, ref_batches = [], []
    for prefetch in (0, 2):
        sampler = EpisodeSampler(cat_of, np.arange(C), N, K, Q, num_threads=2)
        bank = FeatureBank(feats=feats, text=torch.zeros(C, 4), ids=sampler.ids, categories=np.arange(C))
        loader = EpisodeLoader(bank, sampler, B, pin_memory=False, prefetch=prefetch)
        random.seed(11); torch.manual_seed(12)
        it = iter(loader)
        for i in range(4):
            b = next(it)
            snap = (random.getstate(), torch.get_rng_state().clone())
            if prefetch == 0:
                ref_batches.append({k: v.copy() for k, v in b.host.items()})
                ref_states.append(snap)
            else:
                for k, v in b.host.items():
                    assert np.array_equal(v, ref_batches[i][k]), (i, k)
                assert snap[0] == ref_states[i][0]
                assert torch.equal(snap[1], ref_states[i][1])
        loader.close()


def test_sharded_loaders_draw_one_global_stream():
    """EpisodeLoader(shard=(r, W)): the ranks' batches are the slices of the single-process batch of W * B tasks
    (same global generator streams on every rank), with and without the prefetch thread."""
    from fumi_b200.data.bank import FeatureBank
    from fumi_b200.data.loader import EpisodeLoader
    rs = np.random.RandomState(4)
    C, N, K, Q, B, W = 30, 5, 2, 4, 3, 2
    sizes = rs.randint(K + Q, K + Q + 20, size=C)
    cat_of = np.repeat(np.arange(C), sizes)
    rs.shuffle(cat_of)
    feats = torch.zeros(len(cat_of), 4)

    def run(batch, shard, prefetch):
        sampler = EpisodeSampler(cat_of, np.arange(C), N, K, Q, num_threads=2)
        bank = FeatureBank(feats=feats, text=torch.zeros(C, 4), ids=sampler.ids, categories=np.arange(C))
        loader = EpisodeLoader(bank, sampler, batch, pin_memory=False, prefetch=prefetch, shard=shard)
        random.seed(21); torch.manual_seed(22)
        it = iter(loader)
        out = [{k: np.asarray(v).copy() for k, v in next(it).host.items()} for _ in range(3)]
        loader.close()
        return out, random.getstate(), torch.get_rng_state().clone()

    whole, py_w, t_w = run(W * B, None, 0)
    for prefetch in (0, 2):
        for r in range(W):
            part, py_r, t_r = run(B, (r, W), prefetch)
            for a, b in zip(whole, part):
                for k in a:
                    assert np.array_equal(a[k][r * B:(r + 1) * B], b[k]), (prefetch, r, k)
            assert py_r == py_w and torch.equal(t_r, t_w)
