"""Flat CPU restatement of the reference's episodic task sampler (TEST INFRASTRUCTURE).

Follows fumi/dataset/data.py:125-188 (get_inat_anim: ClassSplitter(shuffle=True, K, Q).seed(0)),
data.py:73-84 (BatchMetaDataLoader(shuffle=True)), data.py:377-414 (class split + per-class
ascending image-id arrays), data.py:533-581 (per-sample payload) and the torchmeta==1.7.0
pieces they call (restated in oracle/torchmeta_shim.py).

It uses the *real* generators the reference uses -- Python ``random`` (class tuples), numpy
``RandomState`` (hash-seeded per-class permutation + the split's shared RandomState(0)
shuffles) and the torch global CPU generator (label permutation, DataLoader iterator seed) --
so the product's native MT19937 re-implementation can be checked against it bit-exactly.

Sampler parity is pinned against the reference's own loader driven through the torchmeta
restatement (tests/test_oracle_vs_reference.py); PARITY UNPINNED against real torchmeta.
"""
import random

import numpy as np
import torch


def class_tables(cat_of: np.ndarray, categories: np.ndarray):
    """Per split-class ascending image ids (data.py:395-414). cat_of[i] = category of image i."""
    order = np.argsort(cat_of, kind="stable")
    sorted_cat = cat_of[order]
    starts = np.searchsorted(sorted_cat, categories, side="left")
    ends = np.searchsorted(sorted_cat, categories, side="right")
    return [order[s:e].astype(np.int64) for s, e in zip(starts, ends)]


class FlatSampler:
    """One split's loader.  ``ids_per_class[c]`` = ascending image ids of split-class c."""

    def __init__(self, ids_per_class, num_ways, num_shots, num_query):
        self.ids = ids_per_class
        self.C = len(ids_per_class)
        self.N, self.K, self.Q = num_ways, num_shots, num_query
        self.shared = np.random.RandomState(0)          # train_split.seed(0), data.py:150/167/184

    @staticmethod
    def new_iterator():
        """iter(loader): DataLoader draws its base seed from the torch global generator."""
        torch.empty((), dtype=torch.int64).random_()

    def next_batch(self, batch_size):
        N, K, Q = self.N, self.K, self.Q
        tuples = [tuple(random.sample(range(self.C), N)) for _ in range(batch_size)]
        sup = np.empty((batch_size, N * K), np.int64)
        qry = np.empty((batch_size, N * Q), np.int64)
        for b, tup in enumerate(tuples):
            h = hash(tup)
            for p, c in enumerate(tup):
                n_c = len(self.ids[c])
                if n_c < K + Q:
                    raise ValueError(f"The number of samples for one class ({n_c}) is smaller than the "
                                     f"minimum number of samples per class required ({K + Q}).")
                seed = (h + c + 0) % (2 ** 32)
                perm = np.random.RandomState(seed).permutation(n_c)
                s = perm[:K]
                self.shared.shuffle(s)
                q = perm[K:K + Q]
                self.shared.shuffle(q)
                sup[b, p * K:(p + 1) * K] = self.ids[c][s]
                qry[b, p * Q:(p + 1) * Q] = self.ids[c][q]
        label_perm = np.stack([torch.randperm(N).numpy() for _ in range(batch_size)])
        return {
            "classes": np.asarray(tuples, np.int64),              # [B,N] split-class per tuple position
            "label_perm": label_perm.astype(np.int64),            # [B,N] label of tuple position p
            "sup_ids": sup, "qry_ids": qry,                       # global image ids (bank rows)
            "sup_targets": np.repeat(label_perm, K, axis=1).astype(np.int64),
            "qry_targets": np.repeat(label_perm, Q, axis=1).astype(np.int64),
        }
