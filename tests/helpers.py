"""Shared test helpers: golden fixtures -> regenerated banks -> flat episode batches."""
import functools
import hashlib
import os

import numpy as np

from fumi_b200.data.synth import make_bank

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
BANK_KEYS = ("num_images", "num_classes", "im_dim", "text_dim", "min_per_class", "seed")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


@functools.lru_cache(maxsize=4)
def _bank(*vals):
    return make_bank(**dict(zip(BANK_KEYS, vals)))


def load_golden(name):
    g = dict(np.load(os.path.join(GOLD, name + ".npz"), allow_pickle=False))
    bank = _bank(*[int(g["bank_" + k]) for k in BANK_KEYS])
    assert sha(bank.feats) == str(g["bank_feats_sha"]), "synthetic bank generator drifted"
    assert sha(bank.text) == str(g["bank_text_sha"])
    return g, bank


def params_of(g, prefix="param:"):
    return {k[len(prefix):]: v for k, v in g.items() if k.startswith(prefix)}


def class_text_rows(bank, sup_ids, sup_y, num_ways):
    """[B,N] category whose description conditions label i (fumi.py:207-210: first support row
    with targets == i)."""
    B = sup_ids.shape[0]
    cats = np.empty((B, num_ways), np.int64)
    for b in range(B):
        for i in range(num_ways):
            j = int(np.nonzero(sup_y[b] == i)[0][0])
            cats[b, i] = bank.cat_of[sup_ids[b, j]]
    return cats


def flat_batch(g, bank, num_ways):
    cats = class_text_rows(bank, g["sup_ids"], g["sup_y"], num_ways)
    return dict(sup_x=bank.feats[g["sup_ids"]], sup_y=g["sup_y"], qry_x=bank.feats[g["qry_ids"]],
                qry_y=g["qry_y"], class_text=bank.text[cats], class_cats=cats,
                sup_text=bank.text[bank.cat_of[g["sup_ids"]]])


def relerr(a, b):
    """max |a-b| / max |b|  (tensor-normalised relative error; tolerance 1e-4 per BASELINE.json)."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def argv_int(g, flag, default):
    toks = str(g["argv"]).split()
    return int(toks[toks.index(flag) + 1]) if flag in toks else default
