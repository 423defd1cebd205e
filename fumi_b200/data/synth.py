"""Synthetic iNat-Anim-shaped banks (the Zenodo data is unavailable offline).

Recipe: SURVEY.md section 8(d).  One ``np.random.RandomState(seed)`` stream, drawn in this
order: class means ``mu[C,D]``, class sizes, the annotation order ``cat_of[M]`` (image id ==
bank row), the image features (row chunks, same stream), the description embeddings.
Features are non-negative like post-ReLU ResNet pool features.

The class split mirrors the reference (fumi/dataset/data.py:377-386): ``np.random.seed(0)``
shuffle of ``arange(C)``, then 60/20/20 train/val/test, unsorted.
"""
from dataclasses import dataclass

import numpy as np

# Paper-scale iNat-Anim shape (recalled in SURVEY.md 8(d); parameters, not constants).
INAT_ANIM_IMAGES = 195_605
INAT_ANIM_CLASSES = 673


@dataclass
class SynthBank:
    feats: np.ndarray      # f32 [M, D]  image-feature bank, row == image id
    text: np.ndarray       # f32 [C, T]  description embedding of category c
    cat_of: np.ndarray     # i64 [M]     category of image id


def make_bank(num_images=INAT_ANIM_IMAGES, num_classes=INAT_ANIM_CLASSES, im_dim=2048,
              text_dim=768, min_per_class=60, seed=2022, chunk_rows=8192) -> SynthBank:
    C, M, D, T = num_classes, num_images, im_dim, text_dim
    if M < min_per_class * C:
        raise ValueError(f"num_images {M} < min_per_class*num_classes {min_per_class * C}")
    rs = np.random.RandomState(seed)
    mu = rs.randn(C, D).astype(np.float32)
    extra = M - min_per_class * C
    n_c = min_per_class + rs.multinomial(extra, rs.dirichlet(2.0 * np.ones(C)))
    cat_of = np.repeat(np.arange(C, dtype=np.int64), n_c)
    rs.shuffle(cat_of)
    feats = np.empty((M, D), dtype=np.float32)
    for r0 in range(0, M, chunk_rows):
        r1 = min(M, r0 + chunk_rows)
        noise = rs.randn(r1 - r0, D).astype(np.float32)
        np.maximum(0.5 * mu[cat_of[r0:r1]] + noise, 0.0, out=feats[r0:r1])
    text = (0.3 * rs.randn(C, T)).astype(np.float32)
    return SynthBank(feats=feats, text=text, cat_of=cat_of)


def class_split(num_classes: int):
    """(train, val, test) category arrays exactly as fumi/dataset/data.py:377-386."""
    perm = np.arange(num_classes)
    np.random.RandomState(0).shuffle(perm)   # == np.random.seed(0); np.random.shuffle(...)
    a, b = int(0.6 * num_classes), int(0.8 * num_classes)
    return perm[:a].copy(), perm[a:b].copy(), perm[b:].copy()
