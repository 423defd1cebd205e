"""On-disk input conversion: inat_anim.json + embedding matrices -> bank.npz -> loaders."""
import json
import os
import types

import numpy as np
import pytest
import torch

from fumi_b200.data.convert import convert_inat_anim
from fumi_b200.data.loader import load_arrays
from fumi_b200.sampler import class_tables


def test_convert_inat_anim_roundtrip(tmp_path):
    rs = np.random.RandomState(0)
    C, M, D, T = 12, 90, 8, 6
    cat = rs.randint(0, C, size=M)
    order = rs.permutation(M)                                        # the image list need not be sorted by id
    ann = {"categories": [{"name": f"c{i}", "common_name": f"n{i}", "description": f"d{i}"} for i in range(C)],
           "images": [{"id": int(i)} for i in order],
           "annotations": [{"category_id": int(c)} for c in cat]}
    feats, text = rs.randn(M, D).astype(np.float32), rs.randn(C, T).astype(np.float32)
    jp, ip, tp = tmp_path / "inat_anim.json", tmp_path / "im.npy", tmp_path / "text.npy"
    jp.write_text(json.dumps(ann)); np.save(ip, feats); np.save(tp, text)
    out = tmp_path / "data" / "iNat-Anim" / "bank.npz"
    info = convert_inat_anim(str(jp), str(ip), str(tp), str(out))
    assert info == dict(num_images=M, num_classes=C, im_dim=D, text_dim=T)
    args = types.SimpleNamespace(synthetic=False, data_dir=str(tmp_path / "data"))
    f2, t2, c2 = load_arrays(args)
    assert np.array_equal(f2, feats) and np.array_equal(t2, text) and np.array_equal(c2, cat)
    # class tables as dataset/data.py:395-414 builds them: ascending image ids per category
    offsets, ids = class_tables(c2, np.arange(C))
    for c in range(C):
        assert np.array_equal(ids[offsets[c]:offsets[c + 1]], np.flatnonzero(cat == c))


def test_convert_rejects_bad_inputs(tmp_path):
    ann = {"categories": [{}], "images": [{"id": 0}, {"id": 2}], "annotations": [{"category_id": 0}] * 3}
    jp = tmp_path / "a.json"
    jp.write_text(json.dumps(ann))
    np.save(tmp_path / "im.npy", np.zeros((2, 4), np.float32)); np.save(tmp_path / "t.npy", np.zeros((1, 4), np.float32))
    with pytest.raises(ValueError, match="image ids"):
        convert_inat_anim(str(jp), str(tmp_path / "im.npy"), str(tmp_path / "t.npy"), str(tmp_path / "o.npz"))
    with pytest.raises(ImportError, match="h5py"):
        ann["images"] = [{"id": 0}, {"id": 1}]
        jp.write_text(json.dumps(ann))
        convert_inat_anim(str(jp), str(tmp_path / "missing.hdf5"), str(tmp_path / "t.npy"), str(tmp_path / "o.npz"))
