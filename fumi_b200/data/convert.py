"""One-shot converter from the reference's on-disk inputs to the HBM bank format (`<data_dir>/iNat-Anim/bank.npz`).

Reference inputs (fumi/dataset/data.py:373-430):
  * `inat_anim.json`: `categories[i].{name, common_name, description}`, `images[i].id`,
    `annotations[id].category_id` (the annotation list is indexed by image id, data.py:395-402);
  * `image_embeddings_{resnet-152,resnet-34}.hdf5['images']`: f32 [M, D], row = image id (data.py:429-430, 545);
  * one description embedding per category (the reference computes BERT pooled outputs at load time with
    `--precompute_bert`, data.py:472-495; the weights are a download, so here they come precomputed as `[C, T]`
    in category order, `.npy`).
Output arrays: feats [M, D] f32 (row = image id), text [C, T] f32, cat_of [M] i64 -- what
`fumi_b200.data.loader.load_arrays` reads.  The class split itself (np.random.seed(0) shuffle, 60/20/20) is applied
by the loaders exactly as data.py:377-386 does.

    python -m fumi_b200.data.convert --json inat_anim.json --images image_embeddings_resnet-152.hdf5 \\
        --text description_embeddings.npy --out <data_dir>/iNat-Anim/bank.npz
"""
import argparse
import json
import os

import numpy as np


def _load_matrix(path, key=None):
    if path.endswith((".hdf5", ".h5")):
        try:
            import h5py
        except ImportError as e:          # not part of this image; the reference environment has it
            raise ImportError("reading the reference's .hdf5 embeddings needs h5py (pip install h5py), "
                              "or export the dataset to .npy first") from e
        with h5py.File(path, "r") as f:
            return np.asarray(f[key or "images"], dtype=np.float32)
    if path.endswith(".npz"):
        return np.asarray(np.load(path)[key or "images"], dtype=np.float32)
    return np.asarray(np.load(path), dtype=np.float32)


def convert_inat_anim(json_path, image_embeddings, text_embeddings, out_path):
    with open(json_path) as f:
        ann = json.load(f)
    num_classes, num_images = len(ann["categories"]), len(ann["images"])
    ids = np.asarray([im["id"] for im in ann["images"]], np.int64)
    if not np.array_equal(np.sort(ids), np.arange(num_images)):
        raise ValueError("image ids must be 0..M-1 (they index the embedding matrix and the annotation list)")
    cat_of = np.empty(num_images, np.int64)
    for i in ids:                                     # annotations[id]['category_id'], data.py:397-402
        cat_of[i] = ann["annotations"][int(i)]["category_id"]
    if cat_of.min() < 0 or cat_of.max() >= num_classes:
        raise ValueError("category ids must be 0..C-1")
    feats = _load_matrix(image_embeddings, "images")
    text = _load_matrix(text_embeddings, "text")
    if feats.ndim != 2 or feats.shape[0] != num_images:
        raise ValueError(f"image embeddings are {feats.shape}, expected [{num_images}, D]")
    if text.ndim != 2 or text.shape[0] != num_classes:
        raise ValueError(f"description embeddings are {text.shape}, expected [{num_classes}, T] in category order")
    os.makedirs(os.path.dirname(os.path.abspath(out_path)), exist_ok=True)
    np.savez(out_path, feats=feats, text=text, cat_of=cat_of)
    return dict(num_images=num_images, num_classes=num_classes, im_dim=int(feats.shape[1]), text_dim=int(text.shape[1]))


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--json", required=True)
    ap.add_argument("--images", required=True, help=".hdf5 (dataset 'images'), .npy or .npz")
    ap.add_argument("--text", required=True, help="[C, T] description embeddings in category order (.npy)")
    ap.add_argument("--out", required=True)
    a = ap.parse_args(argv)
    print(convert_inat_anim(a.json, a.images, a.text, a.out))


if __name__ == "__main__":
    main()
