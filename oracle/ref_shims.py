"""Run the UNMODIFIED reference (/root/reference/fumi) in this container (TEST INFRASTRUCTURE).

Recipe: SURVEY.md Appendix C.  Zero edits to reference files; everything missing from the
image is stubbed in ``sys.modules`` *before* the reference modules are imported:

* torchmeta           -> oracle.torchmeta_shim (restated algorithm)
* h5py.File(p)['images'] -> np.load(p, mmap) (fancy row indexing works the same)
* gensim / nltk       -> empty stubs (only used by the out-of-scope GloVe/w2v encoders)
* transformers.AdamW  -> torch.optim.AdamW (symbol removed in transformers 5.x)
* BertTokenizer/BertModel -> fakes: description string "i" -> token [[i]] ->
  last_hidden_state = text_bank[i][None]  (data.py:441-449, 472-495)
* wandb               -> inert stub (evaluate() never touches it; the loops do)
* FuMI's in-place ``hyper_params -= ...`` (fumi.py:168) raises on torch>=2; the hypernet
  output is returned as a Tensor subclass whose ``__isub__`` is out-of-place, which is the
  torch-1.8.1 semantics the authors ran (Trap 2).

The reference cannot travel to the GPU box: this module is used only HERE, by
``oracle/make_golden.py`` and by tests that skip when /root/reference is absent.
"""
import importlib
import importlib.machinery
import json
import os
import sys
import types

import numpy as np
import torch

from . import torchmeta_shim

REF_ROOT = os.environ.get("FUMI_REF", "/root/reference/fumi")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "models", "fumi.py"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


class _H5File(object):
    """h5py.File stand-in: '<path>' is an .npy file holding the 'images' dataset."""

    def __init__(self, path, mode="r"):
        self._arr = np.load(path if os.path.exists(path) else path + ".npy", mmap_mode="r")

    def __getitem__(self, key):
        assert key == "images"
        return self._arr


class _FakeTokens(dict):
    pass


class FakeBertTokenizer(object):
    @classmethod
    def from_pretrained(cls, name):
        return cls()

    def __call__(self, descriptions, **kw):
        ids = torch.tensor([[int(d)] for d in descriptions], dtype=torch.int64)
        return _FakeTokens(input_ids=ids, attention_mask=torch.ones_like(ids))


class FakeBertModel(object):
    text_bank = None  # f32 [C, T], set by build_dataset_dir()

    class _Cfg(object):
        hidden_size = 768

    @classmethod
    def from_pretrained(cls, name):
        m = cls()
        m.config = cls._Cfg()
        m.config.hidden_size = int(cls.text_bank.shape[1])
        return m

    def to(self, device):
        return self

    def __call__(self, input_ids=None, attention_mask=None, output_attentions=False):
        out = types.SimpleNamespace()
        out.last_hidden_state = torch.from_numpy(
            np.asarray(FakeBertModel.text_bank)[input_ids[:, 0].numpy()]).unsqueeze(1)
        return out


class OutOfPlace(torch.Tensor):
    """Tensor whose ``-=`` rebinds instead of mutating (fumi.py:168 under torch>=2)."""

    def __isub__(self, other):
        return self - other


_loaded = {}


def load_reference():
    """Import the reference's modules; returns a namespace (fumi, maml, am3, utils, data)."""
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    if not reference_available():
        raise RuntimeError(f"reference not found under {REF_ROOT}")
    torchmeta_shim.install()
    _stub("h5py", File=_H5File)
    g = _stub("gensim")
    g.corpora = _stub("gensim.corpora")
    g.utils = _stub("gensim.utils", tokenize=lambda s: s.split())
    g.downloader = _stub("gensim.downloader")
    n = _stub("nltk", download=lambda *a, **k: None)
    n.corpus = _stub("nltk.corpus", stopwords=types.SimpleNamespace(words=lambda lang: []))
    run = types.SimpleNamespace(dir="/tmp/fumi_oracle_wandb", name="oracle")
    os.makedirs(run.dir, exist_ok=True)
    _stub("wandb", init=lambda **k: run, log=lambda *a, **k: None, watch=lambda *a, **k: None,
          save=lambda *a, **k: None, finish=lambda: None, run=run,
          config=types.SimpleNamespace(update=lambda *a, **k: None),
          restore=lambda *a, **k: None)
    import transformers
    from transformers import BertModel, BertTokenizer  # noqa: F401  (forces the lazy module)
    transformers.AdamW = torch.optim.AdamW
    sys.modules["transformers"].AdamW = torch.optim.AdamW
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    # the reference uses top-level package names `models`, `utils`, `dataset` (main.py:11-16)
    for k in [k for k in sys.modules if k in ("models", "utils", "dataset")
              or k.startswith(("models.", "utils.", "dataset."))]:
        del sys.modules[k]
    data = importlib.import_module("dataset.data")
    data.BertTokenizer = FakeBertTokenizer
    data.BertModel = FakeBertModel
    utils = importlib.import_module("utils.utils")
    fumi = importlib.import_module("models.fumi")
    maml = importlib.import_module("models.maml")
    am3 = importlib.import_module("models.am3")
    # FuMI only: rebinding `-=` (Appendix C step 5); no reference file is edited.
    if not getattr(fumi.FUMI, "_oop_patched", False):
        orig_forward = fumi.FUMI.forward

        def forward(self, text_embed):
            return orig_forward(self, text_embed).as_subclass(OutOfPlace)
        fumi.FUMI.forward = forward
        fumi.FUMI._oop_patched = True
    _loaded.update(fumi=fumi, maml=maml, am3=am3, utils=utils, data=data)
    return types.SimpleNamespace(**_loaded)


def build_dataset_dir(root, bank, image_embedding_model="resnet-152"):
    """Write ``<root>/iNat-Anim/{inat_anim.json, image_embeddings_*.hdf5(.npy)}`` for ``bank``.

    JSON schema read at data.py:373-418: categories[i].{name,common_name,description},
    images[i].id, annotations[id].category_id.  description "i" indexes FakeBertModel.text_bank.
    """
    d = os.path.join(root, "iNat-Anim")
    os.makedirs(d, exist_ok=True)
    C, M = bank.text.shape[0], bank.feats.shape[0]
    ann = {
        "categories": [{"name": str(c), "common_name": str(c), "description": str(c)} for c in range(C)],
        "images": [{"id": i} for i in range(M)],
        "annotations": [{"category_id": int(bank.cat_of[i])} for i in range(M)],
    }
    with open(os.path.join(d, "inat_anim.json"), "w") as f:
        json.dump(ann, f)
    np.save(os.path.join(d, f"image_embeddings_{image_embedding_model}.hdf5.npy"), bank.feats)
    FakeBertModel.text_bank = bank.text
    return root


def make_args(ref, data_dir, argv=()):
    """Reference argparse (utils.py:19-229) + args.device as main.parse_args (main.py:141-149)."""
    args = ref.utils.parser().parse_args(["--data_dir", data_dir, "--disable_cuda", "--wandb_offline",
                                          *argv])
    args.device = torch.device("cpu")
    return args
