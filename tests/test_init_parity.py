"""Initial weights: `torch.manual_seed(args.seed)` followed by the model constructor must reproduce the reference's
initial state_dict bit for bit (same layer creation order and default init; SURVEY.md A.5, main.py:51-53,
utils.py:232-266), with the reference's state_dict key names in the reference's order (checkpoint compatibility).
The golden `param:*` arrays are the unmodified reference's state_dict right after init (oracle/make_golden.py)."""
import numpy as np
import pytest
import torch

from fumi_b200 import utils
from helpers import load_golden, params_of

CASES = [("fumi", "fumi_train_n5k5_d512"), ("fumi", "fumi_train_n5k5_d512_tanh"), ("fumi", "fumi_train_n20k5_d512"),
         ("fumi", "fumi_test_n5k1_full"), ("maml", "maml_train_n5k5_d512"), ("maml", "maml_test_n5k5_d512"),
         ("am3", "am3_test_n10k5_d512")]


@pytest.mark.parametrize("model_name,golden", CASES)
def test_seeded_init_is_bit_identical_to_reference(model_name, golden):
    g, _ = load_golden(golden)
    argv = ["--model", model_name, "--im_emb_dim", str(int(g["bank_im_dim"])), "--text_emb_dim",
            str(int(g["bank_text_dim"])), *str(g["argv"]).split()]
    args = utils.parser().parse_args(argv)
    torch.manual_seed(args.seed)                       # main.py:51 (np / random seeds do not touch the init)
    model = utils.build_model(args, {})
    want = params_of(g)
    sd = model.state_dict()
    assert list(sd.keys()) == list(want.keys()), "state_dict key names / order differ from the reference"
    for k, v in sd.items():
        assert v.dtype == torch.float32
        assert np.array_equal(v.numpy(), want[k]), f"{golden}: {k} differs from the reference's seeded init"
