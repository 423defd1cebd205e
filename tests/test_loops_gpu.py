"""Reference loop/driver semantics on the GPU path: main.py flow on a small synthetic
bank -- get_dataset, seed order, training_run (initial validation pass, off-by-one batch counts, periodic
validation + checkpoint, best-checkpoint reload), test_loop, checkpoint round trip."""
import os
import random
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def make_args(tmp, model="fumi", **kw):
    from fumi_b200 import utils
    argv = ["--model", model, "--synthetic", "--wandb_offline", "--batch_size", "6", "--num_shots", "2",
            "--num_shots_test", "4", "--epochs", "3", "--eval_freq", "2", "--num_ep_test", "24", "--im_emb_dim", "64",
            "--text_emb_dim", "16", "--image_embedding_model", "resnet-152", "--log_dir", str(tmp),
            "--num_test_adapt_steps", "3", "--dropout", "0.25"]
    a = utils.parser().parse_args(argv)
    a.device = torch.device("cuda", 0)
    for k, v in kw.items():
        setattr(a, k, v)
    return a


@pytest.fixture()
def small_bank(monkeypatch):
    monkeypatch.setenv("FUMI_SYNTH_IMAGES", str(60 * 70))
    monkeypatch.setenv("FUMI_SYNTH_CLASSES", "60")


class CountingLoader:
    def __init__(self, loader):
        self.loader, self.batches, self.iters = loader, 0, 0

    def __iter__(self):
        self.iters += 1
        for b in self.loader:
            self.batches += 1
            yield b


def test_fumi_training_run_and_test_loop_follow_the_reference_counts(tmp_path, small_bank):
    from fumi_b200 import fumi, utils
    from fumi_b200.data.loader import get_dataset
    args = make_args(tmp_path)
    tl, vl, te, dictionary = get_dataset(args)
    torch.manual_seed(args.seed); np.random.seed(args.seed); random.seed(args.seed)      # main.py:51-53
    model = utils.init_model(args, dictionary)
    opt = utils.init_optim(args, model)
    p0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    tl, vl = CountingLoader(tl), CountingLoader(vl)
    max_test_batches = int(args.num_ep_test / args.batch_size)                            # 4
    fumi.training_run(args, model, opt, tl, vl, max_test_batches // 2)                    # main.py:84-85
    # initial validation + one validation at batch_idx 2: each runs max_num_batches + 1 batches (fumi.py:324)
    assert vl.iters == 2 and vl.batches == 2 * (max_test_batches // 2 + 1)
    # training: breaks when batch_idx > epochs - 1, i.e. after epochs + 1 batches (fumi.py:288)
    assert tl.batches == args.epochs + 1
    assert os.path.exists(os.path.join(utils.run_dir(args), "ckpt.pth.tar"))
    moved = max((model.state_dict()[k] - p0[k]).abs().max().item() for k in p0)
    assert 0 < moved <= 5 * args.lr * 1.01                                               # Adam: <= lr per step
    te = CountingLoader(te)
    loss, acc, preds, targets = fumi.test_loop(args, model, te, max_test_batches)
    assert te.batches == max_test_batches + 1 and len(preds) == max_test_batches + 1
    assert np.isfinite(loss) and 0.0 <= acc <= 1.0
    assert preds[0].shape == targets[0].shape == (args.batch_size, 5 * 20) and preds[0].dtype == torch.float32


def test_checkpoint_round_trip_keeps_the_reference_schema(tmp_path, small_bank):
    from fumi_b200 import utils
    from fumi_b200.data.loader import get_dataset
    args = make_args(tmp_path)
    tl, vl, te, dictionary = get_dataset(args)
    torch.manual_seed(1)
    model = utils.init_model(args, dictionary)
    opt = utils.init_optim(args, model)
    batch = next(iter(tl))
    model.evaluate(args, batch, opt, task="train")
    ck = {"batch_idx": 0, "state_dict": model.state_dict(), "best_loss": 1.0, "optimizer": opt.state_dict(),
          "args": utils.args_dict(args)}
    utils.save_checkpoint(ck, True, args)
    saved = torch.load(os.path.join(utils.run_dir(args), "best.pth.tar"), weights_only=False)
    assert set(saved) == {"batch_idx", "state_dict", "best_loss", "optimizer", "args"}          # utils.py:271-277
    assert list(saved["state_dict"]) == ["im_net.linear0.weight", "im_net.linear0.bias", "im_net.linear1.weight",
                                         "im_net.linear1.bias", "hyper_net.0.weight", "hyper_net.0.bias",
                                         "hyper_net.2.weight", "hyper_net.2.bias"]
    want = {k: v.clone() for k, v in model.state_dict().items()}
    model.evaluate(args, next(iter(tl)), opt, task="train")                                   # move away
    utils.load_checkpoint(model, opt, args.device, os.path.join(utils.run_dir(args), "best.pth.tar"))
    for k, v in model.state_dict().items():
        assert torch.equal(v, want[k]), k
    model.evaluate(args, next(iter(tl)), opt, task="train")                                   # flat views still live


def test_maml_and_am3_loops(tmp_path, small_bank):
    from fumi_b200 import am3, maml, utils
    from fumi_b200.data.loader import get_dataset
    args = make_args(tmp_path, model="maml")
    tl, vl, te, d = get_dataset(args)
    torch.manual_seed(args.seed); random.seed(args.seed)
    model = utils.init_model(args, d)
    opt = utils.init_optim(args, model)
    maml.training_run(args, model, opt, tl, vl, 2)
    loss, acc = maml.test_loop(args, model, te, 2)
    assert np.isfinite(loss) and 0.0 <= acc <= 1.0
    args = make_args(tmp_path, model="am3")
    model = utils.init_model(args, d)
    out = am3.test_loop(args, model, te, 2)
    assert len(out) == 11 and np.isfinite(out[0]) and len(out[6]) == 3 * args.batch_size


def test_unsupported_configurations_fail_loudly(tmp_path, small_bank):
    from fumi_b200 import utils
    from fumi_b200.data.loader import get_dataset
    args = make_args(tmp_path)
    args.im_hid_dim = [128, 64]
    tl, _, _, d = get_dataset(args)
    model = utils.init_model(args, d)
    with pytest.raises(NotImplementedError, match="im_hid_dim"):
        model.evaluate(args, next(iter(tl)), utils.init_optim(args, model), task="train")
    with pytest.raises(NotImplementedError):
        utils.init_model(make_args(tmp_path, init_all_layers=True), d)
    with pytest.raises(NotImplementedError):
        utils.init_optim(make_args(tmp_path, optim="rmsprop"), model)
