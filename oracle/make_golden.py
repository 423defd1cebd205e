"""Generate tests/golden/*.npz by running the UNMODIFIED reference through oracle/ref_shims.py.

Run here (the reference cannot travel):  python -m oracle.make_golden
The reference ships no golden vectors; these fixtures are outputs of the
reference itself: its loader (sampled image ids / labels) and FUMI.evaluate / maml.evaluate /
AM3.evaluate on synthetic banks.  Banks are NOT stored: each fixture records the
fumi_b200.data.synth.make_bank arguments and a checksum, and tests regenerate them.
"""
import copy
import hashlib
import os
import random
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from fumi_b200.data.synth import make_bank  # noqa: E402
from oracle import ref_shims  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def checksum(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def seed_all(s):
    torch.manual_seed(s)
    np.random.seed(s)
    random.seed(s)


def flat(batch):
    (sid, stext, sx), sy = batch["train"]
    (qid, qtext, qx), qy = batch["test"]
    return dict(sup_ids=sid.numpy(), sup_y=sy.numpy(), qry_ids=qid.numpy(), qry_y=qy.numpy())


def setup(bank_kw, argv):
    bank = make_bank(**bank_kw)
    d = tempfile.mkdtemp(prefix="fumi_gold_")
    ref = ref_shims.load_reference()
    model_name = "resnet-34" if bank_kw["im_dim"] == 512 else "resnet-152"
    ref_shims.build_dataset_dir(d, bank, model_name)
    args = ref_shims.make_args(ref, d, ["--image_embedding_model", model_name,
                                        "--im_emb_dim", str(bank_kw["im_dim"]),
                                        "--text_emb_dim", str(bank_kw["text_dim"]), *argv])
    loaders = ref.data.get_dataset(args)            # seeds random/np/torch with 0 three times
    seed_all(args.seed)                             # main.py:51-53
    model = ref.utils.init_model(args, loaders[3], watch=False)
    optim = ref.utils.init_optim(args, model)
    return ref, bank, args, loaders, model, optim


def meta(bank_kw, bank, argv):
    out = {("bank_" + k): np.asarray(v) for k, v in bank_kw.items()}
    out["bank_feats_sha"] = np.asarray(checksum(bank.feats))
    out["bank_text_sha"] = np.asarray(checksum(bank.text))
    out["argv"] = np.asarray(" ".join(argv))
    return out


# ------------------------------------------------------------------------------------- sampler
def golden_sampler():
    """Image ids / labels of consecutive batches, train and test iterators interleaved."""
    # 200 classes -> 120/40/40 split, so 20-way tuples exist in every split
    bank_kw = dict(num_images=200 * 70, num_classes=200, im_dim=16, text_dim=8, min_per_class=60, seed=7)
    for tag, argv, nb in [("n5k5b4", ["--num_ways", "5", "--num_shots", "5", "--batch_size", "4"], 3),
                          ("n5k1b3", ["--num_ways", "5", "--num_shots", "1", "--batch_size", "3"], 3),
                          ("n10k5b2", ["--num_ways", "10", "--num_shots", "5", "--batch_size", "2",
                                       "--num_shots_test", "20"], 2),
                          ("n20k5b2", ["--num_ways", "20", "--num_shots", "5", "--batch_size", "2",
                                       "--num_shots_test", "16"], 2)]:
        ref, bank, args, (tl, vl, te, _), model, optim = setup(bank_kw, ["--model", "maml", *argv])
        out = meta(bank_kw, bank, argv)
        # stream order: val iterator + 1 batch, train iterator + nb batches interleaved with test
        order = []
        it_v = iter(vl); order.append(("val", flat(next(it_v))))
        it_t = iter(tl); it_e = iter(te)
        for _ in range(nb):
            order.append(("train", flat(next(it_t))))
            order.append(("test", flat(next(it_e))))
        out["order"] = np.asarray([o[0] for o in order])
        for i, (_, f) in enumerate(order):
            for k, v in f.items():
                out[f"b{i}_{k}"] = v
        np.savez_compressed(os.path.join(GOLD, f"sampler_{tag}.npz"), **out)
        print("sampler", tag, [o[0] for o in order])


# ------------------------------------------------------------------------------------- episodes
def _capture_adapted(model):
    """Record (im_params, hyper_params) of every im_forward call; last per task = query call."""
    calls = []
    orig = model.im_forward

    def wrapped(im_embeds, im_params, hyper_params):
        calls.append((im_embeds.shape[0], {k: v.detach().clone() for k, v in im_params.items()},
                      hyper_params.detach().clone()))
        return orig(im_embeds, im_params, hyper_params)
    model.im_forward = wrapped
    return calls


def _capture_logits(ref_mod):
    """Record the logits every get_accuracy() call sees (fumi.py:185 / maml.py:183)."""
    rec = []
    orig = ref_mod.get_accuracy

    def wrapped(logits, targets):
        rec.append(logits.detach().clone().numpy())
        return orig(logits, targets)
    ref_mod.get_accuracy = wrapped
    return rec, (lambda: setattr(ref_mod, "get_accuracy", orig))


def golden_fumi(tag, bank_kw, argv, task, steps_flag, store_params=True, n_batches=1):
    ref, bank, args, (tl, vl, te, _), model, optim = setup(bank_kw, ["--model", "fumi", *argv])
    out = meta(bank_kw, bank, argv)
    loader = tl if task == "train" else te
    it = iter(loader)
    for _ in range(n_batches):
        batch = next(it)
    out.update(flat(batch))
    p0 = {k: v.detach().clone().numpy() for k, v in model.state_dict().items()}
    if store_params:
        for k, v in p0.items():
            out["param:" + k] = v
    else:
        for k, v in p0.items():
            out["param_sha:" + k] = np.asarray(checksum(v))
    calls = _capture_adapted(model)
    logits, restore = _capture_logits(ref.fumi)
    loss, acc, preds, targets = model.evaluate(args, batch, optim, task=task)
    restore()
    B, NK = batch["train"][1].shape
    steps = getattr(args, steps_flag)
    per_task = steps + 1
    assert len(calls) == B * per_task
    out["loss"], out["acc"] = np.asarray(loss), np.asarray(acc)
    out["preds"] = preds.detach().numpy().astype(np.int64)
    out["logits"] = np.stack(logits)
    out["hp0"] = np.stack([calls[b * per_task][2].numpy() for b in range(B)]) if steps > 0 else np.zeros(0)
    q = [calls[b * per_task + steps] for b in range(B)]
    out["hp_adapted"] = np.stack([c[2].numpy() for c in q])
    out["W1_adapted"] = np.stack([c[1]["linear1.weight"].numpy() for c in q])
    out["b1_adapted"] = np.stack([c[1]["linear1.bias"].numpy() for c in q])
    out["b0_adapted"] = np.stack([c[1]["linear0.bias"].numpy() for c in q])
    # W0 is 2 MB per task at D=2048: keep task 0 only, rows 0..15
    out["W0_adapted_t0_rows16"] = q[0][1]["linear0.weight"].numpy()[:16]
    if task == "train":
        for k, p in model.named_parameters():
            out["grad:" + k] = p.grad.detach().numpy()
        for k, v in model.state_dict().items():
            out["post:" + k] = v.detach().numpy()
        out["lr"], out["wd"] = np.asarray(args.lr), np.asarray(args.weight_decay)
    out["alpha"], out["steps"] = np.asarray(args.step_size), np.asarray(steps)
    np.savez_compressed(os.path.join(GOLD, f"{tag}.npz"), **out)
    print(tag, "loss", loss, "acc", acc, "size", os.path.getsize(os.path.join(GOLD, f"{tag}.npz")) >> 10, "KiB")


def golden_maml(tag, bank_kw, argv, task, steps_flag):
    ref, bank, args, (tl, vl, te, _), model, optim = setup(bank_kw, ["--model", "maml", *argv])
    out = meta(bank_kw, bank, argv)
    batch = next(iter(tl if task == "train" else te))
    out.update(flat(batch))
    for k, v in model.state_dict().items():
        out["param:" + k] = v.detach().clone().numpy()
    logits, restore = _capture_logits(ref.maml)
    loss, acc = ref.maml.evaluate(args, model, batch, optim, task=task)
    restore()
    out["loss"], out["acc"] = np.asarray(loss), np.asarray(acc)
    out["logits"] = np.stack(logits)
    out["preds"] = out["logits"].argmax(-1).astype(np.int64)
    if task == "train":
        for k, p in model.named_parameters():
            out["grad:" + k] = p.grad.detach().numpy()
        for k, v in model.state_dict().items():
            out["post:" + k] = v.detach().numpy()
        out["lr"], out["wd"] = np.asarray(args.lr), np.asarray(args.weight_decay)
    out["alpha"], out["steps"] = np.asarray(args.step_size), np.asarray(getattr(args, steps_flag))
    out["first_order"] = np.asarray(bool(args.first_order))
    np.savez_compressed(os.path.join(GOLD, f"{tag}.npz"), **out)
    print(tag, "loss", loss, "acc", acc, "size", os.path.getsize(os.path.join(GOLD, f"{tag}.npz")) >> 10, "KiB")


def golden_am3(tag, bank_kw, argv):
    ref, bank, args, (tl, vl, te, _), model, optim = setup(bank_kw, ["--model", "am3", *argv])
    out = meta(bank_kw, bank, argv)
    batch = next(iter(te))
    out.update(flat(batch))
    for k, v in model.state_dict().items():
        out["param:" + k] = v.detach().clone().numpy()
    with torch.no_grad():
        r = model.evaluate(batch=batch, optimizer=None, scheduler=None, num_ways=args.num_ways,
                           device=args.device, task="test")
    loss, acc, f1, prec, rec, lam, preds, trues, qidx, sidx, slam = r
    out.update(loss=np.asarray(loss), acc=np.asarray(acc), f1=np.asarray(f1), prec=np.asarray(prec),
               rec=np.asarray(rec), avg_lamda=np.asarray(lam), preds=np.asarray(preds, np.int64),
               sup_lamda=np.asarray(slam))
    np.savez_compressed(os.path.join(GOLD, f"{tag}.npz"), **out)
    print(tag, "loss", loss, "acc", acc, "size", os.path.getsize(os.path.join(GOLD, f"{tag}.npz")) >> 10, "KiB")


def golden_am3_train(tag, bank_kw, argv):
    """One AM3 meta-train step (am3.py:154-196) at --dropout 0: loss, metrics, all gradients, post-Adam parameters."""
    ref, bank, args, (tl, vl, te, _), model, optim = setup(bank_kw, ["--model", "am3", *argv])
    out = meta(bank_kw, bank, argv)
    batch = next(iter(tl))
    out.update(flat(batch))
    for k, v in model.state_dict().items():
        out["param:" + k] = v.detach().clone().numpy()
    r = model.evaluate(batch=batch, optimizer=optim, scheduler=None, num_ways=args.num_ways, device=args.device,
                       task="train")
    loss, acc, f1, prec, rec, lam = r
    out.update(loss=np.asarray(loss), acc=np.asarray(acc), f1=np.asarray(f1), prec=np.asarray(prec), rec=np.asarray(rec),
               avg_lamda=np.asarray(lam))
    for k, p in model.named_parameters():
        if p.grad is not None:
            out["grad:" + k] = p.grad.detach().numpy()
    for k, v in model.state_dict().items():
        out["post:" + k] = v.detach().numpy()
    out["lr"], out["wd"] = np.asarray(args.lr), np.asarray(args.weight_decay)
    np.savez_compressed(os.path.join(GOLD, f"{tag}.npz"), **out)
    print(tag, "loss", loss, "acc", acc, "size", os.path.getsize(os.path.join(GOLD, f"{tag}.npz")) >> 10, "KiB")


def main():
    os.makedirs(GOLD, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == "am3_train":       # only the fixture added in round 2
        small = dict(num_images=120 * 70, num_classes=120, im_dim=512, text_dim=64, min_per_class=60, seed=11)
        golden_am3_train("am3_train_n10k5_d512", small,
                         ["--num_ways", "10", "--num_shots", "5", "--num_shots_test", "6", "--batch_size", "3",
                          "--dropout", "0"])
        return
    golden_sampler()
    small = dict(num_images=120 * 70, num_classes=120, im_dim=512, text_dim=64, min_per_class=60, seed=11)
    full = dict(num_images=60 * 70, num_classes=60, im_dim=2048, text_dim=768, min_per_class=60, seed=2022)
    # config 2 shape (5-way 5-shot meta-train, 5 steps), reduced D/T so params+grads fit a fixture
    golden_fumi("fumi_train_n5k5_d512", small,
                ["--num_ways", "5", "--num_shots", "5", "--num_shots_test", "8", "--batch_size", "3",
                 "--dropout", "0"], "train", "num_train_adapt_steps")
    golden_fumi("fumi_train_n5k5_d512_tanh", small,
                ["--num_ways", "5", "--num_shots", "5", "--num_shots_test", "8", "--batch_size", "2",
                 "--dropout", "0", "--norm_hypernet", "--lr", "1e-3", "--step_size", "0.05"],
                "train", "num_train_adapt_steps")
    # config 5 shape (20-way 5-shot, 10 steps)
    golden_fumi("fumi_train_n20k5_d512", small,
                ["--num_ways", "20", "--num_shots", "5", "--num_shots_test", "6", "--batch_size", "2",
                 "--dropout", "0", "--num_train_adapt_steps", "10"], "train", "num_train_adapt_steps")
    # config 1 (5-way 1-shot meta-test, 100 steps) and the headline 5-way 5-shot meta-test, full dims
    golden_fumi("fumi_test_n5k1_full", full,
                ["--num_ways", "5", "--num_shots", "1", "--batch_size", "2"], "test",
                "num_test_adapt_steps", store_params=True)
    golden_fumi("fumi_test_n5k5_full", full,
                ["--num_ways", "5", "--num_shots", "5", "--batch_size", "2"], "test",
                "num_test_adapt_steps", store_params=False)
    # config 3 (MAML, same kernels), second- and first-order
    golden_maml("maml_train_n5k5_d512", small,
                ["--num_ways", "5", "--num_shots", "5", "--num_shots_test", "8", "--batch_size", "3"],
                "train", "num_train_adapt_steps")
    golden_maml("maml_train_n5k5_d512_fo", small,
                ["--num_ways", "5", "--num_shots", "5", "--num_shots_test", "8", "--batch_size", "3",
                 "--first_order"], "train", "num_train_adapt_steps")
    golden_maml("maml_test_n5k5_d512", small,
                ["--num_ways", "5", "--num_shots", "5", "--batch_size", "2", "--num_test_adapt_steps", "20"],
                "test", "num_test_adapt_steps")
    # config 4 (AM3 10-way 5-shot meta-test)
    golden_am3("am3_test_n10k5_d512", small,
               ["--num_ways", "10", "--num_shots", "5", "--batch_size", "3"])
    golden_am3_train("am3_train_n10k5_d512", small,
                     ["--num_ways", "10", "--num_shots", "5", "--num_shots_test", "6", "--batch_size", "3",
                      "--dropout", "0"])


if __name__ == "__main__":
    main()
