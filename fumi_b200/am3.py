"""AM3 baseline: host-side mirror of the reference's fumi/models/am3.py.

The model definition with reference parameter names and construction order; ``evaluate`` for every task mode --
prototype + text mixing, distances, argmin, CE on the batched kernels (SURVEY.md section 8 row A1 / BASELINE config 4),
and for ``task == "train"`` the hand-written backward (fumi_am3_bwd + the dense-layer gradient kernels), the optimizer
and scheduler steps (am3.py:154-196); ``training_run`` / ``test_loop`` (am3.py:215-367).
"""
import os

import numpy as np
import torch
import torch.nn as nn

from . import engine
from .average_meter import AverageMeter


class AM3(nn.Module):
    """reference: am3.py:16-88."""

    def __init__(self, im_encoder, im_emb_dim, text_encoder, text_emb_dim=300, text_hid_dim=300, prototype_dim=512,
                 dropout=0.7, fine_tune=False, dictionary=None, pooling_strat="mean", lamda_fixed=None):
        super().__init__()
        self.im_emb_dim = im_emb_dim
        self.text_encoder_type = text_encoder
        self.text_emb_dim = text_emb_dim
        self.text_hid_dim = text_hid_dim
        self.prototype_dim = prototype_dim
        self.dropout = dropout
        self.fine_tune = fine_tune
        self.dictionary = dictionary
        self.pooling_strat = pooling_strat
        self.lamda_fixed = lamda_fixed
        if im_encoder in ("precomputed", "resnet"):
            self.image_encoder = nn.Linear(self.im_emb_dim, self.prototype_dim)
        else:
            raise NameError(f"{im_encoder} not allowed as image encoder")
        if text_encoder in ("BERT", "precomputed"):
            self.text_encoder = nn.Identity()
        elif text_encoder == "rand":
            self.text_encoder = nn.Linear(self.text_emb_dim, self.text_emb_dim)
        elif text_encoder in ("w2v", "glove", "RNN", "RNNhid"):
            raise NotImplementedError(f"text encoder {text_encoder!r} needs downloaded word vectors; "
                                      "use precomputed description embeddings (BERT/precomputed)")
        else:
            raise NameError(f"{text_encoder} not allowed as text encoder")
        if not self.fine_tune:
            for p in self.text_encoder.parameters():
                p.requires_grad = False
        self.g = nn.Sequential(nn.Linear(self.text_emb_dim, self.text_hid_dim), nn.ReLU(), nn.Dropout(p=self.dropout),
                               nn.Linear(self.text_hid_dim, self.prototype_dim))
        self.h = nn.Sequential(nn.Linear(self.prototype_dim, self.text_hid_dim), nn.ReLU(), nn.Dropout(p=self.dropout),
                               nn.Linear(self.text_hid_dim, 1))
        self._engine = None

    def _get_engine(self, device):
        if self._engine is None or self._engine.device != torch.device(device):
            self._engine = engine.EpisodeEngine(device)
        return self._engine

    def forward(self, inputs, im_only=False):
        """am3.py:90-126 (plain torch; the hot path goes through evaluate)."""
        idx, text, im = inputs
        im_embeddings = self.image_encoder(im)
        if im_only:
            return im_embeddings
        B, NK, _ = text.shape
        if self.text_encoder_type == "rand":
            text_embeddings = 2 * torch.rand(size=(B, NK, self.prototype_dim), device=im_embeddings.device) - 1
        else:
            text_embeddings = self.g(self.text_encoder(text))
        lamda = torch.sigmoid(self.h(text_embeddings))
        return im_embeddings, text_embeddings, lamda

    def evaluate(self, batch, optimizer, scheduler, num_ways, device, task="train"):
        """am3.py:128-212.  Returns the reference's 11-tuple for task == 'test', 6-tuple for 'val'."""
        train = task == "train"
        self.train(train)
        if self.text_encoder_type == "rand":
            raise NotImplementedError("text_encoder='rand' draws prototypes from the host RNG; not built")
        if train:
            # optimizer.zero_grad() of am3.py:189; FusedAdam keeps its flat gradient views alive
            (optimizer if hasattr(optimizer, "_flat") else self).zero_grad()
        res = self._get_engine(device).am3_batch(self, batch, num_ways, train=train)
        if train:
            optimizer.step()                      # am3.py:190-193
            if scheduler:
                scheduler.step()
        eb = res["batch"]
        B, NQ = eb.qry_y.shape
        if train:
            la = res["loss_acc"].cpu().numpy()        # (mean CE, mean lamda), summed over ranks when sharded
            loss = la[0]
        else:
            loss = (res["task_loss"].sum() / float(B * NQ)).cpu().numpy()      # mean over all queries (utils.py:402)
        acc, prec, rec, f1 = macro_scores(res["confusion"].cpu().numpy())                 # utils.py:323-326
        avg_lamda = la[1] if train else res["sup_lamda"].mean().cpu().numpy()
        if task == "test":
            to_np = lambda t: t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)
            preds = res["preds"].cpu().numpy()
            return (loss, acc, f1, prec, rec, avg_lamda, preds, eb.qry_y, to_np(eb.qry_ids), to_np(eb.sup_ids),
                    res["sup_lamda"].cpu().numpy())
        return loss, acc, f1, prec, rec, avg_lamda


def macro_scores(counts):
    """accuracy_score and precision_recall_fscore_support(average="macro") of sklearn (utils/utils.py:16,323-326) from
    the confusion counts [target, prediction] the device produced (fumi_confusion_counts): same labels (those present
    in the targets or the predictions), same divisions with zero_division -> 0, same unweighted means."""
    counts = np.asarray(counts, dtype=np.int64)
    tp, pred_sum, true_sum = np.diag(counts), counts.sum(0), counts.sum(1)
    present = (pred_sum + true_sum) > 0
    tp, pred_sum, true_sum = tp[present], pred_sum[present], true_sum[present]

    def div(a, b):
        out = np.zeros(len(a), dtype=np.float64)
        np.divide(a, b, out=out, where=b != 0)
        return out

    acc = float(np.diag(counts).sum() / counts.sum())
    prec, rec = div(tp.astype(np.float64), pred_sum.astype(np.float64)), div(tp.astype(np.float64), true_sum.astype(np.float64))
    f1 = div(2.0 * tp.astype(np.float64), true_sum.astype(np.float64) + pred_sum.astype(np.float64))
    return acc, float(np.average(prec)), float(np.average(rec)), float(np.average(f1))


def training_run(args, model, optimizer, train_loader, val_loader, max_test_batches):
    """am3.py:215-305: initial validation, train step per batch, validation + checkpoint every eval_freq batches (batch 0
    included, unlike FuMI's loop), patience, best-checkpoint reload."""
    from . import utils
    best_loss, best_acc = test_loop(args, model, val_loader, max_test_batches)[:2]
    print(f"\ninitial loss: {best_loss}, acc: {best_acc}")
    best_batch_idx = 0
    if type(optimizer) == tuple:
        opt, scheduler = optimizer
    else:
        opt, scheduler = optimizer, None
    saved_best = False
    try:
        for batch_idx, batch in enumerate(train_loader):
            train_loss, train_acc, train_f1, train_prec, train_rec, train_lamda = model.evaluate(
                batch=batch, optimizer=opt, scheduler=scheduler, num_ways=args.num_ways, device=args.device, task="train")
            utils.log({"train/acc": train_acc, "train/f1": train_f1, "train/prec": train_prec, "train/rec": train_rec,
                       "train/loss": train_loss, "train/avg_lamda": train_lamda,
                       "num_episodes": (batch_idx + 1) * args.batch_size}, step=batch_idx)
            if batch_idx % args.eval_freq == 0:
                val = test_loop(args, model, val_loader, max_test_batches)
                val_loss, val_acc, val_f1, val_prec, val_rec, val_lamda = val[:6]
                is_best = val_loss < best_loss
                if is_best:
                    best_loss = val_loss
                    best_batch_idx = batch_idx
                utils.log({"val/acc": val_acc, "val/f1": val_f1, "val/prec": val_prec, "val/rec": val_rec,
                           "val/loss": val_loss, "val/avg_lamda": val_lamda}, step=batch_idx)
                utils.save_checkpoint({"batch_idx": batch_idx, "state_dict": model.state_dict(), "best_loss": best_loss,
                                       "optimizer": opt.state_dict(), "args": utils.args_dict(args)}, is_best, args)
                saved_best |= is_best
                print(f"\nBatch {batch_idx+1}/{args.epochs}: \ntrain/loss: {train_loss}, train/acc: {train_acc}, "
                      f"train/avg_lamda: {train_lamda}\nval/loss: {val_loss}, val/acc: {val_acc}, val/avg_lamda: {val_lamda}")
            if (batch_idx > args.epochs - 1) or (args.patience > 0 and batch_idx - best_batch_idx > args.patience):
                break
    except KeyboardInterrupt:
        pass
    best_file = os.path.join(utils.run_dir(args), "best.pth.tar")
    if saved_best and os.path.exists(best_file):                   # am3.py:302-303 (only a checkpoint of THIS run)
        model, _ = utils.load_checkpoint(model, opt, args.device, best_file)
    return model


def test_loop(args, model, test_dataloader, max_num_batches):
    """am3.py:308-367."""
    meters = [AverageMeter() for _ in range(6)]
    test_preds, test_trues, query_idx, support_idx, support_lamdas = [], [], [], [], []
    for batch_idx, batch in enumerate(test_dataloader):
        (test_loss, test_acc, test_f1, test_prec, test_rec, lamda, preds, trues, query, support,
         support_lamda) = model.evaluate(batch=batch, optimizer=None, scheduler=None, num_ways=args.num_ways,
                                         device=args.device, task="test")
        for mtr, v in zip(meters, (test_acc, test_f1, test_prec, test_rec, test_loss, lamda)):
            mtr.update(v)
        test_preds += preds.tolist()
        test_trues += trues.tolist()
        query_idx += query.tolist()
        support_idx += support.tolist()
        support_lamdas += support_lamda.tolist()
        if batch_idx > max_num_batches - 1:
            break
    acc, f1, prec, rec, loss, lam = [mtr.avg for mtr in meters]
    return loss, acc, f1, prec, rec, lam, test_preds, test_trues, query_idx, support_idx, support_lamdas
