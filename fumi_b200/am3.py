"""AM3 baseline: host-side mirror of the reference's fumi/models/am3.py for meta-test scoring.

In scope (SURVEY.md section 8, row A1 / BASELINE config 4): the model definition with reference
parameter names and construction order, and ``evaluate(task != "train")`` / ``test_loop`` -- prototype
+ text mixing, distances, argmin, CE -- on the batched kernels.  AM3 *training* is outside the
episodic inner-loop path (section 8(f) rank 3) and raises NotImplementedError.
"""
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
        if task == "train":
            raise NotImplementedError("AM3 meta-training is outside the accelerated episodic inner-loop path "
                                      "(SURVEY.md section 8(f)); only meta-test scoring is built")
        self.eval()
        if self.text_encoder_type == "rand":
            raise NotImplementedError("text_encoder='rand' draws prototypes from the host RNG; not built")
        res = self._get_engine(device).am3_batch(self, batch, num_ways)
        eb = res["batch"]
        B, NQ = eb.qry_y.shape
        loss = (res["task_loss"].sum() / float(B * NQ)).cpu().numpy()          # mean over all queries (utils.py:402)
        preds = res["preds"].cpu().numpy()
        flat_preds, flat_targets = preds.reshape(-1), eb.qry_y.cpu().numpy().reshape(-1)
        from sklearn.metrics import accuracy_score, precision_recall_fscore_support    # utils.py:16,323-326
        acc = accuracy_score(flat_targets, flat_preds)
        prec, rec, f1, _ = precision_recall_fscore_support(flat_targets, flat_preds, average="macro")
        avg_lamda = res["sup_lamda"].mean().cpu().numpy()
        if task == "test":
            to_np = lambda t: t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)
            return (loss, acc, f1, prec, rec, avg_lamda, preds, eb.qry_y, to_np(eb.qry_ids), to_np(eb.sup_ids),
                    res["sup_lamda"].cpu().numpy())
        return loss, acc, f1, prec, rec, avg_lamda


def training_run(args, model, optimizer, train_loader, val_loader, max_test_batches):
    raise NotImplementedError("AM3 meta-training is outside the accelerated path (SURVEY.md section 8(f))")


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
