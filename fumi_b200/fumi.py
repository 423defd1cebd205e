"""FuMI model + loops: host-side mirror of the reference's fumi/models/fumi.py.

Same class / function names, constructor arguments, parameter names and registration order
(so ``torch.manual_seed(s)`` gives bit-identical initial weights and reference checkpoints
load), same ``evaluate`` signature and return convention.  The per-task Python loop of the
reference (fumi.py:148-185) is replaced by the batched sm_100a kernels behind the C ABI
(include/fumi_b200.h) via fumi_b200.engine; there is no CPU path.
"""
import math
import os
from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import engine
from .average_meter import AverageMeter

try:                                    # logging only; never on the compute path
    import wandb                        # reference: fumi.py:1,248-254
except Exception:                       # pragma: no cover
    wandb = None


class FUMI(nn.Module):
    """reference: fumi/models/fumi.py:18-107 (constructor), 109-113 (forward)."""

    def __init__(self, n_way=5, im_emb_dim=2048, im_hid_dim=[64], text_encoder="BERT", text_emb_dim=300,
                 text_hid_dim=1024, dropout_rate=0.0, dictionary=None, pooling_strat="mean",
                 init_all_layers=False, norm_hypernet=True, fine_tune=False, init_bias=False):
        super().__init__()
        self.n_way = n_way
        self.im_emb_dim = im_emb_dim
        self.im_hid_dim = list(im_hid_dim)
        self.text_encoder_type = text_encoder
        self.text_emb_dim = text_emb_dim
        self.text_hid_dim = text_hid_dim
        self.dropout_rate = dropout_rate
        self.dictionary = dictionary
        self.pooling_strat = pooling_strat
        self.norm_hypernet = norm_hypernet
        self.fine_tune = fine_tune
        self.init_bias = init_bias

        if text_encoder in ("BERT", "precomputed"):
            self.text_encoder = nn.Identity()                 # embeddings precomputed (fumi.py:47-49)
        elif text_encoder == "rand":
            self.text_encoder = nn.Linear(text_emb_dim, text_emb_dim)
        elif text_encoder in ("w2v", "glove", "RNN", "RNNhid"):
            # fumi/models/common.py needs gensim downloads; out of scope
            raise NotImplementedError(f"text encoder {text_encoder!r} needs downloaded word vectors; "
                                      "use precomputed description embeddings (BERT/precomputed)")
        else:
            raise NameError(f"{text_encoder} not allowed as text encoder")
        if not self.fine_tune:
            for p in self.text_encoder.parameters():
                p.requires_grad = False

        hyper_net_layers = [nn.Linear(self.text_emb_dim, self.text_hid_dim), nn.ReLU()]
        self.init_all_layers = init_all_layers
        if init_all_layers:
            raise NotImplementedError("Entire model hypernet initialisation removed")   # fumi.py:102
        head = nn.Linear(self.text_hid_dim, self.im_hid_dim[-1] + 1)                    # weights + bias
        if self.init_bias:
            # hyper_weight_layer_init('relu','normc',...,adjust_weights=False, adjust_bias=True)
            # (fumi.py:81-84, utils/hypernet_init.py:137-167,88-117,23-25): zero weight, normc bias
            with torch.no_grad():
                head.weight.zero_()
                b = head.bias.view(1, -1)
                b.normal_(0, 1)
                b *= math.sqrt(2.0) / torch.sqrt(b.pow(2).sum(1, keepdim=True))
        hyper_net_layers.append(head)
        im_net_layers = OrderedDict()
        if len(self.im_hid_dim) > 0:
            im_net_layers["linear0"] = nn.Linear(self.im_emb_dim, self.im_hid_dim[0])
            im_net_layers["relu0"] = nn.ReLU()
            if self.dropout_rate > 0:
                im_net_layers["dropout0"] = nn.Dropout(self.dropout_rate)
            for i in range(len(self.im_hid_dim) - 1):
                im_net_layers["linear" + str(i + 1)] = nn.Linear(self.im_hid_dim[i], self.im_hid_dim[i + 1])
                im_net_layers["relu" + str(i + 1)] = nn.ReLU()
                if self.dropout_rate > 0:
                    im_net_layers["dropout" + str(i + 1)] = nn.Dropout(self.dropout_rate)
        self.im_net = nn.Sequential(im_net_layers)
        if self.norm_hypernet:
            hyper_net_layers.append(nn.Tanh())
        self.hyper_net = nn.Sequential(*hyper_net_layers)
        self._engine = None
        self.dropout_seed = 0          # counter-based dropout masks: advanced once per train batch

    # -- reference API -------------------------------------------------------------------------
    def forward(self, text_embed):
        """Hyper-network forward pass (text -> image params), fumi.py:109-113.  On a CUDA tensor this
        runs the library's dense-layer kernels; kept differentiable-free (use evaluate to train)."""
        if text_embed.is_cuda:
            return self._get_engine(text_embed.device).hypernet(self, text_embed.reshape(-1, text_embed.shape[-1])
                                                                ).reshape(*text_embed.shape[:-1], -1)
        return self.hyper_net(text_embed)

    def get_hyper_params(self, text, targets, device, attn_mask=None):
        """fumi.py:198-212: text row of the first support sample of each label -> hypernet."""
        if self.text_encoder_type == "rand":
            text_encoding = 2 * torch.rand(text.shape[0], self.text_emb_dim, device=text.device) - 1
        else:
            text_encoding = self.text_encoder(text.unsqueeze(0)).squeeze(0)
        first = engine.first_row_of_each_label(targets.unsqueeze(0), self.n_way)[0]
        return self(text_encoding[first].to(device))

    def im_forward(self, im_embeds, im_params, hyper_params):
        """fumi.py:214-218 (plain torch; the hot path does not come through here)."""
        out = im_embeds
        for name, mod in self.im_net.named_children():
            if isinstance(mod, nn.Linear):
                out = F.linear(out, im_params[name + ".weight"], im_params[name + ".bias"])
            else:
                out = mod(out)
        return out @ hyper_params[:, :-1].t() + hyper_params[:, -1]

    def _get_engine(self, device):
        if self._engine is None or self._engine.device != torch.device(device):
            self._engine = engine.EpisodeEngine(device)
        return self._engine

    def evaluate(self, args, batch, optimizer, task="train"):
        """One meta-batch (fumi.py:115-196).  ``batch`` is either the reference's torchmeta dict
        ({'train': [[ids, text, im], targets], 'test': ...}) or a fumi_b200.data.EpisodeBatch of
        indices into a GPU-resident FeatureBank.

        Returns (loss np 0-d f32, acc np 0-d f32, test_preds f32 [B,NQ] on device, test_targets i64)."""
        if task == "train":
            self.train()
            # optimizer.zero_grad() of the reference (fumi.py:190); FusedAdam keeps its flat gradient views alive (one
            # memset), Module.zero_grad() would drop them and fall back to per-tensor Adam / all-reduce launches
            (optimizer if hasattr(optimizer, "_flat") else self).zero_grad()
        else:
            self.eval()
        eng = self._get_engine(args.device)
        steps = args.num_train_adapt_steps if task == "train" else args.num_test_adapt_steps
        res = eng.fumi_batch(self, batch, steps=steps, step_size=args.step_size, train=(task == "train"))
        if task == "train":
            # optimizer.zero_grad(); outer_loss.backward(); optimizer.step()   (fumi.py:190-193)
            optimizer.step()
        # the single device sync per batch, as the reference's .cpu().numpy() (fumi.py:195-196); on one GPU it waits
        # for the forward's loss only (engine.loss_acc_early), the backward and the Adam step keep running behind it
        la = eng.read_loss_acc(res)
        return la[0], la[1], res["preds"].to(torch.float32), res["qry_y"]


def training_run(args, model, optimizer, train_loader, val_loader, max_test_batches):
    """FUMI training loop (fumi.py:220-299): initial validation pass, per-batch train step,
    validation + checkpoint every eval_freq batches, patience, best-checkpoint reload."""
    from . import utils
    best_loss, best_acc, _, _ = test_loop(args, model, val_loader, max_test_batches)
    print(f"\ninitial loss: {best_loss}, acc: {best_acc}")
    best_batch_idx = 0
    if type(optimizer) == tuple:
        opt, scheduler = optimizer
    else:
        opt, scheduler = optimizer, None
    saved_best = False
    try:
        for batch_idx, batch in enumerate(train_loader):
            train_loss, train_acc, _, _ = model.evaluate(args=args, batch=batch, optimizer=opt, task="train")
            utils.log({"train/acc": train_acc, "train/loss": train_loss,
                       "num_episodes": (batch_idx + 1) * args.batch_size}, step=batch_idx)
            if batch_idx % args.eval_freq == 0 and batch_idx != 0:
                val_loss, val_acc, _, _ = test_loop(args, model, val_loader, max_test_batches)
                is_best = val_loss < best_loss
                if is_best:
                    best_loss = val_loss
                    best_batch_idx = batch_idx
                utils.log({"val/acc": val_acc, "val/loss": val_loss}, step=batch_idx)
                utils.save_checkpoint({"batch_idx": batch_idx, "state_dict": model.state_dict(),
                                       "best_loss": best_loss, "optimizer": opt.state_dict(),
                                       "args": utils.args_dict(args)}, is_best, args)
                saved_best |= is_best
                print(f"\nBatch {batch_idx+1}/{args.epochs}: \ntrain/loss: {train_loss}, train/acc: {train_acc}"
                      f"\nval/loss: {val_loss}, val/acc: {val_acc}")
            # break after max iters or early stopping (off-by-one kept: fumi.py:288-291)
            if (batch_idx > args.epochs - 1) or (args.patience > 0 and batch_idx - best_batch_idx > args.patience):
                break
    except KeyboardInterrupt:
        pass
    best_file = os.path.join(utils.run_dir(args), "best.pth.tar")
    if saved_best and os.path.exists(best_file):                   # fumi.py:296-297 (only a checkpoint of THIS run)
        model, _ = utils.load_checkpoint(model, opt, args.device, best_file)
    if saved_best and torch.distributed.is_available() and torch.distributed.is_initialized():
        for t in list(model.parameters()) + list(model.buffers()):   # rank 0 holds the checkpoint files
            torch.distributed.broadcast(t.data, 0)
    return model


def test_loop(args, model, test_loader, max_num_batches):
    """fumi.py:302-326; runs max_num_batches + 1 batches like the reference (break after processing)."""
    avg_test_acc, avg_test_loss = AverageMeter(), AverageMeter()
    test_preds, test_targets = [], []
    for batch_idx, batch in enumerate(test_loader):
        test_loss, test_acc, preds, target = model.evaluate(args=args, batch=batch, optimizer=None, task="test")
        avg_test_acc.update(test_acc)
        avg_test_loss.update(test_loss)
        test_preds.append(preds)
        test_targets.append(target)
        if batch_idx > max_num_batches - 1:
            break
    return avg_test_loss.avg, avg_test_acc.avg, test_preds, test_targets


def get_accuracy(logits, targets):
    _, predictions = torch.max(logits, dim=-1)
    return torch.mean(predictions.eq(targets).float())
