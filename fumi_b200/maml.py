"""MAML baseline: host-side mirror of the reference's fumi/models/maml.py.

Same names / signatures (PureImageNetwork, evaluate, training_run, test_loop); the per-task loop
(maml.py:158-183) runs on the same batched inner-loop kernels as FuMI, with the shared
``lin_final`` parameter as the head initialisation of every task.
"""
from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import engine
from .average_meter import AverageMeter


class PureImageNetwork(nn.Module):
    """reference: maml.py:15-33.  Parameters: net.lin_0.*, net.lin_1.*, net.lin_final.*"""

    def __init__(self, im_embed_dim=2048, n_way=5, hidden_dims=None):
        super().__init__()
        self.im_embed_dim = im_embed_dim
        self.n_way = n_way
        layers = OrderedDict()
        in_dim = im_embed_dim
        if hidden_dims is not None:
            for i, hid_dim in enumerate(hidden_dims):
                layers["lin_" + str(i)] = nn.Linear(in_dim, hid_dim)
                layers["relu_" + str(i)] = nn.ReLU()
                in_dim = hid_dim
        layers["lin_final"] = nn.Linear(in_dim, n_way)
        self.net = nn.Sequential(layers)
        self.hidden_dims = list(hidden_dims) if hidden_dims is not None else []
        self._engine = None

    def forward(self, inputs, params=None):
        """maml.py:31-33: optional explicit parameter dict keyed 'net.<layer>.<weight|bias>'."""
        if params is None:
            return self.net(inputs)
        out = inputs
        for name, mod in self.net.named_children():
            if isinstance(mod, nn.Linear):
                out = F.linear(out, params[f"net.{name}.weight"], params.get(f"net.{name}.bias"))
            else:
                out = mod(out)
        return out

    def _get_engine(self, device):
        if self._engine is None or self._engine.device != torch.device(device):
            self._engine = engine.EpisodeEngine(device)
        return self._engine


def evaluate(args, model, batch, optimizer, task="train"):
    """One meta-batch (maml.py:134-193).  Returns (loss np 0-d f32, acc np 0-d f32)."""
    model.train()                 # maml.py:143 (no dropout layers: no behavioural effect)
    if task == "train":
        # optimizer.zero_grad() of maml.py:188.  Only on the train path: test_loop passes optimizer=None, and a
        # Module.zero_grad() there would drop the gradients out of FusedAdam's flat buffer (the engine writes no
        # gradient in test mode, so there is nothing to clear).
        (optimizer if hasattr(optimizer, "_flat") else model).zero_grad()
    eng = model._get_engine(args.device)
    steps = args.num_train_adapt_steps if task == "train" else args.num_test_adapt_steps
    res = eng.maml_batch(model, batch, steps=steps, step_size=args.step_size, train=(task == "train"),
                         first_order=bool(args.first_order))
    if task == "train":
        optimizer.step()          # maml.py:188-191
    la = eng.read_loss_acc(res)  # (one GPU: waits for the forward's loss only, engine.loss_acc_early)
    evaluate.last = res           # logits / preds of the last call, for inspection and tests
    return la[0], la[1]


def training_run(args, model, optimizer, train_loader, val_loader, max_test_batches):
    """maml.py:36-107 (no best-checkpoint reload at the end, unlike FuMI)."""
    from . import utils
    best_loss, best_acc = test_loop(args, model, val_loader, max_test_batches)
    print(f"\ninitial loss: {best_loss}, acc: {best_acc}")
    best_batch_idx = 0
    try:
        for batch_idx, batch in enumerate(train_loader):
            train_loss, train_acc = evaluate(args=args, model=model, batch=batch, optimizer=optimizer, task="train")
            utils.log({"train/acc": train_acc, "train/loss": train_loss,
                       "num_episodes": (batch_idx + 1) * args.batch_size}, step=batch_idx)
            if batch_idx % args.eval_freq == 0 and batch_idx != 0:
                val_loss, val_acc = test_loop(args, model, val_loader, max_test_batches)
                is_best = val_loss < best_loss
                if is_best:
                    best_loss = val_loss
                    best_batch_idx = batch_idx
                utils.log({"val/acc": val_acc, "val/loss": val_loss}, step=batch_idx)
                utils.save_checkpoint({"batch_idx": batch_idx, "state_dict": model.state_dict(),
                                       "best_loss": best_loss, "optimizer": optimizer.state_dict(),
                                       "args": utils.args_dict(args)}, is_best, args)
                print(f"\nBatch {batch_idx+1}/{args.epochs}: \ntrain/loss: {train_loss}, train/acc: {train_acc}"
                      f"\nval/loss: {val_loss}, val/acc: {val_acc}")
            if (batch_idx > args.epochs - 1) or (args.patience > 0 and batch_idx - best_batch_idx > args.patience):
                break
    except KeyboardInterrupt:
        pass
    return model


def test_loop(args, model, test_loader, max_num_batches):
    """maml.py:110-131."""
    avg_test_acc, avg_test_loss = AverageMeter(), AverageMeter()
    for batch_idx, batch in enumerate(test_loader):
        test_loss, test_acc = evaluate(args=args, model=model, batch=batch, optimizer=None, task="test")
        avg_test_acc.update(test_acc)
        avg_test_loss.update(test_loss)
        if batch_idx > max_num_batches - 1:
            break
    return avg_test_loss.avg, avg_test_acc.avg


def get_accuracy(logits, targets):
    _, predictions = torch.max(logits, dim=-1)
    return torch.mean(predictions.eq(targets).float())
