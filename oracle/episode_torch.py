"""Autograd restatement of the reference's per-task Python loop (TEST INFRASTRUCTURE).

This is the *port* of the reference's CPU path used as ``cpu_baseline`` by bench.py and as a
second, independent check of oracle/episode_np.py: the same torch ops in the same order as
  fumi/models/fumi.py:148-193   (for task in batch: hypernet -> n_steps x [forward, CE,
                                 autograd.grad(create_graph=True), out-of-place SGD] -> query CE)
  fumi/models/maml.py:158-191
with torchmeta's MetaLinear / gradient_update_parameters written inline (F.linear with an
explicit parameter dict; ``p - step_size * grad``).  ``dropout_p`` > 0 adds the reference's
Dropout layers after both ReLUs of im_net in train mode (fumi.py:93-99: torch Bernoulli masks from the
global generator -- used by the timing baseline; parity is defined at --dropout 0 / eval mode or
with injected masks).
"""
from collections import OrderedDict

import torch
import torch.nn.functional as F


def _im_forward(x, im, hp, dropout_p=0.0):
    h = F.relu(F.linear(x, im["linear0.weight"], im["linear0.bias"]))
    if dropout_p > 0:
        h = F.dropout(h, dropout_p, training=True)
    h = F.relu(F.linear(h, im["linear1.weight"], im["linear1.bias"]))
    if dropout_p > 0:
        h = F.dropout(h, dropout_p, training=True)
    out = torch.matmul(h, torch.unsqueeze(hp[:, :-1], 2))            # fumi.py:216
    out = torch.squeeze(out) + torch.unsqueeze(hp[:, -1], 1)         # fumi.py:217
    return torch.transpose(out, 0, 1)


def fumi_batch(params, batch, alpha, steps, tanh=False, train=False, dropout_p=0.0):
    """params: OrderedDict of leaf tensors keyed by reference state_dict names.  batch tensors:
    sup_x [B,NK,D], sup_y [B,NK], qry_x [B,NQ,D], qry_y [B,NQ], class_text [B,N,T].
    If train, leaves get .grad of loss (summed over tasks / B), as fumi.py:187-192."""
    B = batch["sup_x"].shape[0]
    outer = torch.zeros((), dtype=batch["sup_x"].dtype)
    acc = torch.zeros((), dtype=batch["sup_x"].dtype)
    preds, logits = [], []
    for b in range(B):
        u = F.relu(F.linear(batch["class_text"][b], params["hyper_net.0.weight"], params["hyper_net.0.bias"]))
        hp = F.linear(u, params["hyper_net.2.weight"], params["hyper_net.2.bias"])
        if tanh:
            hp = torch.tanh(hp)
        im = OrderedDict((k[len("im_net."):], v) for k, v in params.items() if k.startswith("im_net."))
        for _ in range(steps):
            logit = _im_forward(batch["sup_x"][b], im, hp, dropout_p if train else 0.0)
            inner = F.cross_entropy(logit, batch["sup_y"][b])
            g_hp = torch.autograd.grad(inner, hp, create_graph=True)[0]
            g_im = torch.autograd.grad(inner, list(im.values()), create_graph=True)
            hp = hp - alpha * g_hp
            im = OrderedDict((k, p - alpha * g) for (k, p), g in zip(im.items(), g_im))
        ql = _im_forward(batch["qry_x"][b], im, hp, dropout_p if train else 0.0)
        outer = outer + F.cross_entropy(ql, batch["qry_y"][b])
        p = ql.max(dim=-1)[1]
        preds.append(p)
        logits.append(ql.detach())
        acc = acc + (p == batch["qry_y"][b]).float().mean()
    outer = outer / B
    acc = acc / B
    if train:
        for v in params.values():
            v.grad = None
        outer.backward()
    return dict(loss=outer.detach(), acc=acc.detach(), preds=torch.stack(preds), logits=torch.stack(logits))


def maml_batch(params, batch, alpha, steps, first_order=False, train=False):
    B = batch["sup_x"].shape[0]
    outer = torch.zeros((), dtype=batch["sup_x"].dtype)
    acc = torch.zeros((), dtype=batch["sup_x"].dtype)
    preds, logits = [], []

    def fwd(x, p):
        h = F.relu(F.linear(x, p["net.lin_0.weight"], p["net.lin_0.bias"]))
        h = F.relu(F.linear(h, p["net.lin_1.weight"], p["net.lin_1.bias"]))
        return F.linear(h, p["net.lin_final.weight"], p["net.lin_final.bias"])

    for b in range(B):
        p = OrderedDict(params)
        for _ in range(steps):
            inner = F.cross_entropy(fwd(batch["sup_x"][b], p), batch["sup_y"][b])
            g = torch.autograd.grad(inner, list(p.values()), create_graph=not first_order)
            p = OrderedDict((k, v - alpha * gi) for (k, v), gi in zip(p.items(), g))
        ql = fwd(batch["qry_x"][b], p)
        outer = outer + F.cross_entropy(ql, batch["qry_y"][b])
        pr = ql.max(dim=-1)[1]
        preds.append(pr)
        logits.append(ql.detach())
        acc = acc + (pr == batch["qry_y"][b]).float().mean()
    outer = outer / B
    acc = acc / B
    if train:
        for v in params.values():
            v.grad = None
        outer.backward()
    return dict(loss=outer.detach(), acc=acc.detach(), preds=torch.stack(preds), logits=torch.stack(logits))
