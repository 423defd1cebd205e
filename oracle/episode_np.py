"""Analytic numpy restatement of the FuMI / MAML episode (TEST INFRASTRUCTURE).

Follows, line by line:
  fumi/models/fumi.py:109-113,198-212  hypernetwork  text row of label i -> hp[i] (64 w + 1 b)
  fumi/models/fumi.py:214-218          im_forward    MLP(X; W0,b0,W1,b1) . hp[:, :-1]^T + hp[:, -1]
  fumi/models/fumi.py:160-176          inner step    CE, grad wrt hp and im_net params at the
                                                     pre-update point, SGD update (out-of-place)
  fumi/models/fumi.py:178-193          query scoring, argmax (first max), loss/B, backward
  fumi/models/maml.py:134-193          same loop with the shared lin_final head, --first_order
  torchmeta gradient_update_parameters (restated in torchmeta_shim.py)
  fumi/utils/utils.py:280-283          Adam(lr, weight_decay as L2 in the gradient)

Everything is written out by hand (softmax-CE gradient, ReLU/dropout masks as constants, the
exact second-order meta-gradient = reverse sweep through the unrolled SGD steps with
Hessian-vector products), in the *direct* form (per-task W0 materialised).  The product's
CUDA kernels use the algebraically identical Gram form; this file is their checker.

Dropout: the reference draws Bernoulli masks from the torch global generator
(SURVEY.md B.6); here masks are *injected* (``masks`` argument; entries 0 or 1/(1-p)), one
per forward (each inner step and the query pass), so a counter-based device RNG can be checked.
"""
import numpy as np


def _softmax(L):
    m = L.max(axis=1, keepdims=True)
    e = np.exp(L - m)
    return e / e.sum(axis=1, keepdims=True)


def _ce_mean(L, y):
    m = L.max(axis=1, keepdims=True)
    lse = m[:, 0] + np.log(np.exp(L - m).sum(axis=1))
    return float(np.mean(lse - L[np.arange(len(y)), y]))


def hypernet_forward(hyper, text, tanh=False):
    """hyper = (Wh1[Th,T], bh1[Th], Wh2[65,Th], bh2[65]); text [R,T] -> (u [R,Th], hp [R,65])."""
    Wh1, bh1, Wh2, bh2 = hyper
    pre = text @ Wh1.T + bh1
    u = np.maximum(pre, 0)
    hp = u @ Wh2.T + bh2
    if tanh:
        hp = np.tanh(hp)
    return u, hp


def hypernet_backward(hyper, text, u, hp, dhp, tanh=False):
    Wh1, bh1, Wh2, bh2 = hyper
    if tanh:
        dhp = dhp * (1.0 - hp * hp)
    dWh2 = dhp.T @ u
    dbh2 = dhp.sum(0)
    du = (dhp @ Wh2) * (u > 0)
    dWh1 = du.T @ text
    dbh1 = du.sum(0)
    return dWh1, dbh1, dWh2, dbh2


def _mlp(X, th, m0, m1, force=None, ties=None):
    """force = (R0 bool [n,H0], R1 bool [n,H1]) overrides the ReLU gates (used to compare against an
    implementation whose rounding resolved a near-zero pre-activation the other way); every
    overridden entry's |pre-activation| is appended to `ties`."""
    W0, b0, W1, b1, Wh, bh = th
    Z0 = X @ W0.T + b0
    R0 = Z0 > 0
    if force is not None:
        if ties is not None:
            ties.extend(np.abs(Z0[R0 != force[0]]).tolist())
        R0 = force[0]
    M0 = R0.astype(X.dtype)
    if m0 is not None:
        M0 = M0 * m0
    H0 = Z0 * M0
    Z1 = H0 @ W1.T + b1
    R1 = Z1 > 0
    if force is not None:
        if ties is not None:
            ties.extend(np.abs(Z1[R1 != force[1]]).tolist())
        R1 = force[1]
    M1 = R1.astype(X.dtype)
    if m1 is not None:
        M1 = M1 * m1
    H1 = Z1 * M1
    L = H1 @ Wh.T + bh
    return M0, H0, M1, H1, L


def episode(X, y, Xq, yq, im, hp0, alpha, steps, masks=None, want_grad=False, first_order=False,
            loss_scale=1.0, relu_gates=None):
    """One task.  im = (W0,b0,W1,b1) meta-parameters; hp0 [N,65] head init (hypernet output for
    FuMI, lin_final [W|b] for MAML).  masks: None or dict(m0=[S+1,n?,H0], m1=[S+1,.,H1]) given as
    lists: masks['sup'][s] = (m0 [NK,H0], m1 [NK,H1]) and masks['qry'] = (m0q, m1q).

    Returns dict(loss, acc, preds, logits, step_losses, adapted=(W0,b0,W1,b1,hp) and, if
    want_grad, grads=(dW0,db0,dW1,db1,dhp0) of loss*loss_scale)."""
    dt = X.dtype
    W0, b0, W1, b1 = [a.astype(dt) for a in im]
    hp = hp0.astype(dt)
    n, N = X.shape[0], hp.shape[0]
    Y = np.zeros((n, N), dt)
    Y[np.arange(n), y] = 1
    th = [W0, b0, W1, b1, hp[:, :-1].copy(), hp[:, -1].copy()]
    cache, step_losses, ties = [], [], []
    for s in range(steps):
        m0, m1 = (masks["sup"][s] if masks is not None else (None, None))
        M0, H0, M1, H1, L = _mlp(X, th, m0, m1, None if relu_gates is None else relu_gates[s], ties)
        step_losses.append(_ce_mean(L, y))
        P = _softmax(L)
        dL = (P - Y) / n
        dWh = dL.T @ H1
        dbh = dL.sum(0)
        dZ1 = (dL @ th[4]) * M1
        dW1 = dZ1.T @ H0
        db1 = dZ1.sum(0)
        dZ0 = (dZ1 @ th[2]) * M0
        dW0 = dZ0.T @ X
        db0 = dZ0.sum(0)
        cache.append((th, M0, H0, M1, H1, P, dL, dZ1, dZ0))
        g = [dW0, db0, dW1, db1, dWh, dbh]
        th = [t - dt.type(alpha) * gi for t, gi in zip(th, g)]
    m0q, m1q = (masks["qry"] if masks is not None else (None, None))
    M0q, H0q, M1q, H1q, Lq = _mlp(Xq, th, m0q, m1q)
    loss = _ce_mean(Lq, yq)
    preds = Lq.argmax(axis=1)                     # first max on ties, as torch.max (fumi.py:180)
    acc = float(np.mean(preds == yq))
    out = dict(loss=loss, acc=acc, preds=preds.astype(np.int64), logits=Lq, step_losses=step_losses, relu_ties=ties,
               adapted=(th[0], th[1], th[2], th[3], np.concatenate([th[4], th[5][:, None]], 1)))
    if not want_grad:
        return out
    # ---- reverse sweep -------------------------------------------------------------------
    mq = Xq.shape[0]
    Yq = np.zeros((mq, N), dt)
    Yq[np.arange(mq), yq] = 1
    dLq = (_softmax(Lq) - Yq) * dt.type(loss_scale / mq)
    a_Wh = dLq.T @ H1q
    a_bh = dLq.sum(0)
    dZ1q = (dLq @ th[4]) * M1q
    a_W1 = dZ1q.T @ H0q
    a_b1 = dZ1q.sum(0)
    dZ0q = (dZ1q @ th[2]) * M0q
    a_W0 = dZ0q.T @ Xq
    a_b0 = dZ0q.sum(0)
    a = [a_W0, a_b0, a_W1, a_b1, a_Wh, a_bh]
    if not first_order:
        al = dt.type(alpha)
        for s in reversed(range(steps)):
            (W0s, b0s, W1s, b1s, Whs, bhs), M0, H0, M1, H1, P, dL, dZ1, dZ0 = cache[s]
            gW0, gb0, gW1, gb1, gWh, gbh = [-al * ai for ai in a]      # adjoints of the step's grads
            # reverse of: dZ0 = dH0*M0 ; dW0 = dZ0^T X ; db0 = sum dZ0
            r_dH0 = (X @ gW0.T + gb0) * M0
            # reverse of: dH0 = dZ1 @ W1
            r_dZ1 = r_dH0 @ W1s.T
            r_W1 = dZ1.T @ r_dH0
            # reverse of: dW1 = dZ1^T H0 ; db1 = sum dZ1
            r_dZ1 = r_dZ1 + H0 @ gW1.T + gb1
            r_H0 = dZ1 @ gW1
            # reverse of: dZ1 = dH1*M1 ; dH1 = dL @ Wh
            r_dH1 = r_dZ1 * M1
            r_dL = r_dH1 @ Whs.T
            r_Wh = dL.T @ r_dH1
            # reverse of: dWh = dL^T H1 ; dbh = sum dL
            r_dL = r_dL + H1 @ gWh.T + gbh
            r_H1 = dL @ gWh
            # reverse of: dL = (softmax(L) - Y)/n
            r_L = P * (r_dL - (P * r_dL).sum(axis=1, keepdims=True)) / dt.type(n)
            # reverse of the forward MLP
            r_H1 = r_H1 + r_L @ Whs
            r_Wh = r_Wh + r_L.T @ H1
            r_bh = r_L.sum(0)
            r_Z1 = r_H1 * M1
            r_H0 = r_H0 + r_Z1 @ W1s
            r_W1 = r_W1 + r_Z1.T @ H0
            r_b1 = r_Z1.sum(0)
            r_Z0 = r_H0 * M0
            r_W0 = r_Z0.T @ X
            r_b0 = r_Z0.sum(0)
            a = [a[0] + r_W0, a[1] + r_b0, a[2] + r_W1, a[3] + r_b1, a[4] + r_Wh, a[5] + r_bh]
    out["grads"] = (a[0], a[1], a[2], a[3], np.concatenate([a[4], a[5][:, None]], 1))
    return out


def fumi_batch(params, batch, alpha, steps, tanh=False, masks=None, want_grad=False, dtype=np.float32,
               relu_gates=None):
    """FuMI meta-batch (fumi.py:115-196).  params: dict keyed by the reference state_dict names.
    batch: dict(sup_x [B,NK,D], sup_y, qry_x [B,NQ,D], qry_y, class_text [B,N,T] = description
    embedding of the class carrying label i, picked as fumi.py:207-210)."""
    g = lambda k: np.asarray(params[k], dtype)
    hyper = (g("hyper_net.0.weight"), g("hyper_net.0.bias"), g("hyper_net.2.weight"), g("hyper_net.2.bias"))
    im = (g("im_net.linear0.weight"), g("im_net.linear0.bias"), g("im_net.linear1.weight"), g("im_net.linear1.bias"))
    B = batch["sup_x"].shape[0]
    res, grads = [], None
    for b in range(B):
        text = np.asarray(batch["class_text"][b], dtype)
        u, hp0 = hypernet_forward(hyper, text, tanh)
        r = episode(np.asarray(batch["sup_x"][b], dtype), batch["sup_y"][b], np.asarray(batch["qry_x"][b], dtype),
                    batch["qry_y"][b], im, hp0, alpha, steps, masks=None if masks is None else masks[b],
                    want_grad=want_grad, loss_scale=1.0 / B,
                    relu_gates=None if relu_gates is None else relu_gates[b])
        r["hp0"] = hp0
        if want_grad:
            dW0, db0, dW1, db1, dhp0 = r["grads"]
            dh = hypernet_backward(hyper, text, u, hp0, dhp0, tanh)
            gb = [dh[0], dh[1], dh[2], dh[3], dW0, db0, dW1, db1]
            grads = gb if grads is None else [x + y for x, y in zip(grads, gb)]
        res.append(r)
    out = dict(loss=dtype(np.sum([r["loss"] for r in res], dtype=dtype) / B),
               acc=dtype(np.sum([r["acc"] for r in res], dtype=dtype) / B),
               preds=np.stack([r["preds"] for r in res]), logits=np.stack([r["logits"] for r in res]),
               tasks=res)
    if want_grad:
        names = ["hyper_net.0.weight", "hyper_net.0.bias", "hyper_net.2.weight", "hyper_net.2.bias",
                 "im_net.linear0.weight", "im_net.linear0.bias", "im_net.linear1.weight", "im_net.linear1.bias"]
        out["grads"] = dict(zip(names, grads))
    return out


def maml_batch(params, batch, alpha, steps, first_order=False, want_grad=False, dtype=np.float32):
    """MAML meta-batch (maml.py:134-193): shared lin_final head, no dropout."""
    g = lambda k: np.asarray(params[k], dtype)
    im = (g("net.lin_0.weight"), g("net.lin_0.bias"), g("net.lin_1.weight"), g("net.lin_1.bias"))
    hp0 = np.concatenate([g("net.lin_final.weight"), g("net.lin_final.bias")[:, None]], 1)
    B = batch["sup_x"].shape[0]
    res, grads = [], None
    for b in range(B):
        r = episode(np.asarray(batch["sup_x"][b], dtype), batch["sup_y"][b], np.asarray(batch["qry_x"][b], dtype),
                    batch["qry_y"][b], im, hp0, alpha, steps, want_grad=want_grad, first_order=first_order,
                    loss_scale=1.0 / B)
        if want_grad:
            dW0, db0, dW1, db1, dhp0 = r["grads"]
            gb = [dW0, db0, dW1, db1, dhp0[:, :-1], dhp0[:, -1]]
            grads = gb if grads is None else [x + y for x, y in zip(grads, gb)]
        res.append(r)
    out = dict(loss=dtype(np.sum([r["loss"] for r in res], dtype=dtype) / B),
               acc=dtype(np.sum([r["acc"] for r in res], dtype=dtype) / B),
               preds=np.stack([r["preds"] for r in res]), logits=np.stack([r["logits"] for r in res]),
               tasks=res)
    if want_grad:
        names = ["net.lin_0.weight", "net.lin_0.bias", "net.lin_1.weight", "net.lin_1.bias",
                 "net.lin_final.weight", "net.lin_final.bias"]
        out["grads"] = dict(zip(names, grads))
    return out


def adam_step(p, g, m, v, step, lr, wd=0.0, b1=0.9, b2=0.999, eps=1e-8):
    """torch.optim.Adam single-tensor update (utils.py:280-283): L2 weight decay added to the
    gradient; step is 1-based.  Returns (p, m, v) new arrays, computed in the dtype of p."""
    dt = p.dtype
    g = g + dt.type(wd) * p
    m = m + (g - m) * dt.type(1 - b1)               # torch: exp_avg.lerp_(grad, 1-beta1)
    v = v * dt.type(b2) + dt.type(1 - b2) * g * g
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    step_size = lr / bc1
    denom = np.sqrt(v) / dt.type(np.sqrt(bc2)) + dt.type(eps)
    p = p - dt.type(step_size) * (m / denom)
    return p, m, v


def am3_batch(params, batch, num_ways, lamda_fixed=None, dtype=np.float32):
    """AM3 meta-test scoring (am3.py:90-126,159-200; utils.py:302-402), eval mode (dropout off).
    batch: sup_x [B,NK,D], sup_y, sup_text [B,NK,T], qry_x [B,NQ,D], qry_y."""
    g = lambda k: np.asarray(params[k], dtype)
    Wim, bim = g("image_encoder.weight"), g("image_encoder.bias")
    Wg0, bg0, Wg3, bg3 = g("g.0.weight"), g("g.0.bias"), g("g.3.weight"), g("g.3.bias")
    Wh0, bh0, Wh3, bh3 = g("h.0.weight"), g("h.0.bias"), g("h.3.weight"), g("h.3.bias")
    sx, qx = np.asarray(batch["sup_x"], dtype), np.asarray(batch["qry_x"], dtype)
    st = np.asarray(batch["sup_text"], dtype)
    sy, qy = batch["sup_y"], batch["qry_y"]
    B, NK = sy.shape
    es = sx @ Wim.T + bim
    eq = qx @ Wim.T + bim
    t = np.maximum(st @ Wg0.T + bg0, 0) @ Wg3.T + bg3
    lam = 1.0 / (1.0 + np.exp(-(np.maximum(t @ Wh0.T + bh0, 0) @ Wh3.T + bh3)))
    if lamda_fixed == 0:
        lam = np.zeros_like(lam)
    elif lamda_fixed == 1:
        lam = np.ones_like(lam)
    P = es.shape[-1]
    protos = np.zeros((B, num_ways, P), dtype)
    for b in range(B):
        for c in range(num_ways):
            sel = sy[b] == c
            cnt = max(int(sel.sum()), 1)
            ebar = es[b][sel].sum(0) / dtype(cnt)
            tbar = t[b][sel].sum(0) / dtype(cnt)
            lbar = lam[b][sel].sum(0) / dtype(cnt)
            protos[b, c] = lbar * ebar + (1 - lbar) * tbar
    d = ((protos[:, None, :, :] - eq[:, :, None, :]) ** 2).sum(-1)          # [B,NQ,N]
    preds = d.argmin(-1)                                                     # first min (utils.py:317)
    L = -d.reshape(-1, num_ways)
    loss = _ce_mean(L, qy.reshape(-1))
    return dict(loss=dtype(loss), preds=preds.astype(np.int64), dist=d, protos=protos, lamda=lam[..., 0],
                acc=float(np.mean(preds == qy)), avg_lamda=dtype(lam.mean()))
