"""Host orchestration of the episodic hot path over the C ABI (include/fumi_b200.h).

One meta-batch of B tasks = a fixed sequence of launches on the current CUDA stream:

  FuMI (fumi/models/fumi.py:115-196)                      MAML (fumi/models/maml.py:134-193)
  1 hypernetwork over the class text rows                 (head table = lin_final [W|b])
  2 first-layer projection of the feature rows   proj = X W0^T        (fumi_linear_fwd)
  3 Gram blocks of every task                    G = X X^T            (fumi_gram)
  4 fused inner loop + query scoring                                  (fumi_episode_fwd)
  train only:
  5 fused second-order backward                                       (fumi_episode_bwd)
  6 dW0 = d_proj^T X, reductions of the per-CTA partials, hypernetwork backward
  7 (multi-GPU) one all-reduce of the flat meta-gradient; the optimizer step follows in evaluate().

PyTorch supplies device memory, streams and torch.distributed only; every arithmetic step is a
kernel of libfumi_b200.so.  There is no CPU path: a non-CUDA device raises.
"""
import ctypes as C
import os

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from .data.bank import EpisodeBatch

H0, H1 = 256, 64
HD = H1 + 1
# dense layers: 0 fp32 FMA; 1 tcgen05 3xTF32; 2 = 1 + the two bank-sized contractions and the Gram blocks on fp16 planes
DEFAULT_PRECISION = 2


def first_row_of_each_label(targets, num_ways):
    """[B,N] index of the first row with label i (fumi.py:208-210: (targets==i).nonzero()[0][0])."""
    B, n = targets.shape
    eq = targets.unsqueeze(-1) == torch.arange(num_ways, device=targets.device).view(1, 1, -1)   # [B,n,N]
    pos = torch.arange(n, device=targets.device).view(1, n, 1).expand(B, n, num_ways)
    first = torch.where(eq, pos, torch.full_like(pos, n)).min(dim=1).values
    if bool((first >= n).any()):
        raise IndexError("index 0 is out of bounds for dimension 0 with size 0")   # label without support row
    return first


class EpisodeEngine:
    def __init__(self, device, precision=None):
        self.device = torch.device(device)
        _lib.require_cuda(self.device, "fumi_b200.EpisodeEngine")
        self.L = _lib.lib()
        self.precision = DEFAULT_PRECISION if precision is None else precision
        self.launches = 0                # kernels launched through this engine (bench: gpu_launches)
        self.profile = None              # dict name -> [cuda event pairs] while bench.py profiles kernels

    def _call(self, name, fn, *args):
        """Invoke one C-ABI entry point; under bench.py's per-kernel pass bracket it with CUDA events
        on the launching stream."""
        if self.profile is None or self.device.type != "cuda":
            return _lib.check(fn(*args), name)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = _lib.check(fn(*args), name)
        e1.record()
        self.profile.setdefault(name, []).append((e0, e1))
        return rc

    # ------------------------------------------------------------------ thin wrappers over the C ABI
    def _stream(self):
        return _lib.stream_ptr(self.device) if self.device.type == "cuda" else None

    def _new(self, *shape, dtype=torch.float32):
        return torch.empty(*shape, dtype=dtype, device=self.device)

    def linear_fwd(self, x, w, bias=None, act=0, precision=None):
        M, K = x.shape
        N = w.shape[0]
        y = self._new(M, N)
        self._call("fumi_linear_fwd", self.L.fumi_linear_fwd, _lib.ptr(x), _lib.ptr(w), _lib.ptr(bias), _lib.ptr(y), M, N, K, act,
                                          self.precision if precision is None else precision, self._stream())
        self.launches += 1
        return y

    def linear_wgrad(self, dy, x, dw, db=None, accumulate=False, precision=None):
        M, N = dy.shape
        K = x.shape[1]
        self._call("fumi_linear_wgrad", self.L.fumi_linear_wgrad, _lib.ptr(dy), _lib.ptr(x), _lib.ptr(dw), _lib.ptr(db), M, N, K,
                                            int(accumulate), self.precision if precision is None else precision,
                                            self._stream())
        self.launches += 2 + (db is not None)

    def linear_dgrad(self, dy, w, gate=None):
        M, N = dy.shape
        K = w.shape[1]
        dx = self._new(M, K)
        self._call("fumi_linear_dgrad", self.L.fumi_linear_dgrad, _lib.ptr(dy), _lib.ptr(w), _lib.ptr(gate), _lib.ptr(dx), M, N, K,
                                            self._stream())
        self.launches += 1
        return dx

    # ---- tcgen05 3xTF32 path (precision 1): operands are (hi, lo) plane pairs ---------------------
    def split_tf32(self, x):
        x = x.contiguous()
        hi, lo = torch.empty_like(x), torch.empty_like(x)
        self._call("fumi_split_tf32", self.L.fumi_split_tf32, _lib.ptr(x), _lib.ptr(hi), _lib.ptr(lo), x.numel(),
                   self._stream())
        self.launches += 1
        return hi, lo

    def transpose_split_tf32(self, x):
        """x [R,C] -> (hiT, loT) [C, ldt] with ldt = R rounded up to 32 (K-major planes over the rows)."""
        R, Cc = x.shape
        ldt = (R + 31) // 32 * 32
        hiT, loT = self._new(Cc, ldt), self._new(Cc, ldt)
        self._call("fumi_transpose_split_tf32", self.L.fumi_transpose_split_tf32, _lib.ptr(x), _lib.ptr(hiT),
                   _lib.ptr(loT), R, Cc, ldt, self._stream())
        self.launches += 1
        return hiT, loT

    def gemm_tc(self, a, b, bias=None, act=0, out=None, accumulate=False, split_k=0, K=None):
        """out[M,N] (=|+=) act(A . B^T + bias) on tcgen05; a, b = (hi, lo) planes [rows, ld] with K <= ld."""
        M, lda = a[0].shape
        N, ldb = b[0].shape
        K = min(lda, ldb) if K is None else K
        if out is None:
            out = self._new(M, N)
        self._call("fumi_gemm_tf32x3", self.L.fumi_gemm_tf32x3, _lib.ptr(a[0]), _lib.ptr(a[1]), _lib.ptr(b[0]),
                   _lib.ptr(b[1]), _lib.ptr(bias), _lib.ptr(out), M, N, K, lda, ldb, out.stride(0), act,
                   int(accumulate), split_k, self._stream())
        self.launches += 1
        return out

    # ---- tcgen05 3 x fp16 path (precision 2): scaled fp16 (hi, lo) planes + the device scalar max|x| --------
    def absmax(self, x):
        out = torch.empty(1, dtype=torch.float32, device=x.device)
        self._call("fumi_absmax", self.L.fumi_absmax, _lib.ptr(x), x.numel(), _lib.ptr(out), self._stream())
        self.launches += 1
        return out

    def split_f16(self, x):
        x = x.contiguous()
        am = self.absmax(x)
        hi = torch.empty(x.shape, dtype=torch.float16, device=x.device)
        lo = torch.empty_like(hi)
        self._call("fumi_split_f16", self.L.fumi_split_f16, _lib.ptr(x), _lib.ptr(am), _lib.ptr(hi), _lib.ptr(lo),
                   x.numel(), self._stream())
        self.launches += 1
        return hi, lo, am

    def transpose_split_f16(self, x):
        """x [R,C] -> (hiT, loT, absmax), planes [C, ldt] fp16 with ldt = R rounded up to 64."""
        x = x.contiguous()
        R, Cc = x.shape
        ldt = (R + 63) // 64 * 64
        am = self.absmax(x)
        hiT = torch.empty((Cc, ldt), dtype=torch.float16, device=x.device)
        loT = torch.empty_like(hiT)
        self._call("fumi_transpose_split_f16", self.L.fumi_transpose_split_f16, _lib.ptr(x), _lib.ptr(am), _lib.ptr(hiT),
                   _lib.ptr(loT), R, Cc, ldt, self._stream())
        self.launches += 1
        return hiT, loT, am

    def gemm_f16(self, a, b, bias=None, act=0, out=None, accumulate=False, split_k=0, K=None):
        """out[M,N] (=|+=) act(A . B^T + bias); a, b = (hi, lo, absmax) from split_f16 / transpose_split_f16."""
        M, lda = a[0].shape
        N, ldb = b[0].shape
        K = min(lda, ldb) if K is None else K
        if out is None:
            out = self._new(M, N)
        self._call("fumi_gemm_f16x3", self.L.fumi_gemm_f16x3, _lib.ptr(a[0]), _lib.ptr(a[1]), _lib.ptr(b[0]),
                   _lib.ptr(b[1]), _lib.ptr(a[2]), _lib.ptr(b[2]), _lib.ptr(bias), _lib.ptr(out), M, N, K, lda, ldb,
                   out.stride(0), act, int(accumulate), split_k, self._stream())
        self.launches += 1
        return out

    def _feat_planes16(self, feats, bank, transposed):
        key = "_f16T" if transposed else "_f16"
        fn = self.transpose_split_f16 if transposed else self.split_f16
        if bank is not None:
            if getattr(bank, key, None) is None:
                setattr(bank, key, fn(feats))
            return getattr(bank, key)
        return fn(feats)

    def _feat_planes(self, feats, bank):
        """(hi, lo) of the feature matrix; cached on the FeatureBank (static data, split once)."""
        if bank is not None:
            if getattr(bank, "_tf32", None) is None:
                bank._tf32 = self.split_tf32(feats)
            return bank._tf32
        return self.split_tf32(feats)

    def _feat_planes_T(self, feats, bank):
        if bank is not None:
            if getattr(bank, "_tf32T", None) is None:
                bank._tf32T = self.transpose_split_tf32(feats)
            return bank._tf32T
        return self.transpose_split_tf32(feats)

    def project_rows(self, feats, bank, w0):
        """proj = X W0^T (first image layer over every feature row, no bias)."""
        if self.precision == 2 and feats.shape[1] % 8 == 0:
            return self.gemm_f16(self._feat_planes16(feats, bank, False), self.split_f16(w0))
        if self.precision >= 1:
            return self.gemm_tc(self._feat_planes(feats, bank), self.split_tf32(w0))
        return self.linear_fwd(feats, w0, None, act=0, precision=0)

    def wgrad_rows(self, d_proj, feats, bank, dw0):
        """dW0 = d_proj^T X."""
        if self.precision == 2:
            R = d_proj.shape[0]
            self.gemm_f16(self.transpose_split_f16(d_proj), self._feat_planes16(feats, bank, True), out=dw0, K=R)
        elif self.precision >= 1:
            R = d_proj.shape[0]
            self.gemm_tc(self.transpose_split_tf32(d_proj), self._feat_planes_T(feats, bank), out=dw0, K=R)
        else:
            self.linear_wgrad(d_proj, feats, dw0, precision=0)

    def gram(self, feats, sup_rows, qry_rows, bank=None):
        B, NK = sup_rows.shape
        NQ = qry_rows.shape[1]
        g = self._new(B, NK + NQ, NK)
        D = feats.shape[1]
        if self.precision == 2 and NK <= 32 and NK + NQ <= 192 and D % 64 == 0:      # fp16 bank planes (split once)
            hi, lo, am = self._feat_planes16(feats, bank, False)
            self._call("fumi_gram", self.L.fumi_gram_f16, _lib.ptr(hi), _lib.ptr(lo), _lib.ptr(am), feats.shape[0], D,
                       _lib.ptr(sup_rows), _lib.ptr(qry_rows), B, NK, NQ, _lib.ptr(g), self._stream())
        else:
            self._call("fumi_gram", self.L.fumi_gram, _lib.ptr(feats), feats.shape[0], D, _lib.ptr(sup_rows),
                       _lib.ptr(qry_rows), B, NK, NQ, _lib.ptr(g), self._stream())
        self.launches += 1
        return g

    def make_cfg(self, N, NK, NQ, steps, step_size, dropout_p=0.0, dropout_seed=0, task_offset=0, first_order=False,
                 save=False):
        return _lib.EpisodeCfg(num_ways=N, num_support=NK, num_query=NQ, hid0=H0, hid1=H1, steps=steps,
                               step_size=step_size, dropout_p=dropout_p, dropout_seed=dropout_seed,
                               task_offset=task_offset, first_order=int(first_order), reserved=int(save))

    def stash_layout(self, cfg):
        lay = _lib.StashLayout()
        _lib.check(self.L.fumi_stash_layout(C.byref(cfg), C.byref(lay)), "fumi_stash_layout")
        return lay

    def episode_fwd(self, cfg, proj, eb, gram, b0, w1, b1, head_table, head_rows):
        B, NQ, N = eb.sup_rows.shape[0], cfg.num_query, cfg.num_ways
        per = _lib.check(self.L.fumi_episode_stash_floats(C.byref(cfg)), "fumi_episode_stash_floats")
        slots = B if cfg.reserved else min(B, max(1, _lib.check(self.L.fumi_device_sm_count(), "sm_count")))
        out = dict(logits=self._new(B, NQ, N), preds=self._new(B, NQ, dtype=torch.int64), task_loss=self._new(B),
                   task_acc=self._new(B), stash=self._new(slots * per), stash_per_task=per)
        self._call("fumi_episode_fwd", self.L.fumi_episode_fwd, 
            C.byref(cfg), B, _lib.ptr(proj), _lib.ptr(eb.sup_rows), _lib.ptr(eb.qry_rows), _lib.ptr(eb.sup_y),
            _lib.ptr(eb.qry_y), _lib.ptr(gram), _lib.ptr(b0), _lib.ptr(w1), _lib.ptr(b1), _lib.ptr(head_table),
            _lib.ptr(head_rows), _lib.ptr(out["logits"]), _lib.ptr(out["preds"]), _lib.ptr(out["task_loss"]),
            _lib.ptr(out["task_acc"]), _lib.ptr(out["stash"]), self._stream())
        self.launches += 1
        return out

    def episode_bwd(self, cfg, proj, eb, gram, stash, loss_scale, d_proj):
        B, N = eb.sup_rows.shape[0], cfg.num_ways
        P = _lib.check(self.L.fumi_episode_bwd_parts(), "fumi_episode_bwd_parts")
        parts = torch.zeros(P * (H0 + H0 * H1 + H1), dtype=torch.float32, device=self.device)
        pb0, pw1, pb1 = parts[:P * H0], parts[P * H0:P * (H0 + H0 * H1)], parts[P * (H0 + H0 * H1):]
        d_head = self._new(B, N, HD)
        self._call("fumi_episode_bwd", self.L.fumi_episode_bwd, 
            C.byref(cfg), B, _lib.ptr(proj), _lib.ptr(eb.sup_rows), _lib.ptr(eb.qry_rows), _lib.ptr(eb.sup_y),
            _lib.ptr(eb.qry_y), _lib.ptr(gram), _lib.ptr(stash), float(loss_scale), _lib.ptr(d_proj), _lib.ptr(d_head),
            _lib.ptr(pb0), _lib.ptr(pw1), _lib.ptr(pb1), self._stream())
        self.launches += 2
        return d_head, (pb0, pw1, pb1, P)

    def reduce_parts(self, parts, P, out):
        n = out.numel()
        self._call("fumi_reduce_parts", self.L.fumi_reduce_parts, _lib.ptr(parts), P, n, _lib.ptr(out), 0, self._stream())
        self.launches += 1

    def loss_acc(self, task_loss, task_acc):
        out = self._new(2)
        self._call("fumi_reduce_loss_acc", self.L.fumi_reduce_loss_acc, _lib.ptr(task_loss), _lib.ptr(task_acc), task_loss.numel(),
                                               _lib.ptr(out), self._stream())
        self.launches += 1
        return out

    def loss_acc_early(self, la, reduce=True):
        """Start the device -> host copy of [loss, acc] as soon as the forward has produced them (they do not depend on
        the backward or the optimizer step), into a pinned buffer with an event behind it.  `read_loss_acc` then waits for
        that event only, so the caller gets the step's loss (fumi.py:195-196) while the device is still in the backward,
        and enqueues the next step without ever letting the stream run dry.  With more than one rank and `reduce` (a
        training step: every rank is in the gradient all-reduce anyway) the two scalars are averaged over ranks first,
        by a 2-float all-reduce on a side stream (the compute stream never waits for it); evaluation passes stay local,
        as before -- no collective is added where only some ranks may be evaluating."""
        if self.device.type != "cuda" or os.environ.get("FUMI_EARLY_LOSS", "1") == "0":
            return None
        world = self._world() if reduce else 1
        main = torch.cuda.current_stream(self.device)
        if getattr(self, "_la_host", None) is None:
            self._la_host = torch.empty(2, dtype=torch.float32, pin_memory=True)
            self._la_event = torch.cuda.Event()
            self._la_tmp = torch.empty(2, dtype=torch.float32, device=self.device)
            self._la_stream = torch.cuda.Stream(self.device)
        if world == 1:
            self._la_host.copy_(la, non_blocking=True)
            self._la_event.record(main)
            return self._la_event
        side = self._la_stream
        side.wait_stream(main)
        la.record_stream(side)
        with torch.cuda.stream(side):
            self._la_tmp.copy_(la)
            dist.all_reduce(self._la_tmp)
            self._la_tmp /= world
            self._la_host.copy_(self._la_tmp, non_blocking=True)
            self._la_event.record(side)
        return self._la_event

    def read_loss_acc(self, res):
        """[loss, acc] of a batch as a host array (the one device sync per batch of the reference API)."""
        ev = res.get("loss_acc_event")
        if ev is None:
            return res["loss_acc"].cpu().numpy()
        ev.synchronize()
        return self._la_host.numpy().copy()

    # ------------------------------------------------------------------ batch plumbing
    def _to_dev(self, t, dtype=None):
        if isinstance(t, np.ndarray):
            t = torch.from_numpy(np.ascontiguousarray(t))
        t = t.to(self.device, non_blocking=True)
        if dtype is not None and t.dtype != dtype:
            t = t.to(dtype)
        return t.contiguous()

    def unpack(self, batch, num_ways, want_text):
        """-> (EpisodeBatch on device, feats [R,D], text rows [Rt,T] or None, head_rows [B,N] or None)"""
        if isinstance(batch, EpisodeBatch):
            eb = batch.to(self.device)
            return eb, eb.bank.feats, (eb.bank.text if want_text else None), (eb.head_class if want_text else None)
        # the reference's torchmeta batch dict (fumi.py:129-144): [[ids, text, im], targets]
        (tr_ids, tr_text, tr_im), tr_y = batch["train"]
        (te_ids, te_text, te_im), te_y = batch["test"]
        B, NK, D = tr_im.shape
        NQ = te_im.shape[1]
        tr_y, te_y = self._to_dev(tr_y, torch.int64), self._to_dev(te_y, torch.int64)
        feats = torch.cat([self._to_dev(tr_im, torch.float32).reshape(B * NK, D),
                           self._to_dev(te_im, torch.float32).reshape(B * NQ, D)], 0)
        ar = torch.arange(B * (NK + NQ), device=self.device, dtype=torch.int64)
        eb = EpisodeBatch(bank=None, sup_rows=ar[:B * NK].view(B, NK).contiguous(),
                          qry_rows=ar[B * NK:].view(B, NQ).contiguous(), sup_y=tr_y, qry_y=te_y,
                          sup_ids=tr_ids, qry_ids=te_ids, head_class=None)
        text_rows = head_rows = None
        if want_text:
            first = first_row_of_each_label(tr_y, num_ways)                         # [B,N]
            tr_text = self._to_dev(tr_text, torch.float32)
            text_rows = torch.gather(tr_text, 1, first.unsqueeze(-1).expand(B, num_ways, tr_text.shape[-1]))
            text_rows = text_rows.reshape(B * num_ways, -1).contiguous()
            head_rows = torch.arange(B * num_ways, device=self.device, dtype=torch.int64).view(B, num_ways)
        return eb, feats, text_rows, head_rows

    @staticmethod
    def _grad(p):
        if p.grad is None:
            p.grad = torch.zeros_like(p)
        return p.grad

    def _world(self):
        if dist.is_available() and dist.is_initialized():
            return dist.get_world_size()
        return 1

    def _flat_grads(self, params):
        """FusedAdam's flat gradient buffer [n gradients | loss, acc] when every p.grad still lives in it."""
        opt = getattr(params[0], "_fumi_flat_opt", None) if params else None
        return opt.flat_grad(params) if opt is not None else None

    def _allreduce_begin(self, params, head):
        """Multi-GPU outer step, part 1: the gradient bucket that is complete first -- the first
        `head` floats of the flat buffer = dW0, 2 MB of the 3 MB -- is summed over ranks asynchronously while the
        hypernetwork backward and the small reductions still run on the compute stream.  Returns the handle(s)."""
        if self._world() == 1:
            return None
        flat = self._flat_grads(params)
        if flat is None or head <= 0:
            return None
        return (flat, head, dist.all_reduce(flat[:head], async_op=True))

    def _allreduce_finish(self, params, loss_acc, pending):
        """Part 2: [remaining gradients | loss, acc] in ONE all-reduce (the two scalars ride in the flat buffer's
        tail); the per-tensor path is kept for optimizers without a flat buffer.  loss_acc becomes the mean."""
        world = self._world()
        if world == 1:
            return
        flat = self._flat_grads(params)
        if flat is not None:
            head = pending[1] if pending is not None else 0
            flat[-2:].copy_(loss_acc)
            w = dist.all_reduce(flat[head:], async_op=True)
            if pending is not None:
                pending[2].wait()
            w.wait()
            loss_acc.copy_(flat[-2:])
        else:
            if pending is not None:
                pending[2].wait()
            for p in params:
                dist.all_reduce(p.grad)
            dist.all_reduce(loss_acc)
        loss_acc /= world

    # ------------------------------------------------------------------ FuMI
    def hypernet(self, model, text_rows, keep=False, bank=None):
        """hyper_net(text): Linear-ReLU-Linear(-Tanh)  (fumi.py:70-107,109-113).  `bank`: the FeatureBank whose static
        description rows `text_rows` are (their operand planes are split once and cached on it)."""
        l0, l2 = model.hyper_net[0], model.hyper_net[2]
        if self.precision >= 1 and text_rows.shape[1] % 4 == 0:
            if bank is not None and text_rows is bank.text:
                if getattr(bank, "_text_tf32", None) is None:
                    bank._text_tf32 = self.split_tf32(text_rows)
                tplanes = bank._text_tf32
            else:
                tplanes = self.split_tf32(text_rows)
            u = self.gemm_tc(tplanes, self.split_tf32(l0.weight), bias=l0.bias, act=1)
        else:
            u = self.linear_fwd(text_rows, l0.weight, l0.bias, act=1, precision=0)
        hp = self.linear_fwd(u, l2.weight, l2.bias, act=2 if model.norm_hypernet else 0, precision=0)
        return (hp, u) if keep else hp

    def _check_im_net(self, hidden, emb_dim):
        if list(hidden) != [H0, H1]:
            raise NotImplementedError(f"this build adapts an image MLP with --im_hid_dim {H0} {H1} (the reference "
                                      f"default); got {list(hidden)}")
        if emb_dim % 4:
            raise NotImplementedError("--im_emb_dim must be a multiple of 4")

    def fumi_batch(self, model, batch, steps, step_size, train, return_state=False):
        self._check_im_net(model.im_hid_dim, model.im_emb_dim)
        N = model.n_way
        eb, feats, text_rows, head_rows = self.unpack(batch, N, want_text=True)
        if model.text_encoder_type == "rand":
            text_rows = 2 * torch.rand(text_rows.shape[0], model.text_emb_dim, device=self.device) - 1   # fumi.py:200-202
        B, NK = eb.sup_rows.shape
        NQ = eb.qry_rows.shape[1]
        lin0, lin1 = model.im_net.linear0, model.im_net.linear1
        p_drop = float(model.dropout_rate) if (train and model.dropout_rate > 0) else 0.0
        if p_drop > 0:
            model.dropout_seed += 1
        world = self._world()
        rank = dist.get_rank() if world > 1 else 0
        cfg = self.make_cfg(N, NK, NQ, steps, step_size, dropout_p=p_drop,
                            dropout_seed=(int(getattr(model, "dropout_base_seed", 0)) << 20) + model.dropout_seed,
                            task_offset=rank * B, save=train or return_state)
        # the bank-sized projection goes first: after the host sync that ends every step (loss / accuracy come back,
        # fumi.py:195-196) it gives the device 0.4 ms of work while the host enqueues the small hypernetwork launches
        proj = self.project_rows(feats, eb.bank, lin0.weight)
        hp_table, u = self.hypernet(model, text_rows, keep=True, bank=eb.bank)
        gram = self.gram(feats, eb.sup_rows, eb.qry_rows, bank=eb.bank)
        out = self.episode_fwd(cfg, proj, eb, gram, lin0.bias, lin1.weight, lin1.bias, hp_table, head_rows)
        la = self.loss_acc(out["task_loss"], out["task_acc"])
        res = dict(loss_acc=la, loss_acc_event=self.loss_acc_early(la, reduce=train), preds=out["preds"], logits=out["logits"], qry_y=eb.qry_y, task_loss=out["task_loss"],
                   task_acc=out["task_acc"], cfg=cfg, stash=out["stash"] if (train or return_state) else None,
                   hp_table=hp_table, head_rows=head_rows, batch=eb)
        if not train:
            return res
        # ---- outer backward (fumi.py:190-192): gradients of sum_b task_loss / B_global
        d_proj = torch.zeros_like(proj)
        d_head, (pb0, pw1, pb1, P) = self.episode_bwd(cfg, proj, eb, gram, out["stash"], 1.0 / (B * world), d_proj)
        self.reduce_parts(pb0, P, self._grad(lin0.bias))
        self.reduce_parts(pw1, P, self._grad(lin1.weight))
        self.reduce_parts(pb1, P, self._grad(lin1.bias))
        self.wgrad_rows(d_proj, feats, eb.bank, self._grad(lin0.weight))         # dW0 = d_proj^T X
        params = [p for p in model.parameters() if p.requires_grad]
        pending = self._allreduce_begin(params, lin0.weight.numel() if params[0] is lin0.weight else 0)
        d_hp = torch.zeros_like(hp_table)
        self._call("fumi_scatter_add_rows", self.L.fumi_scatter_add_rows, _lib.ptr(d_head), _lib.ptr(head_rows.reshape(-1)), B * N, HD,
                                                _lib.ptr(d_hp), self._stream())
        self.launches += 1
        if model.norm_hypernet:
            self._call("fumi_tanh_bwd", self.L.fumi_tanh_bwd, _lib.ptr(hp_table), _lib.ptr(d_hp), d_hp.numel(), self._stream())
            self.launches += 1
        l0, l2 = model.hyper_net[0], model.hyper_net[2]
        self.linear_wgrad(d_hp, u, self._grad(l2.weight), self._grad(l2.bias), precision=0)
        d_u = self.linear_dgrad(d_hp, l2.weight, gate=u)
        self.linear_wgrad(d_u, text_rows, self._grad(l0.weight), self._grad(l0.bias), precision=0)
        self._allreduce_finish(params, la, pending)
        return res

    # ------------------------------------------------------------------ MAML
    def maml_batch(self, model, batch, steps, step_size, train, first_order=False, return_state=False):
        self._check_im_net(model.hidden_dims, model.im_embed_dim)
        N = model.n_way
        eb, feats, _, _ = self.unpack(batch, N, want_text=False)
        B, NK = eb.sup_rows.shape
        NQ = eb.qry_rows.shape[1]
        lin0, lin1, fin = model.net.lin_0, model.net.lin_1, model.net.lin_final
        world = self._world()
        cfg = self.make_cfg(N, NK, NQ, steps, step_size, first_order=first_order, save=train or return_state)
        head_table = torch.cat([fin.weight, fin.bias.unsqueeze(1)], 1).contiguous()      # [N, 65]
        proj = self.project_rows(feats, eb.bank, lin0.weight)
        gram = self.gram(feats, eb.sup_rows, eb.qry_rows, bank=eb.bank)
        out = self.episode_fwd(cfg, proj, eb, gram, lin0.bias, lin1.weight, lin1.bias, head_table, None)
        la = self.loss_acc(out["task_loss"], out["task_acc"])
        res = dict(loss_acc=la, loss_acc_event=self.loss_acc_early(la, reduce=train), preds=out["preds"], logits=out["logits"], qry_y=eb.qry_y, cfg=cfg,
                   stash=out["stash"] if (train or return_state) else None, batch=eb)
        if not train:
            return res
        d_proj = torch.zeros_like(proj)
        d_head, (pb0, pw1, pb1, P) = self.episode_bwd(cfg, proj, eb, gram, out["stash"], 1.0 / (B * world), d_proj)
        self.reduce_parts(pb0, P, self._grad(lin0.bias))
        self.reduce_parts(pw1, P, self._grad(lin1.weight))
        self.reduce_parts(pb1, P, self._grad(lin1.bias))
        self.wgrad_rows(d_proj, feats, eb.bank, self._grad(lin0.weight))
        params = list(model.parameters())
        pending = self._allreduce_begin(params, lin0.weight.numel() if params[0] is lin0.weight else 0)
        d_fin = self._new(N * HD)
        self.reduce_parts(d_head.reshape(B, N * HD), B, d_fin)                   # shared head: sum over tasks
        d_fin = d_fin.view(N, HD)
        self._grad(fin.weight).copy_(d_fin[:, :H1])
        self._grad(fin.bias).copy_(d_fin[:, H1])
        self._allreduce_finish(params, la, pending)
        return res

    # ------------------------------------------------------------------ AM3
    def _dropout(self, x, seed, layer, p):
        if p > 0:
            self._call("fumi_dropout_apply", self.L.fumi_dropout_apply, _lib.ptr(x), x.shape[0], x.shape[1], C.c_uint64(seed),
                       C.c_uint32(layer), float(p), self._stream())
            self.launches += 1
        return x

    def am3_batch(self, model, batch, num_ways, train=False):
        """AM3.evaluate (am3.py:128-212): prototypes, distances, argmin, CE; with train=True also the gradients of
        the mean query CE wrt every parameter (loss.backward(), am3.py:188-190), left in p.grad.

        The text path is evaluated once per class description (one row of bank.text), as the support samples of a class
        share their text (am3.py:123 maps the same description K times).  In train mode the Dropout layers inside g / h
        (am3.py:66-88) use the counter-based masks of the episode kernels, one mask per class description and meta-batch
        (the reference draws one per support sample from the torch generator; parity is defined at --dropout 0)."""
        N = num_ways
        if isinstance(batch, EpisodeBatch):
            eb = batch.to(self.device)
            feats, text_rows, class_rows = eb.bank.feats, eb.bank.text, eb.head_class
        else:
            eb, feats, text_rows, class_rows = self.unpack(batch, N, want_text=True)
        P = model.prototype_dim
        p_drop = float(model.dropout) if train else 0.0
        if train:
            model.dropout_seed = getattr(model, "dropout_seed", 0) + 1
        seed = (int(getattr(model, "dropout_base_seed", 0)) << 20) + getattr(model, "dropout_seed", 0)
        if self.precision == 2 and feats.shape[1] % 8 == 0:      # the bank's fp16 planes (split once, shared with FuMI / MAML)
            emb = self.gemm_f16(self._feat_planes16(feats, eb.bank, False), self.split_f16(model.image_encoder.weight),
                                bias=model.image_encoder.bias)
        elif self.precision >= 1:
            emb = self.gemm_tc(self._feat_planes(feats, eb.bank), self.split_tf32(model.image_encoder.weight),
                               bias=model.image_encoder.bias)
        else:
            emb = self.linear_fwd(feats, model.image_encoder.weight, model.image_encoder.bias, precision=0)
        g0, g3, h0, h3 = model.g[0], model.g[3], model.h[0], model.h[3]
        u = self._dropout(self.linear_fwd(text_rows, g0.weight, g0.bias, act=1, precision=0), seed, 0, p_drop)
        t = self.linear_fwd(u, g3.weight, g3.bias, precision=0)
        z = self._dropout(self.linear_fwd(t, h0.weight, h0.bias, act=1, precision=0), seed, 1, p_drop)
        lam = self.linear_fwd(z, h3.weight, h3.bias, act=3, precision=0)
        B, NK = eb.sup_rows.shape
        NQ = eb.qry_rows.shape[1]
        protos, dist_, preds = self._new(B, N, P), self._new(B, NQ, N), self._new(B, NQ, dtype=torch.int64)
        task_loss = self._new(B)
        fixed = -1 if model.lamda_fixed is None else int(model.lamda_fixed)
        self._call("fumi_am3_score", self.L.fumi_am3_score,
            _lib.ptr(emb), _lib.ptr(t), _lib.ptr(lam.reshape(-1)), _lib.ptr(eb.sup_rows), _lib.ptr(eb.qry_rows),
            _lib.ptr(eb.sup_y), _lib.ptr(eb.qry_y), _lib.ptr(class_rows), B, N, NK, NQ, P, fixed, _lib.ptr(protos),
            _lib.ptr(dist_), _lib.ptr(preds), _lib.ptr(task_loss), self._stream())
        self.launches += 1
        # per-support-row lamda as the reference returns it (am3.py:208): lamda of the row's class
        label_lam = torch.gather(lam.reshape(-1)[class_rows], 1, eb.sup_y)
        if fixed == 0:
            label_lam = torch.zeros_like(label_lam)
        elif fixed == 1:
            label_lam = torch.ones_like(label_lam)
        counts = torch.zeros(N, N, dtype=torch.int64, device=self.device)      # confusion counts (utils.py:323-326)
        self._call("fumi_confusion_counts", self.L.fumi_confusion_counts, _lib.ptr(eb.qry_y), _lib.ptr(preds),
                   B * NQ, N, _lib.ptr(counts), self._stream())
        self.launches += 1
        res = dict(task_loss=task_loss, preds=preds, dist=dist_, protos=protos, sup_lamda=label_lam, batch=eb, confusion=counts)
        if not train:
            return res
        # ---- backward (am3.py:188-190): gradients of sum_b sum_q CE / (B_global * NQ)
        world = self._world()
        Cn = text_rows.shape[0]
        d_emb = torch.zeros_like(emb)
        d_tp, d_lam = self._new(B, N, P), self._new(B, N)
        self._call("fumi_am3_bwd", self.L.fumi_am3_bwd,
            _lib.ptr(emb), _lib.ptr(t), _lib.ptr(lam.reshape(-1)), _lib.ptr(eb.sup_rows), _lib.ptr(eb.qry_rows),
            _lib.ptr(eb.sup_y), _lib.ptr(eb.qry_y), _lib.ptr(class_rows), B, N, NK, NQ, P, fixed, _lib.ptr(protos),
            _lib.ptr(dist_), 1.0 / float(B * world * NQ), _lib.ptr(d_emb), _lib.ptr(d_tp), _lib.ptr(d_lam), self._stream())
        self.launches += 1
        rows_flat = class_rows.reshape(-1).contiguous()
        d_t2 = torch.zeros(2, Cn, P, dtype=torch.float32, device=self.device)          # [from prototypes | through h]
        d_lt = torch.zeros(Cn, 1, dtype=torch.float32, device=self.device)
        self._call("fumi_scatter_add_rows", self.L.fumi_scatter_add_rows, _lib.ptr(d_tp), _lib.ptr(rows_flat), B * N, P,
                   _lib.ptr(d_t2[0]), self._stream())
        self._call("fumi_scatter_add_rows", self.L.fumi_scatter_add_rows, _lib.ptr(d_lam), _lib.ptr(rows_flat), B * N, 1,
                   _lib.ptr(d_lt), self._stream())
        self.launches += 2
        # lamda = sigmoid(h(t)):  h = Linear(P, Th) - ReLU - Dropout - Linear(Th, 1)
        self._call("fumi_sigmoid_bwd", self.L.fumi_sigmoid_bwd, _lib.ptr(lam), _lib.ptr(d_lt), Cn, self._stream())
        self.launches += 1
        self.linear_wgrad(d_lt, z, self._grad(h3.weight), self._grad(h3.bias), precision=0)
        d_z = self._dropout(self.linear_dgrad(d_lt, h3.weight, gate=z), seed, 1, p_drop)   # z > 0 <=> ReLU gate and mask kept
        self.linear_wgrad(d_z, t, self._grad(h0.weight), self._grad(h0.bias), precision=0)
        self._call("fumi_linear_dgrad", self.L.fumi_linear_dgrad, _lib.ptr(d_z), _lib.ptr(h0.weight), None,
                   _lib.ptr(d_t2[1]), Cn, h0.weight.shape[0], P, self._stream())
        self.launches += 1
        d_t = self._new(Cn, P)
        self.reduce_parts(d_t2.reshape(2, Cn * P), 2, d_t.reshape(-1))
        # t = g(text):  g = Linear(T, Th) - ReLU - Dropout - Linear(Th, P)
        self.linear_wgrad(d_t, u, self._grad(g3.weight), self._grad(g3.bias), precision=0)
        d_u = self._dropout(self.linear_dgrad(d_t, g3.weight, gate=u), seed, 0, p_drop)
        self.linear_wgrad(d_u, text_rows, self._grad(g0.weight), self._grad(g0.bias), precision=0)
        # image encoder: dW = d_emb^T X over the bank, db = column sums
        self.wgrad_rows(d_emb, feats, eb.bank, self._grad(model.image_encoder.weight))
        self.reduce_parts(d_emb, d_emb.shape[0], self._grad(model.image_encoder.bias))
        params = [p for p in model.parameters() if p.requires_grad]
        la = torch.stack([task_loss.sum() / float(B * NQ), label_lam.mean()])
        self._allreduce_finish(params, la, None)
        res["loss_acc"] = la
        return res
