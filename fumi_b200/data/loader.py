"""Episodic loaders over a GPU-resident feature bank: drop-in for dataset.data.get_dataset.

Reference: fumi/dataset/data.py:25-86 (get_dataset -> three BatchMetaDataLoader) and 125-188
(get_inat_anim: train split with Q = --num_shots_test, val/test with Q = int(100 / N), each
ClassSplitter seeded 0).  A loader here has the same iteration contract -- ``iter(loader)`` starts a
new DataLoader iterator (which burns the torch base seed), ``next`` yields one meta-batch, forever --
but the batch is an EpisodeBatch of indices into the split's FeatureBank instead of collated tensors.
"""
import json
import os
import queue
import random
import threading

import numpy as np
import torch

from .bank import EpisodeBatch, FeatureBank
from .synth import INAT_ANIM_CLASSES, INAT_ANIM_IMAGES, class_split, make_bank
from ..sampler import EpisodeSampler

_KEYS = ("classes", "label_perm", "sup_ids", "qry_ids", "sup_y", "qry_y", "head_class", "sup_rows", "qry_rows")


class EpisodeLoader:
    """prefetch = n > 0: a background thread samples up to n meta-batches ahead on private copies of the two
    generator states (snapshotted when the iterator is created); when a batch is handed out, the global
    `random` / torch generator states are set to what they were right after that batch was drawn, so the
    host program observes the same generator sequence as with the synchronous loader."""

    def __init__(self, bank, sampler, batch_size, pin_memory=True, prefetch=0, device_sampler=None, shard=None):
        self.bank, self.sampler, self.batch_size = bank, sampler, int(batch_size)
        # shard = (rank, world): every rank draws the SAME global stream of `world * batch_size` tasks per step (so the
        # sampled tasks are bit-identical to a single-process run with that meta-batch) and keeps tasks
        # [rank * batch_size, (rank + 1) * batch_size) -- SURVEY.md 8(e).  The host-side plan then costs `world` times
        # more per step (5.5 ms per 4096 tasks); the default (shard=None) lets each rank draw its own stream.
        self.shard = None if shard is None else (int(shard[0]), int(shard[1]))
        self._draw = self.batch_size * (self.shard[1] if self.shard else 1)
        self.pin = bool(pin_memory) and bank.feats.is_cuda
        self.prefetch = int(prefetch)
        # device_sampler: the host only advances the sequential generator streams (fumi_sampler_plan); the
        # per (task, class) permutations and the index arrays are produced in HBM (fumi_sampler_expand).
        # Default on a CUDA bank; off = the all-host sampler (fumi_sampler_next) + pinned index buffers.
        self.device_sampler = bank.feats.is_cuda if device_sampler is None else bool(device_sampler)
        self.dataset = sampler          # len(loader.dataset) parity is not meaningful for episodes
        self._stop = None

    def set_shard(self, rank, world):
        """Switch to the global-stream mode (see `shard` above) before the first iterator is created."""
        self.shard = (int(rank), int(world))
        self._draw = self.batch_size * self.shard[1]

    def _mine(self, t):
        """This rank's slice of a [world * batch_size, ...] array / tensor."""
        if self.shard is None:
            return t
        r, B = self.shard[0], self.batch_size
        return t[r * B:(r + 1) * B]

    def _host_buffers(self):
        N, K, Q, B = self.sampler.N, self.sampler.K, self.sampler.Q, self._draw
        widths = dict(classes=N, label_perm=N, sup_ids=N * K, qry_ids=N * Q, sup_y=N * K, qry_y=N * Q,
                      head_class=N, sup_rows=N * K, qry_rows=N * Q)
        ts = {k: torch.empty((B, w), dtype=torch.int64, pin_memory=self.pin) for k, w in widths.items()}
        return ts, {k: t.numpy() for k, t in ts.items()}

    def _expand(self, plan):
        plan = {k: self._mine(v) for k, v in plan.items() if not (self.shard and k == "job_order")}   # global job ids
        d = self.sampler.expand(plan, self.bank.feats.device)
        host = {k: plan[k].numpy().copy() for k in ("classes", "label_perm", "head_class")}   # the plan buffer is reused
        return EpisodeBatch(bank=self.bank, sup_rows=d["sup_rows"], qry_rows=d["qry_rows"], sup_y=d["sup_y"],
                            qry_y=d["qry_y"], sup_ids=d["sup_ids"], qry_ids=d["qry_ids"],
                            head_class=d["head_class"], host=host)

    def next_batch(self):
        if self.device_sampler:
            return self._expand(self.sampler.plan(self._draw, pin_memory=self.pin))
        ts, arrs = self._host_buffers()
        self.sampler.next_batch(self._draw, out=arrs)
        return self._make(ts, arrs)

    def _make(self, ts, arrs):
        ts = {k: self._mine(v) for k, v in ts.items()}
        arrs = {k: self._mine(v) for k, v in arrs.items()}
        return EpisodeBatch(bank=self.bank, sup_rows=ts["sup_rows"], qry_rows=ts["qry_rows"], sup_y=ts["sup_y"],
                            qry_y=ts["qry_y"], sup_ids=arrs["sup_ids"], qry_ids=arrs["qry_ids"],
                            head_class=ts["head_class"], host=arrs)

    def close(self):
        """Stop the prefetch thread and wait for it (a thread still inside a CUDA call at interpreter exit aborts the
        process)."""
        if self._stop is not None:
            self._stop.set()
            self._stop = None
        th, self._thread = getattr(self, "_thread", None), None
        if th is not None and th.is_alive() and th is not threading.current_thread():
            th.join(timeout=5.0)

    def __iter__(self):
        self.sampler.new_iterator()
        if self.prefetch <= 0:
            while True:
                yield self.next_batch()
        from ..sampler import _torch_state_get, _torch_state_set
        self.close()
        ver, key, gauss = random.getstate()
        py = np.asarray(key, np.uint32)
        st, raw = _torch_state_get()
        q = queue.Queue(maxsize=self.prefetch)
        stop = self._stop = threading.Event()

        # device sampler on a CUDA bank: the plan upload and fumi_sampler_expand run on the loader's own stream from
        # the prefetch thread, so they overlap the previous batch's kernels; the consumer's stream waits on an event.
        dev = self.bank.feats.device
        side = None
        if self.device_sampler and dev.type == "cuda":       # one loader stream (and allocator pool) per loader, not per epoch
            if getattr(self, "_side", None) is None:
                self._side = torch.cuda.Stream(dev)
            side = self._side

        # pinned plan buffers are allocated once and reused round-robin: a pinned allocation in steady state
        # (cudaHostAlloc) stalls every CUDA call of the process for milliseconds.  A buffer is rewritten prefetch + 2
        # batches after its upload was enqueued, long after the consumer waited on that batch's event.
        ring = [self.sampler.empty_plan(self._draw, pin_memory=self.pin) for _ in range(self.prefetch + 2)] \
            if self.device_sampler else []
        if side is not None:
            # Same for device memory: a batch's index arrays are allocated on the loader stream and return to its pool
            # only when the consumer's stream has passed them, so the pool keeps growing (cudaMalloc from this thread,
            # which stalls the consumer's launches for tens of milliseconds while the device is busy) until it holds
            # every batch in flight.  Fill it once, up front: prefetch + 6 sets of exactly the tensors of a batch.
            N, K, Q, B = self.sampler.N, self.sampler.K, self.sampler.Q, self.batch_size
            with torch.cuda.stream(side):
                warm = []
                for _ in range(self.prefetch + 6):
                    warm.append([t.to(dev, non_blocking=True) for t in self.sampler.empty_plan(B, pin_memory=False).values()])
                    warm.append([torch.empty((B, N * w_), dtype=torch.int64, device=dev) for w_ in (K, Q, K, Q, K, Q)])
                del warm
            side.synchronize()

        def worker():
            try:
                if side is not None:
                    torch.cuda.set_device(dev)
                produced = 0
                while not stop.is_set():
                    if self.device_sampler:
                        plan = ring[produced % len(ring)]
                        produced += 1
                        self.sampler.plan_states(self._draw, py, st, plan)
                        if side is not None:
                            with torch.cuda.stream(side):
                                batch = self._expand(plan)
                                ev = torch.cuda.Event()
                                ev.record(side)
                            item = (batch, ev, py.copy(), st.copy())
                        else:
                            item = (plan, None, py.copy(), st.copy())
                    else:
                        ts, arrs = self._host_buffers()
                        self.sampler.next_batch_states(self._draw, py, st, arrs)
                        item = (ts, arrs, py.copy(), st.copy())
                    while not stop.is_set():
                        try:
                            q.put(item, timeout=0.1)
                            break
                        except queue.Full:
                            pass
            except Exception as e:          # surfaced on the consumer side
                q.put(e)

        self._thread = threading.Thread(target=worker, daemon=True)
        self._thread.start()
        try:
            while True:
                item = q.get()
                if isinstance(item, Exception):
                    raise item
                ts, arrs, py_after, st_after = item
                random.setstate((ver, tuple(py_after.tolist()), gauss))       # (tolist: C speed; 625 words every batch)
                _torch_state_set(st_after, raw)
                if isinstance(ts, EpisodeBatch):                 # expanded on the loader stream
                    cur = torch.cuda.current_stream(dev)
                    cur.wait_event(arrs)
                    for t in (ts.sup_rows, ts.qry_rows, ts.sup_y, ts.qry_y, ts.sup_ids, ts.qry_ids, ts.head_class):
                        t.record_stream(cur)
                    yield ts
                else:
                    yield self._expand(ts) if arrs is None else self._make(ts, arrs)
        finally:
            stop.set()


def build_loaders(feats, text, cat_of, num_ways, num_shots, num_shots_test, batch_size, device, num_threads=0):
    """(train, val, test) EpisodeLoaders.  feats [M,D] / text [C,T] / cat_of [M] are host arrays."""
    C = text.shape[0]
    loaders = []
    for split, cats in zip(("train", "val", "test"), class_split(C)):
        # each InatAnim constructor reseeds the three global generators (data.py:320-322)
        random.seed(0); np.random.seed(0); torch.manual_seed(0)
        Q = num_shots_test if split == "train" else int(100 / num_ways)
        sampler = EpisodeSampler(cat_of, cats, num_ways, num_shots, Q, num_threads=num_threads)
        bank = FeatureBank(feats=torch.from_numpy(np.ascontiguousarray(feats[sampler.ids])).to(device),
                           text=torch.from_numpy(np.ascontiguousarray(text[cats])).to(device),
                           ids=sampler.ids, categories=np.asarray(cats))
        loaders.append(EpisodeLoader(bank, sampler, batch_size))
    return tuple(loaders)


def load_arrays(args):
    """Host arrays (feats, text, cat_of) from --data_dir, or the synthetic banks with --synthetic.

    On-disk layout accepted: <data_dir>/iNat-Anim/bank.npz with arrays feats [M,D] f32, text [C,T] f32
    (description embeddings, precomputed), cat_of [M] i64.  The reference's inat_anim.json +
    image_embeddings_*.hdf5 pair (data.py:373-430) converts to it with `python -m fumi_b200.data.convert`."""
    if getattr(args, "synthetic", False):
        b = make_bank(num_images=int(os.environ.get("FUMI_SYNTH_IMAGES", INAT_ANIM_IMAGES)),
                      num_classes=int(os.environ.get("FUMI_SYNTH_CLASSES", INAT_ANIM_CLASSES)),
                      im_dim=args.im_emb_dim, text_dim=args.text_emb_dim)
        return b.feats, b.text, b.cat_of
    path = os.path.join(args.data_dir, "iNat-Anim", "bank.npz")
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} not found (pass --synthetic for the synthetic iNat-Anim-shaped banks)")
    z = np.load(path)
    return z["feats"], z["text"], z["cat_of"]


def get_dataset(args):
    """dataset.data.get_dataset (data.py:25-86): (train_loader, val_loader, test_loader, dictionary)."""
    if args.dataset != "inat-anim":
        raise NotImplementedError()          # data.py:71; CUB / supervised-inat-anim are outside the path
    if args.text_encoder != "BERT":
        raise NotImplementedError("only precomputed description embeddings (--text_encoder BERT) are built")
    feats, text, cat_of = load_arrays(args)
    bs = args.tasks_per_batch if getattr(args, "tasks_per_batch", None) else args.batch_size
    tl, vl, te = build_loaders(feats, text, cat_of, args.num_ways, args.num_shots, args.num_shots_test, bs, args.device)
    return tl, vl, te, {}
