"""Host front-end of the native episodic task sampler (fumi_sampler_* in the C ABI).

Replaces the reference loader stack for the path (fumi/dataset/data.py:73-84,125-188 +
torchmeta 1.7.0): the class tables are built once, and each meta-batch is a set of index
arrays (image ids into the HBM feature bank + labels) instead of collated feature tensors.

The reference draws from three global generators.  To stay a drop-in,
the native sampler *borrows* the live state of Python's ``random`` and of the torch CPU
generator for the duration of a call and writes it back, so a host program that seeds them as
fumi/main.py:51-53 does gets the reference's task stream.
"""
import ctypes as C
import random

import numpy as np
import torch

from . import _lib


def _torch_state_get():
    s = torch.get_rng_state().numpy()
    st = np.empty(626, np.uint32)
    st[:624] = s[24:24 + 624 * 8].view(np.uint64).astype(np.uint32)
    st[624] = np.uint32(s[16:24].view(np.uint64)[0])      # next
    st[625] = np.uint32(s[8:12].view(np.int32)[0])        # left
    return st, s


def _torch_state_set(st, raw):
    raw = raw.copy()
    raw[24:24 + 624 * 8].view(np.uint64)[:] = st[:624].astype(np.uint64)
    raw[16:24].view(np.uint64)[0] = np.uint64(st[624])
    raw[8:12].view(np.int32)[0] = np.int32(st[625])
    torch.set_rng_state(torch.from_numpy(raw))


def class_tables(cat_of, categories):
    """(offsets[C+1], ids) -- per split-class ascending image ids (data.py:395-414)."""
    cat_of = np.asarray(cat_of)
    order = np.argsort(cat_of, kind="stable")
    sorted_cat = cat_of[order]
    starts = np.searchsorted(sorted_cat, categories, side="left")
    ends = np.searchsorted(sorted_cat, categories, side="right")
    sizes = ends - starts
    offsets = np.zeros(len(categories) + 1, np.int64)
    np.cumsum(sizes, out=offsets[1:])
    ids = np.concatenate([order[s:e] for s, e in zip(starts, ends)]).astype(np.int64)
    return offsets, ids


class EpisodeSampler:
    """One split's loader: ``next_batch(B)`` -> dict of int64 index arrays (host, pinned if asked)."""

    def __init__(self, cat_of, categories, num_ways, num_shots, num_query, num_threads=0):
        self.categories = np.asarray(categories, np.int64)
        self.N, self.K, self.Q = int(num_ways), int(num_shots), int(num_query)
        self.offsets, self.ids = class_tables(cat_of, self.categories)
        self.num_threads = num_threads
        h = C.c_void_p()
        _lib.check(_lib.lib().fumi_sampler_create(_lib.ptr(self.offsets), _lib.ptr(self.ids),
                                                  len(self.categories), self.N, self.K, self.Q, C.byref(h)),
                   "fumi_sampler_create")
        self._h = h

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        try:
            if h:
                _lib.lib().fumi_sampler_destroy(h)
        except Exception:      # interpreter shutdown
            pass

    def new_iterator(self):
        """Equivalent of iter(loader): the DataLoader draws its base seed from the torch stream."""
        st, raw = _torch_state_get()
        _lib.check(_lib.lib().fumi_sampler_new_iterator(self._h, _lib.ptr(st)), "fumi_sampler_new_iterator")
        _torch_state_set(st, raw)

    def empty_out(self, batch_size):
        B, N, K, Q = int(batch_size), self.N, self.K, self.Q
        return {k: np.empty((B, n), np.int64) for k, n in
                (("classes", N), ("label_perm", N), ("sup_ids", N * K), ("qry_ids", N * Q),
                 ("sup_y", N * K), ("qry_y", N * Q), ("head_class", N),
                 ("sup_rows", N * K), ("qry_rows", N * Q))}

    def next_batch_states(self, batch_size, py_state, torch_state, out):
        """One meta-batch on explicit generator states (uint32[625] / uint32[626], advanced in place).
        Touches no global state: safe to call from a prefetch thread (the C call releases the GIL)."""
        try:
            _lib.check(_lib.lib().fumi_sampler_next(
                self._h, int(batch_size), _lib.ptr(py_state), _lib.ptr(torch_state), *[_lib.ptr(out[k]) for k in
                ("classes", "label_perm", "sup_ids", "qry_ids", "sup_y", "qry_y", "head_class",
                 "sup_rows", "qry_rows")],
                self.num_threads), "fumi_sampler_next")
        except _lib.FumiError as e:
            if "smaller than the minimum" in str(e):
                raise ValueError(str(e).split(": ", 2)[-1]) from None     # torchmeta raises ValueError
            raise
        return out

    # ---- device-resident form: sequential streams on the host, permutations on the GPU ----
    def empty_plan(self, batch_size, pin_memory=False):
        """Host staging tensors of one plan (fumi_sampler_plan outputs)."""
        B, N, KQ = int(batch_size), self.N, self.K + self.Q
        mk = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=pin_memory)
        return {"classes": mk((B, N), torch.int64), "label_perm": mk((B, N), torch.int64),
                "head_class": mk((B, N), torch.int64), "perm_seed": mk((B, N), torch.int32),   # uint32 bits
                "picks": mk((B, N, KQ), torch.int32), "job_order": mk((B, N), torch.int32)}

    def plan_states(self, batch_size, py_state, torch_state, plan):
        """Sequential generator streams of one meta-batch on explicit states (thread-safe like
        next_batch_states); the hash-seeded permutations are left to expand()."""
        try:
            _lib.check(_lib.lib().fumi_sampler_plan(
                self._h, int(batch_size), _lib.ptr(py_state), _lib.ptr(torch_state),
                *[_lib.ptr(plan[k]) for k in ("classes", "label_perm", "head_class", "perm_seed", "picks", "job_order")]),
                "fumi_sampler_plan")
        except _lib.FumiError as e:
            if "smaller than the minimum" in str(e):
                raise ValueError(str(e).split(": ", 2)[-1]) from None
            raise
        return plan

    def plan(self, batch_size, plan=None, pin_memory=False):
        if plan is None:
            plan = self.empty_plan(batch_size, pin_memory)
        ver, key, gauss = random.getstate()
        py = np.asarray(key, np.uint32)
        st, raw = _torch_state_get()
        self.plan_states(batch_size, py, st, plan)
        random.setstate((ver, tuple(int(x) for x in py), gauss))
        _torch_state_set(st, raw)
        return plan

    def device_tables(self, device):
        device = torch.device(device)
        tabs = getattr(self, "_dev_tables", None)
        if tabs is None or tabs[0].device != device:
            tabs = (torch.from_numpy(self.offsets).to(device), torch.from_numpy(self.ids).to(device))
            self._dev_tables = tabs
        return tabs

    def expand(self, plan, device):
        """Upload a plan and run the per (task, class) permutations on `device` (current stream).
        Returns dict(sup_ids, qry_ids, sup_y, qry_y, sup_rows, qry_rows, head_class, classes, label_perm)
        of device tensors -- bit-identical to next_batch()."""
        device = torch.device(device)
        _lib.require_cuda(device, "EpisodeSampler.expand (use next_batch for host-resident indices)")
        offsets, ids = self.device_tables(device)
        d = {k: v.to(device, non_blocking=True) for k, v in plan.items()}
        B, N, K, Q = d["classes"].shape[0], self.N, self.K, self.Q
        out = {k: torch.empty((B, N * w), dtype=torch.int64, device=device)
               for k, w in (("sup_ids", K), ("qry_ids", Q), ("sup_y", K), ("qry_y", Q), ("sup_rows", K), ("qry_rows", Q))}
        stream = _lib.stream_ptr(device) if device.type == "cuda" else None
        _lib.check(_lib.lib().fumi_sampler_expand(
            _lib.ptr(offsets), _lib.ptr(ids), int(np.diff(self.offsets).max()), _lib.ptr(d["classes"]),
            _lib.ptr(d["label_perm"]), _lib.ptr(d["perm_seed"]), _lib.ptr(d["picks"]), _lib.ptr(d.get("job_order")), B, N, K, Q,
            *[_lib.ptr(out[k]) for k in ("sup_ids", "qry_ids", "sup_y", "qry_y", "sup_rows", "qry_rows")], stream),
            "fumi_sampler_expand")
        out.update(head_class=d["head_class"], classes=d["classes"], label_perm=d["label_perm"])
        return out

    def next_batch(self, batch_size, out=None):
        if out is None:
            out = self.empty_out(batch_size)
        ver, key, gauss = random.getstate()
        py = np.asarray(key, np.uint32)
        st, raw = _torch_state_get()
        self.next_batch_states(batch_size, py, st, out)
        random.setstate((ver, tuple(int(x) for x in py), gauss))
        _torch_state_set(st, raw)
        return out
