"""GPU-resident feature bank and index-only episode batches.

The reference loader copies every sampled 2048-d feature row through Python, clones the class
description per sample and collates [B, NK, D] tensors on the host (dataset/data.py:533-581,
torchmeta collate).  Here the split's image features live in HBM once (rows grouped by class, in
the order of the sampler's class table), descriptions are one row per class, and a meta-batch is
just the sampled index arrays.
"""
from dataclasses import dataclass, field
from typing import Optional

import numpy as np
import torch


@dataclass
class FeatureBank:
    feats: torch.Tensor                 # f32 [R, D]  split feature matrix, row r = image ids[r]
    text: torch.Tensor                  # f32 [C, T]  description embedding of split-class c
    ids: np.ndarray                     # i64 [R]     image id of each row (dataset/data.py image ids)
    categories: np.ndarray              # i64 [C]     dataset category of split-class c

    @property
    def device(self):
        return self.feats.device


@dataclass
class EpisodeBatch:
    """Index form of one meta-batch (what the reference's batch dict carries as dense tensors)."""
    bank: Optional[FeatureBank]
    sup_rows: torch.Tensor              # i64 [B, NK] rows of bank.feats
    qry_rows: torch.Tensor              # i64 [B, NQ]
    sup_y: torch.Tensor                 # i64 [B, NK] labels (batch['train'][1])
    qry_y: torch.Tensor                 # i64 [B, NQ] labels (batch['test'][1])
    sup_ids: object = None              # i64 [B, NK] image ids (batch['train'][0][0])
    qry_ids: object = None
    head_class: Optional[torch.Tensor] = None   # i64 [B, N] split-class whose description conditions label i
    host: dict = field(default_factory=dict)

    def to(self, device):
        device = torch.device(device)

        def mv(t):
            if t is None:
                return None
            if isinstance(t, np.ndarray):
                t = torch.from_numpy(t)
            return t.to(device, non_blocking=True).contiguous()
        if isinstance(self.sup_rows, torch.Tensor) and self.sup_rows.device == device:
            return self
        return EpisodeBatch(bank=self.bank, sup_rows=mv(self.sup_rows), qry_rows=mv(self.qry_rows),
                            sup_y=mv(self.sup_y), qry_y=mv(self.qry_y), sup_ids=self.sup_ids, qry_ids=self.qry_ids,
                            head_class=mv(self.head_class), host=self.host)
