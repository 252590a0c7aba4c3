"""Restatement of the `info_nce` dependency (PyPI `info-nce-pytorch`, unpinned in the reference: README.md lists no
version; imported at model.py:14 and trainer.py:13, called at model.py:206-214 (SGL.cal_loss) and :322-332 (HALF)).
ORACLE TEST INFRASTRUCTURE ONLY -- the package is absent from this image, so its published algorithm is restated:

    info_nce(query, positive_key, negative_keys, temperature=0.1, reduction='mean', negative_mode='unpaired'):
        q, p, n   = F.normalize(., dim=-1) of each                      (x / max(||x||_2, 1e-12))
        pos_logit = sum(q * p, dim=1, keepdim=True)                      [B, 1]
        neg_logit = q @ n^T                          ('unpaired')        [B, M]
        logits    = cat([pos_logit, neg_logit], dim=1); labels = 0
        return F.cross_entropy(logits / temperature, labels, reduction)

With no explicit negatives the other rows of `positive_key` serve as negatives (logits = q @ p^T, labels = arange).
The reference only uses the 'unpaired' form with negative_keys == positive_key (the positive appears twice in the
softmax denominator) and the default temperature 0.1 -- its `taugh` config key is never passed on.
"""
import torch
import torch.nn.functional as F
from torch import nn


def _normalize(*xs):
    return [None if x is None else F.normalize(x, dim=-1) for x in xs]


def info_nce(query, positive_key, negative_keys=None, temperature=0.1, reduction='mean', negative_mode='unpaired'):
    if query.dim() != 2 or positive_key.dim() != 2:
        raise ValueError('<query> and <positive_key> must have 2 dimensions.')
    if len(query) != len(positive_key):
        raise ValueError('<query> and <positive_key> must have the same number of samples.')
    query, positive_key, negative_keys = _normalize(query, positive_key, negative_keys)
    if negative_keys is not None:
        positive_logit = torch.sum(query * positive_key, dim=1, keepdim=True)
        if negative_mode == 'unpaired':
            negative_logits = query @ negative_keys.transpose(-2, -1)
        elif negative_mode == 'paired':
            negative_logits = (query.unsqueeze(1) @ negative_keys.transpose(-2, -1)).squeeze(1)
        else:
            raise ValueError(negative_mode)
        logits = torch.cat([positive_logit, negative_logits], dim=1)
        labels = torch.zeros(len(logits), dtype=torch.long, device=query.device)
    else:
        logits = query @ positive_key.transpose(-2, -1)
        labels = torch.arange(len(query), device=query.device)
    return F.cross_entropy(logits / temperature, labels, reduction=reduction)


class InfoNCE(nn.Module):
    def __init__(self, temperature=0.1, reduction='mean', negative_mode='unpaired'):
        super().__init__()
        self.temperature = temperature
        self.reduction = reduction
        self.negative_mode = negative_mode

    def forward(self, query, positive_key, negative_keys=None):
        return info_nce(query, positive_key, negative_keys, temperature=self.temperature, reduction=self.reduction,
                        negative_mode=self.negative_mode)
