"""No-op stand-in (reference imports it at model.py:16; never called on the hot path)."""


def pairwise_cosine_similarity(*a, **k):
    raise NotImplementedError("torchmetrics stub: DOSE models are out of scope")
