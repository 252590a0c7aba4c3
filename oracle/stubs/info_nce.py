"""No-op stand-in: the reference imports InfoNCE (model.py:14, trainer.py:13) but the in-scope
hot path never calls it.  Oracle test infrastructure only."""


class InfoNCE:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        raise NotImplementedError("info_nce stub: contrastive models are out of scope")
