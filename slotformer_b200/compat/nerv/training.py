"""``nerv.training`` subset: BaseModel / BaseParams (see package docstring)."""
import torch
from torch import nn


class BaseParams:
    """Config base class: plain class attributes + dict-like ``get``.

    Call sites: base_slots/datasets/clevrer.py:376 (``params.get(k, d)``),
    scripts/train.py:103 (run-time attribute assignment).
    """

    def get(self, key, default=None):
        return getattr(self, key, default)

    def to_dict(self):
        out = {}
        for klass in reversed(type(self).__mro__):
            out.update({k: v for k, v in vars(klass).items()
                        if not k.startswith('_') and not callable(v)})
        out.update(vars(self))
        return out

    def __repr__(self):
        body = ', '.join(f'{k}={v!r}' for k, v in sorted(self.to_dict().items()))
        return f'{type(self).__name__}({body})'


class BaseModel(nn.Module):
    """``nn.Module`` with the loss hooks the reference trainers call."""

    def calc_train_loss(self, data_dict, out_dict):
        raise NotImplementedError

    @torch.no_grad()
    def calc_eval_loss(self, data_dict, out_dict):
        return self.calc_train_loss(data_dict, out_dict)

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    @property
    def device(self):
        return next(self.parameters()).device
