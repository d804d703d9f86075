"""LoggedModule mirror (reference ovr/modeling/logged_module.py:8-42).

Interface kept: ``log(name, tensor)``, ``log_dict(d)``, the ``log_info`` dict the meta-architecture
prints on a NaN loss (distill_prop_mmss_gcnn.py:444-449), ``_log_print`` and ``_log_raise_nan``.
Difference by design: the reference copies EVERY logged tensor to the host and
issues four scalar syncs per call; here ``log`` only records a detached reference and the statistics
are computed on the device when ``log_info`` is actually read, so the fast path never synchronises.
"""
import torch
import torch.distributed as dist
from torch import nn
from torch.nn import functional as F


def stats(tensor):
    t = tensor.detach()
    if t.is_cuda and t.numel() > 0:
        from .. import ops
        packed = ops.tensor_stats(t).tolist()           # one launch (loco_tensor_stats), one 16-byte copy when the entry is READ
    else:
        f = t.to(torch.float32)
        packed = torch.stack([f.min(), f.max(), f.mean(), f.std() if f.numel() > 1 else f.new_zeros(())]).tolist()
    return {"device": t.device.index, "shape": t.shape, "min": packed[0], "max": packed[1], "mean": packed[2],
            "std": packed[3]}


class _LazyLog(dict):
    """dict whose tensor-statistics entries are materialised on first read."""

    def __getitem__(self, k):
        v = dict.__getitem__(self, k)
        if isinstance(v, _Pending):
            v = stats(v.tensor)
            dict.__setitem__(self, k, v)
        return v

    def items(self):
        return [(k, self[k]) for k in dict.keys(self)]

    def values(self):
        return [self[k] for k in dict.keys(self)]

    def get(self, k, default=None):
        return self[k] if k in self else default

    def __repr__(self):
        return repr(dict(self.items()))


class _Pending:
    __slots__ = ("tensor",)

    def __init__(self, tensor):
        self.tensor = tensor


class LoggedModule(nn.Module):
    def __init__(self):
        super().__init__()
        self.log_info = _LazyLog()
        self._log_print = False
        self._log_raise_nan = False

    def _rank(self):
        return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0

    def log(self, name, tensor):
        self.log_info[name] = _Pending(tensor.detach())
        if self._log_print:
            print(f"RANK {self._rank()}: {name}", self.log_info[name])
        if self._log_raise_nan and torch.isnan(tensor).any():
            raise ValueError()

    def log_dict(self, d):
        self.log_info.update(d)
        if self._log_print:
            print(f"RANK {self._rank()}: {d}")
        if self._log_raise_nan:
            for v in d.values():
                if torch.isnan(v).any():
                    raise ValueError()


def normalize_vec(vec_tensor, dim=1):
    return F.normalize(vec_tensor, p=2, dim=dim)                      # logged_module.py:55-65


def standardize_vec(vec_tensor, dim=1):                               # logged_module.py:68-72
    return (vec_tensor - vec_tensor.mean(dim, keepdim=True)) / (vec_tensor.std(dim, keepdim=True) + 1e-12)
