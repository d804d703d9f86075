"""Minimal stand-ins for the Detectron2 structures the heads receive (``Boxes``, ``Instances``,
``ShapeSpec``).  The drop-in modules are duck-typed: real Detectron2 objects work unchanged; these are
used by the tests/bench and when Detectron2 is not installed."""
from collections import namedtuple

import torch

ShapeSpec = namedtuple("ShapeSpec", ["channels", "height", "width", "stride"], defaults=[None, None, None, None])


class Boxes:
    def __init__(self, tensor):
        tensor = torch.as_tensor(tensor, dtype=torch.float32)
        if tensor.numel() == 0:
            tensor = tensor.reshape(-1, 4)
        self.tensor = tensor

    def __len__(self):
        return self.tensor.shape[0]

    def __getitem__(self, item):
        t = self.tensor[item]
        return Boxes(t.reshape(-1, 4))

    def to(self, *a, **k):
        return Boxes(self.tensor.to(*a, **k))

    @property
    def device(self):
        return self.tensor.device

    def clip(self, box_size):
        h, w = box_size
        self.tensor[:, 0].clamp_(min=0, max=w)
        self.tensor[:, 1].clamp_(min=0, max=h)
        self.tensor[:, 2].clamp_(min=0, max=w)
        self.tensor[:, 3].clamp_(min=0, max=h)

    @staticmethod
    def cat(boxes_list):
        if len(boxes_list) == 0:
            return Boxes(torch.empty(0, 4))
        return Boxes(torch.cat([b.tensor for b in boxes_list], 0))


class Instances:
    def __init__(self, image_size, **fields):
        object.__setattr__(self, "_image_size", tuple(image_size))
        object.__setattr__(self, "_fields", {})
        for k, v in fields.items():
            self.set(k, v)

    @property
    def image_size(self):
        return self._image_size

    def __setattr__(self, name, val):
        if name.startswith("_"):
            object.__setattr__(self, name, val)
        else:
            self.set(name, val)

    def __getattr__(self, name):
        fields = object.__getattribute__(self, "_fields")
        if name not in fields:
            raise AttributeError(f"Cannot find field '{name}' in the given Instances!")
        return fields[name]

    def set(self, name, value):
        self._fields[name] = value

    def has(self, name):
        return name in self._fields

    def get(self, name):
        return self._fields[name]

    def get_fields(self):
        return self._fields

    def __len__(self):
        for v in self._fields.values():
            return len(v)
        return 0

    def __getitem__(self, item):
        ret = Instances(self._image_size)
        for k, v in self._fields.items():
            ret.set(k, v[item])
        return ret
