"""Name-keyed registries — the reference's plugin API for this path.

The reference registers its heads in Detectron2 registries (``ROI_HEADS_REGISTRY``,
roi_emb_heads.py:121,309) and in its own ``MMSS_HEADS_REGISTRY = Registry("MMSS_HEADS")``
(grounding_head.py:10,50).  Detectron2 is not a dependency of this package: when it is importable the
drop-in classes are ALSO registered in its ROI_HEADS_REGISTRY (``register_with_detectron2``) so that
``cfg.MODEL.ROI_HEADS.NAME`` resolves to them; otherwise the stand-in below keeps the same
``register()`` / ``get()`` / ``in`` behaviour.
"""


class Registry(dict):
    def __init__(self, name):
        super().__init__()
        self._name = name

    def register(self, obj=None, *, name=None):
        if obj is None:
            def deco(o):
                self._do_register(name or o.__name__, o)
                return o
            return deco
        self._do_register(name or obj.__name__, obj)
        return obj

    def _do_register(self, name, obj):
        assert name not in self, f"An object named '{name}' was already registered in '{self._name}' registry!"
        self[name] = obj

    def get(self, name):
        if name not in self:
            raise KeyError(f"No object named '{name}' found in '{self._name}' registry!")
        return self[name]


MMSS_HEADS_REGISTRY = Registry("MMSS_HEADS")
ROI_HEADS_REGISTRY = Registry("ROI_HEADS")
BOX_PREDICTORS = Registry("BOX_EMBEDDING_PREDICTORS")


def register_with_detectron2(replace=True):
    """Register the drop-in ROI heads in Detectron2's ROI_HEADS_REGISTRY (no-op + False when absent)."""
    try:
        from detectron2.modeling.roi_heads import ROI_HEADS_REGISTRY as D2
    except Exception:
        return False
    for name, cls in ROI_HEADS_REGISTRY.items():
        if name in D2._obj_map:
            if not replace:
                continue
            del D2._obj_map[name]
        D2._do_register(name, cls)
    return True
