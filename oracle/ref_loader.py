"""ORACLE (test infrastructure only): import the REAL reference GroundingHead from /root/reference.

Only usable in the build container (the GPU box has no /root/reference).  Used by
tests/golden/make_golden.py to generate the committed golden vectors and by
tests/test_oracle_lsm.py (skipped when the reference tree is absent) to re-validate the restatement.

The reference module (/root/reference/ovr/modeling/mmss_heads/grounding_head.py) needs only two
Detectron2 symbols, which are stubbed here:
  * detectron2.utils.registry.Registry       (grounding_head.py:8,10,50)
  * detectron2.utils.events.get_event_storage (logged_module.py:4; never called on this path)
``ovr/__init__.py`` is bypassed with empty namespace packages because it imports every meta-arch
(and hence all of Detectron2).  The hard-coded ``.to("cuda")`` calls (grounding_head.py:43,309-377)
are neutralised on CPU by a context manager that maps "cuda" -> the tensor's own device.
"""
import contextlib
import importlib.util
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("LOCOV_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "ovr/modeling/mmss_heads/grounding_head.py"))


class _Registry(dict):
    def __init__(self, name):
        super().__init__()
        self._name = name

    def register(self, obj=None):
        if obj is None:
            def deco(o):
                self[o.__name__] = o
                return o
            return deco
        self[obj.__name__] = obj
        return obj

    def get(self, name):
        return self[name]


def _stub_modules():
    mods = {}
    for name in ["detectron2", "detectron2.utils", "detectron2.utils.registry", "detectron2.utils.events"]:
        mods[name] = types.ModuleType(name)
    mods["detectron2.utils.registry"].Registry = _Registry
    mods["detectron2.utils.events"].get_event_storage = lambda: None
    for name in ["ovr", "ovr.modeling", "ovr.modeling.mmss_heads"]:
        m = types.ModuleType(name)
        m.__path__ = []
        mods[name] = m
    return mods


def load_reference_grounding_head():
    """Returns the reference's ``GroundingHead`` class (unmodified source, stubbed imports)."""
    if not reference_available():
        raise FileNotFoundError(f"reference tree not found at {REFERENCE_ROOT}")
    saved = {k: sys.modules.get(k) for k in _stub_modules()}
    sys.modules.update(_stub_modules())
    try:
        def _imp(modname, rel):
            spec = importlib.util.spec_from_file_location(modname, os.path.join(REFERENCE_ROOT, rel))
            mod = importlib.util.module_from_spec(spec)
            sys.modules[modname] = mod
            spec.loader.exec_module(mod)
            return mod
        _imp("ovr.modeling.logged_module", "ovr/modeling/logged_module.py")
        gh = _imp("ovr.modeling.mmss_heads.grounding_head", "ovr/modeling/mmss_heads/grounding_head.py")
        return gh.GroundingHead
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        sys.modules.pop("ovr.modeling.logged_module", None)
        sys.modules.pop("ovr.modeling.mmss_heads.grounding_head", None)


@contextlib.contextmanager
def cuda_to_cpu():
    """Make ``tensor.to("cuda")`` a no-op so the reference's hard-coded device strings run on CPU."""
    orig = torch.Tensor.to

    def to(self, *args, **kwargs):
        if args and isinstance(args[0], str) and args[0].startswith("cuda") and not torch.cuda.is_available():
            args = (self.device,) + tuple(args[1:])
        return orig(self, *args, **kwargs)

    torch.Tensor.to = to
    try:
        yield
    finally:
        torch.Tensor.to = orig


class _Cfg(dict):
    __getattr__ = dict.__getitem__


def make_grounding_cfg(alignment="softmax", temperature=10.0, loss="cross_entropy", align_words=True,
                       align_regions=True, distillation=True, text_input="input_embeddings",
                       local_metric="dot", global_metric="aligned_local", negative_mining="hardest",
                       margin=1.0):
    """Attribute-dict with the keys GroundingHead.__init__ reads (grounding_head.py:54-90;
    defaults /root/reference/ovr/config/config.py:51-63, coco_lsm.yaml:64-75)."""
    g = _Cfg(LOCAL_METRIC=local_metric, GLOBAL_METRIC=global_metric, ALIGNMENT=alignment,
             ALIGNMENT_TEMPERATURE=temperature, LOSS=loss, NEGATIVE_MINING=negative_mining,
             TRIPLET_MARGIN=margin, ALIGN_WORDS_TO_REGIONS=align_words, ALIGN_REGIONS_TO_WORDS=align_regions,
             TEXT_INPUT=text_input)
    return _Cfg(MODEL=_Cfg(MMSS_HEAD=_Cfg(GROUNDING=g, DISTILLATION_LOSS=distillation)))
