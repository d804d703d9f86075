"""ORACLE (test infrastructure only): import the REAL reference heads (GroundingHead, EmbeddingFastRCNNOutputLayers,
Embedding[Proposals]Res5ROIHeads) from /root/reference.

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
    """sys.modules entries standing in for Detectron2 / fvcore (restated in oracle/d2_stubs.py) and for the
    ``ovr`` package __init__ files (which import every meta-architecture)."""
    from . import d2_stubs as S
    mods = {}

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__path__ = []
        for k, v in attrs.items():
            setattr(m, k, v)
        mods[name] = m
        return m

    mod("detectron2")
    mod("detectron2.utils")
    mod("detectron2.utils.registry", Registry=_Registry)
    mod("detectron2.utils.events", get_event_storage=lambda: None)
    mod("detectron2.config", configurable=S.configurable)
    mod("detectron2.layers", ShapeSpec=S.ShapeSpec, batched_nms=S.batched_nms, cat=S.cat, cross_entropy=S.cross_entropy,
        nonzero_tuple=S.nonzero_tuple)
    mod("detectron2.structures", Boxes=S.Boxes, Instances=S.Instances, ImageList=None, pairwise_iou=None)
    mod("detectron2.modeling", META_ARCH_REGISTRY=_Registry("META_ARCH"))
    mod("detectron2.modeling.box_regression", Box2BoxTransform=S.Box2BoxTransform)
    mod("detectron2.modeling.roi_heads", ROI_HEADS_REGISTRY=S.ROI_HEADS_REGISTRY, ROIHeads=S.ROIHeads)
    mod("detectron2.modeling.roi_heads.fast_rcnn", fast_rcnn_inference=S.fast_rcnn_inference,
        fast_rcnn_inference_single_image=S.fast_rcnn_inference_single_image, FastRCNNOutputLayers=S.FastRCNNOutputLayers,
        _log_classification_stats=S._log_classification_stats)
    mod("detectron2.modeling.poolers", ROIPooler=S.ROIPooler)
    mod("detectron2.modeling.sampling", subsample_labels=None)
    mod("detectron2.modeling.backbone", build_backbone=None)
    mod("detectron2.modeling.backbone.resnet", BottleneckBlock=S.BottleneckBlock, ResNet=S.ResNet)
    mod("detectron2.modeling.proposal_generator")
    mod("detectron2.modeling.proposal_generator.proposal_utils", add_ground_truth_to_proposals=None)
    mod("fvcore")
    mod("fvcore.nn", giou_loss=S.giou_loss, smooth_l1_loss=S.smooth_l1_loss)
    for name in ["ovr", "ovr.modeling", "ovr.modeling.mmss_heads", "ovr.modeling.roi_heads", "ovr.modeling.meta_arch", "ovr.modeling.language"]:
        mod(name)
    # imported at module level by distill_mmss_gcnn.py:9-14, used only by the meta-architecture class (out of scope)
    mod("ovr.modeling.language.backbone", build_backbone=None)
    mod("ovr.modeling.mmss_heads.mmss_heads", build_mmss_heads=None)
    # multi-token class scoring head: imported by box_emb_head.py:21-23, never instantiated on this path
    mod("ovr.modeling.roi_heads.box_emb_grounding_head", EmbeddingGroundingFastRCNNOutputLayers=type("EmbeddingGroundingFastRCNNOutputLayers", (), {}))
    return mods


_REF_FILES = {
    "ovr.modeling.logged_module": "ovr/modeling/logged_module.py",
    "ovr.modeling.mmss_heads.grounding_head": "ovr/modeling/mmss_heads/grounding_head.py",
    "ovr.modeling.roi_heads.box_emb_head": "ovr/modeling/roi_heads/box_emb_head.py",
    "ovr.modeling.roi_heads.roi_emb_heads": "ovr/modeling/roi_heads/roi_emb_heads.py",
    "ovr.modeling.meta_arch.distill_mmss_gcnn": "ovr/modeling/meta_arch/distill_mmss_gcnn.py",
    "ovr.modeling.roi_heads.box_emb_grounding_head": "ovr/modeling/roi_heads/box_emb_grounding_head.py",
}


def _load_reference_modules(names):
    """Execute the reference's own source files (unmodified) under the stubbed imports; returns {modname: module}."""
    if not reference_available():
        raise FileNotFoundError(f"reference tree not found at {REFERENCE_ROOT}")
    stubs = _stub_modules()
    touched = list(stubs) + list(_REF_FILES)
    saved = {k: sys.modules.get(k) for k in touched}
    sys.modules.update(stubs)
    out = {}
    try:
        for modname in names:
            spec = importlib.util.spec_from_file_location(modname, os.path.join(REFERENCE_ROOT, _REF_FILES[modname]))
            m = importlib.util.module_from_spec(spec)
            sys.modules[modname] = m
            spec.loader.exec_module(m)
            out[modname] = m
        return out
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def load_reference_grounding_head():
    """Returns the reference's ``GroundingHead`` class (unmodified source, stubbed imports)."""
    mods = _load_reference_modules(["ovr.modeling.logged_module", "ovr.modeling.mmss_heads.grounding_head"])
    return mods["ovr.modeling.mmss_heads.grounding_head"].GroundingHead


def load_reference_box_head():
    """Returns the reference's box_emb_head module: ``EmbeddingFastRCNNOutputLayers`` (box_emb_head.py:60-236, its own
    forward / forward_cls_prediction / set_class_embeddings / from_config) on top of the restated Detectron2 base
    class (oracle/d2_stubs.FastRCNNOutputLayers: losses / inference), and ``build_box_predictor`` (:239-249)."""
    mods = _load_reference_modules(["ovr.modeling.logged_module", "ovr.modeling.roi_heads.box_emb_head"])
    return mods["ovr.modeling.roi_heads.box_emb_head"]


def load_reference_grounding_box_head():
    """Returns the reference's box_emb_grounding_head module: ``GroundingModule`` (box_emb_grounding_head.py:60-277, multi-token
    class scoring) and ``EmbeddingGroundingFastRCNNOutputLayers`` (:280-434) on the restated Detectron2 base class."""
    mods = _load_reference_modules(["ovr.modeling.logged_module", "ovr.modeling.roi_heads.box_emb_grounding_head"])
    return mods["ovr.modeling.roi_heads.box_emb_grounding_head"]


def load_reference_roi_heads():
    """Returns the reference's roi_emb_heads module (``EmbeddingRes5ROIHeads`` / ``EmbeddingProposalsRes5ROIHeads``,
    roi_emb_heads.py:121-360) wired to the restated Detectron2 ROIPooler (torchvision roi_align), BottleneckBlock
    res5 stage and the REAL reference box predictor."""
    mods = _load_reference_modules(["ovr.modeling.logged_module", "ovr.modeling.roi_heads.box_emb_head",
                                    "ovr.modeling.roi_heads.roi_emb_heads"])
    return mods["ovr.modeling.roi_heads.roi_emb_heads"]


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


def load_reference_distill():
    """Returns the reference's distill_mmss_gcnn module: ``MultiDistillLoss`` / ``MultiDistillLossJS`` / ``MultiDistillLossL2``
    (distill_mmss_gcnn.py:211-433) are plain torch modules; the meta-architecture class in the same file is never instantiated."""
    mods = _load_reference_modules(["ovr.modeling.logged_module", "ovr.modeling.meta_arch.distill_mmss_gcnn"])
    return mods["ovr.modeling.meta_arch.distill_mmss_gcnn"]


def make_roi_cfg(stage="stt", **over):
    """Attribute-dict with the Detectron2 / OVR config keys the ROI heads and the box predictor read
    (box_emb_head.py:151-177, roi_emb_heads.py:168-241; defaults: ovr/config/config.py + Detectron2 defaults,
    SURVEY.md Appendix C; stage values: configs/coco_stt.yaml, configs/coco_lsm.yaml).  ``over`` = dotted-key overrides,
    e.g. ``{"MODEL.ROI_BOX_HEAD.EMB_DIM": 64}``."""
    lsm = stage == "lsm"
    cfg = _Cfg(
        MODEL=_Cfg(
            MASK_ON=False, KEYPOINT_ON=False,
            ROI_HEADS=_Cfg(NAME="EmbeddingProposalsRes5ROIHeads" if lsm else "EmbeddingRes5ROIHeads", IN_FEATURES=["res4"],
                           NUM_CLASSES=80 if lsm else 48, BATCH_SIZE_PER_IMAGE=200 if lsm else 512, POSITIVE_FRACTION=1.0,
                           PROPOSAL_APPEND_GT=True, SCORE_THRESH_TEST=0.05, NMS_THRESH_TEST=0.5, DETACH_CLASS_PREDICTOR=lsm),
            ROI_BOX_HEAD=_Cfg(NAME="EmbeddingFastRCNNOutputLayers", POOLER_RESOLUTION=14, POOLER_SAMPLING_RATIO=0,
                              POOLER_TYPE="ROIAlignV2", CLS_AGNOSTIC_BBOX_REG=True, EMB_DIM=768, EMBEDDING_BASED=True,
                              FREEZE_EMB_PRED=not lsm, NORMALIZE_EMB_PRED=False, STANDARDIZE_EMB_PRED=False,
                              BBOX_REG_WEIGHTS=(10.0, 10.0, 5.0, 5.0), SMOOTH_L1_BETA=0.0, BBOX_REG_LOSS_TYPE="smooth_l1",
                              BBOX_REG_LOSS_WEIGHT=1.0),
            RESNETS=_Cfg(NUM_GROUPS=1, WIDTH_PER_GROUP=64, RES2_OUT_CHANNELS=256, STRIDE_IN_1X1=True, NORM="FrozenBN",
                         DEFORM_ON_PER_STAGE=[False, False, False, False]),
            # read by EmbeddingGroundingFastRCNNOutputLayers.from_config (box_emb_grounding_head.py:349-361); MAX_TOKENS is the key
            # the reference's config.py never defines — it is set through ``over`` by the cases that need it
            MMSS_HEAD=_Cfg(GROUNDING=_Cfg(LOCAL_METRIC="dot", GLOBAL_METRIC="aligned_local", ALIGNMENT="softmax", ALIGNMENT_TEMPERATURE=10.0)),
        ),
        TEST=_Cfg(DETECTIONS_PER_IMAGE=100),
    )
    for k, v in over.items():
        node = cfg
        parts = k.split(".")
        for p in parts[:-1]:
            node = node[p]
        assert parts[-1] in node or parts[-1] == "MAX_TOKENS", k
        node[parts[-1]] = v
    return cfg
