"""The Detectron2 config keys the hot path reads, with the reference's defaults.

Reference: ovr/config/config.py:4-174 (add_ovr_config) + the Detectron2 defaults listed in
SURVEY.md Appendix C + the two shipped YAMLs (configs/coco_stt.yaml, configs/coco_lsm.yaml).
``get_cfg()`` returns a yacs-like attribute node so the drop-in modules can be built without
Detectron2; a real Detectron2 ``CfgNode`` works the same way (attribute access only).
``MODEL.B200`` holds this library's optional keys (absent = the reference's behaviour):
  PRECISION              "fp32" (three-pass fp32-accurate tensor-core mode, the reference's numerics) or "bf16";
  POOLER_CHANNELS_LAST   RoIAlign writes the pooled tensor in torch.channels_last memory format (same logical shape/values);
  POOLER_BF16            ... and in bf16 (reduced precision: for a res5 stage run in bf16 by the caller).
"""
import copy


class CfgNode(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    def clone(self):
        return copy.deepcopy(self)

    def merge_from_list(self, opts):
        assert len(opts) % 2 == 0
        for k, v in zip(opts[0::2], opts[1::2]):
            node = self
            parts = k.split(".")
            for p in parts[:-1]:
                node = node[p]
            assert parts[-1] in node, f"Non-existent config key: {k}"
            node[parts[-1]] = v
        return self


def _n(**kw):
    return CfgNode(**kw)


def get_cfg(stage="stt"):
    """stage: "stt" (configs/coco_stt.yaml) or "lsm" (configs/coco_lsm.yaml)."""
    lsm = stage == "lsm"
    cfg = _n(
        MODEL=_n(
            MASK_ON=False, KEYPOINT_ON=False,
            LOAD_EMB_PRED_FROM_MMSS_HEAD=True,
            B200=_n(PRECISION="fp32", POOLER_CHANNELS_LAST=False, POOLER_BF16=False),
            ROI_HEADS=_n(NAME="EmbeddingProposalsRes5ROIHeads" if lsm else "EmbeddingRes5ROIHeads",
                         IN_FEATURES=["res4"], NUM_CLASSES=80 if lsm else 48,
                         BATCH_SIZE_PER_IMAGE=200 if lsm else 512, POSITIVE_FRACTION=1.0, IOU_THRESHOLDS=[0.5],
                         IOU_LABELS=[0, 1], PROPOSAL_APPEND_GT=True, SCORE_THRESH_TEST=0.05, NMS_THRESH_TEST=0.5,
                         DETACH_CLASS_PREDICTOR=lsm),
            ROI_BOX_HEAD=_n(NAME="EmbeddingFastRCNNOutputLayers", POOLER_RESOLUTION=14, POOLER_SAMPLING_RATIO=0,
                            POOLER_TYPE="ROIAlignV2", CLS_AGNOSTIC_BBOX_REG=True, EMB_DIM=768, EMBEDDING_BASED=True,
                            FREEZE_EMB_PRED=not lsm, NORMALIZE_EMB_PRED=False, STANDARDIZE_EMB_PRED=False,
                            BBOX_REG_WEIGHTS=(10.0, 10.0, 5.0, 5.0), SMOOTH_L1_BETA=0.0,
                            BBOX_REG_LOSS_TYPE="smooth_l1", BBOX_REG_LOSS_WEIGHT=1.0),
            RESNETS=_n(NUM_GROUPS=1, WIDTH_PER_GROUP=64, RES2_OUT_CHANNELS=256, STRIDE_IN_1X1=True, NORM="FrozenBN",
                       DEFORM_ON_PER_STAGE=[False, False, False, False]),
            MMSS_HEAD=_n(TYPES=("GroundingHead",), DEFAULT_HEAD="GroundingHead", TIE_VL_PROJECTION_WEIGHTS=lsm,
                         IN_FEATURES="res5", SPATIAL_DROPOUT=100 if lsm else -1, DISTILLATION_LOSS=lsm,
                         DISTILLATION_LOSS_TYPE="KD", DISTILLATION_TEMPERATURE=10.0 if lsm else 1.0, DISTILLATION_LOSS_WEIGHT=1.0,
                         DISTILLATION_DETACH_TEACHER=False, DISTILLATION_TEACHER_TRANSFORMER=not lsm,
                         GROUNDING=_n(LOCAL_METRIC="dot", GLOBAL_METRIC="aligned_local", ALIGNMENT="softmax",
                                      ALIGNMENT_TEMPERATURE=10.0, LOSS="cross_entropy", NEGATIVE_MINING="random",
                                      TRIPLET_MARGIN=1.0, ALIGN_WORDS_TO_REGIONS=True, ALIGN_REGIONS_TO_WORDS=True,
                                      TEXT_INPUT="input_embeddings")),
            LANGUAGE_BACKBONE=_n(FREEZE=True),
        ),
        TEST=_n(DETECTIONS_PER_IMAGE=100),
        SEED=1992,
    )
    return cfg
