"""Host-side mirror of the reference's ovr/modeling interfaces for the region-text path."""
from .box_emb_head import Box2BoxTransform, EmbeddingFastRCNNOutputLayers, build_box_predictor  # noqa: F401
from .box_emb_grounding_head import EmbeddingGroundingFastRCNNOutputLayers, GroundingModule  # noqa: F401
from .config import CfgNode, get_cfg  # noqa: F401
from .distill import MultiDistillLoss, MultiDistillLossJS, MultiDistillLossL2, build_distill_loss  # noqa: F401
from .grounding_head import GroundingHead, build_grounding_head  # noqa: F401
from .logged_module import LoggedModule, normalize_vec, standardize_vec  # noqa: F401
from .mmss_heads import build_mmss_heads  # noqa: F401
from .poolers import ROIAlign, ROIPooler  # noqa: F401
from .registry import BOX_PREDICTORS, MMSS_HEADS_REGISTRY, ROI_HEADS_REGISTRY, register_with_detectron2  # noqa: F401
from .roi_emb_heads import EmbeddingProposalsRes5ROIHeads, EmbeddingRes5ROIHeads  # noqa: F401
from .structures import Boxes, Instances, ShapeSpec  # noqa: F401
