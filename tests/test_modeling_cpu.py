"""Host-side mirror of the reference interfaces: names, constructor/config keys, parameter layouts,
registries, and that the product path refuses to run without the CUDA library (no CPU fallback)."""
import pytest
import torch

import locov_b200.modeling as M
from locov_b200 import LocoError
from oracle import lsm_head


def test_registries_expose_the_reference_names():
    assert "GroundingHead" in M.MMSS_HEADS_REGISTRY
    assert "EmbeddingRes5ROIHeads" in M.ROI_HEADS_REGISTRY and "EmbeddingProposalsRes5ROIHeads" in M.ROI_HEADS_REGISTRY
    assert "EmbeddingFastRCNNOutputLayers" in M.BOX_PREDICTORS
    with pytest.raises(KeyError):
        M.MMSS_HEADS_REGISTRY.get("NoSuchHead")


def test_box_predictor_state_dict_and_class_embeddings():
    cfg = M.get_cfg("stt")
    bp = M.build_box_predictor(cfg, 2048)
    sd = bp.state_dict()
    assert sd["emb_pred.weight"].shape == (768, 2048) and sd["emb_pred.bias"].shape == (768,)
    assert sd["bbox_pred.weight"].shape == (4, 2048) and sd["bbox_pred.bias"].shape == (4,)
    assert not bp.emb_pred.weight.requires_grad          # FREEZE_EMB_PRED (coco_stt.yaml:36)
    assert bp.cls_score is None and bp.num_classes is None
    embs = torch.cat([torch.randn(48, 768), torch.zeros(1, 768)])
    bp.set_class_embeddings(embs.numpy())
    assert bp.num_classes == 48 and bp.cls_score.weight.shape == (49, 768)
    assert not bp.cls_score.weight.requires_grad and float(bp.cls_score.bias.abs().sum()) == 0.0
    assert "cls_score.weight" in bp.state_dict()
    bp.set_class_embeddings(torch.cat([torch.randn(65, 768), torch.zeros(1, 768)]))     # re-settable (trainer.py:191)
    assert bp.num_classes == 65
    lsm = M.build_box_predictor(M.get_cfg("lsm"), 2048)
    assert lsm.detach_cls_predictor and lsm.loss_weight["loss_cls"] == 0.0 and lsm.emb_pred.weight.requires_grad


def test_grounding_head_options_and_tying():
    cfg = M.get_cfg("lsm")
    heads = M.build_mmss_heads(cfg, 2048, 768)
    gh = heads["GroundingHead"]
    assert isinstance(gh.v2l_projection.weight, torch.nn.Parameter) and gh.v2l_projection.weight.shape == (768, 2048)
    assert gh.temperature == 10.0 and gh.return_dist and isinstance(gh.log_info, dict)
    bad = M.get_cfg("lsm")
    bad.MODEL.MMSS_HEAD.GROUNDING.ALIGNMENT = "random_top3"
    with pytest.raises(NotImplementedError):
        M.GroundingHead(bad, 2048, 768)
    bad = M.get_cfg("lsm")
    bad.MODEL.MMSS_HEAD.GROUNDING.GLOBAL_METRIC = "reconstruction_mse"
    with pytest.raises(NotImplementedError):
        M.GroundingHead(bad, 2048, 768)


def test_no_cpu_fallback():
    """CPU tensors are rejected loudly — the product path is the CUDA library or nothing."""
    cfg = M.get_cfg("lsm")
    gh = M.GroundingHead(cfg, 32, 16)
    ii, ic, w, b = lsm_head.make_lsm_inputs(B=2, Rg=4, T=5, V=32, D=16)
    with pytest.raises(LocoError):
        gh(ii, ic)
    pool = M.ROIPooler(7, (1 / 16,), 0, "ROIAlignV2")
    with pytest.raises(LocoError):
        pool([torch.randn(1, 4, 8, 8)], [M.Boxes(torch.tensor([[0., 0., 32., 32.]]))])
    with pytest.raises(NotImplementedError):
        M.ROIPooler(7, (1 / 16,), 0, "ROIPool")


def test_logged_module_is_lazy():
    m = M.LoggedModule()
    t = torch.arange(6.0).reshape(2, 3)
    m.log("x", t)
    s = m.log_info["x"]
    assert s["min"] == 0.0 and s["max"] == 5.0 and abs(s["mean"] - 2.5) < 1e-6 and tuple(s["shape"]) == (2, 3)
    m.log_dict({"loss": torch.tensor(1.0)})
    assert float(m.log_info["loss"]) == 1.0


def test_pooler_box_format():
    from locov_b200.modeling.poolers import convert_boxes_to_pooler_format
    r = convert_boxes_to_pooler_format([M.Boxes(torch.ones(2, 4)), M.Boxes(torch.zeros(0, 4)), M.Boxes(2 * torch.ones(1, 4))])
    assert r.tolist() == [[0, 1, 1, 1, 1], [0, 1, 1, 1, 1], [2, 2, 2, 2, 2]]


# ---- the (cfg, input_shape) construction path Detectron2's build_roi_heads / the reference's build_box_predictor use ----
from oracle import box_cases  # noqa: E402
from util import load_golden  # noqa: E402


def _small_cfg(stage, cls_name):
    cfg = M.get_cfg(stage)
    cfg.MODEL.RESNETS.RES2_OUT_CHANNELS = 8
    cfg.MODEL.RESNETS.WIDTH_PER_GROUP = 2
    cfg.MODEL.ROI_BOX_HEAD.EMB_DIM = 32
    cfg.MODEL.ROI_HEADS.NAME = cls_name
    cfg.MODEL.ROI_HEADS.NUM_CLASSES = box_cases.ROI_SHAPE["K"]
    return cfg


def test_heads_are_constructible_the_way_the_registries_call_them():
    """``ROI_HEADS_REGISTRY.get(name)(cfg, input_shape)`` (Detectron2 build_roi_heads) and
    ``BOX_EMBEDDING_PREDICTORS[name](cfg, input_shape)`` (box_emb_head.py:249)."""
    for stage, name in (("stt", "EmbeddingRes5ROIHeads"), ("lsm", "EmbeddingProposalsRes5ROIHeads")):
        cfg = _small_cfg(stage, name)
        heads = M.ROI_HEADS_REGISTRY.get(cfg.MODEL.ROI_HEADS.NAME)(cfg, {"res4": M.ShapeSpec(channels=32, stride=16)})
        assert type(heads).__name__ == name and heads.in_features == ["res4"] and heads.output_shape == 64
        assert heads.num_classes == 6 and heads.batch_size_per_image == cfg.MODEL.ROI_HEADS.BATCH_SIZE_PER_IMAGE
        assert heads.pooler.output_size == (14, 14) and heads.pooler.level_poolers[0].spatial_scale == 1 / 16
        assert len(heads.res5) == 3 and heads.res5[0].stride == 2 and heads.res5[0].shortcut is not None and heads.res5[1].shortcut is None
        assert type(heads.box_predictor).__name__ == "EmbeddingFastRCNNOutputLayers" and heads.box_predictor.emb_dim == 32
    bp = M.BOX_PREDICTORS["EmbeddingFastRCNNOutputLayers"](M.get_cfg("stt"), M.ShapeSpec(channels=2048))
    assert bp.emb_pred.weight.shape == (768, 2048) and bp.loss_weight == {"loss_box_reg": 1.0} and bp.precision == "fp32"
    bp = M.EmbeddingFastRCNNOutputLayers(M.get_cfg("stt"), 2048, precision="bf16")       # keyword override next to a cfg
    assert bp.precision == "bf16"
    bp = M.EmbeddingFastRCNNOutputLayers(64, cls_agnostic_bbox_reg=True, emb_dim=16)    # explicit-argument form
    assert bp.emb_pred.weight.shape == (16, 64)
    # explicit-argument form of the ROI heads with an injected res5
    heads = M.EmbeddingRes5ROIHeads(in_features=["res4"], pooler=M.ROIPooler(7, (1 / 16,), 0, "ROIAlignV2"), res5=torch.nn.Identity(),
                                    box_predictor=bp, num_classes=3)
    assert heads.num_classes == 3 and isinstance(heads.res5, torch.nn.Identity)


@pytest.mark.parametrize("case", sorted(box_cases.ROI_CASES))
def test_reference_state_dicts_load_strictly(case):
    """The state_dict the REAL reference ROI heads produced (tests/golden/roiheads_*.npz) loads with identical keys."""
    z = load_golden("roiheads_" + case)
    cls_name, stage, _ = box_cases.ROI_CASES[case]
    heads = M.ROI_HEADS_REGISTRY.get(cls_name)(_small_cfg(stage, cls_name), {"res4": M.ShapeSpec(channels=32, stride=16)})
    heads.box_predictor.set_class_embeddings(torch.from_numpy(z["cls"]))
    sd = {k[len("state::"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("state::")}
    assert sorted(sd) == sorted(heads.state_dict().keys())
    heads.load_state_dict(sd, strict=True)


def test_box_predictor_state_keys_match_the_reference_class():
    z = load_golden("box_k65_infer")
    bp = M.build_box_predictor(M.get_cfg("stt"), 2048)
    bp.set_class_embeddings(torch.zeros(66, 768))
    assert sorted(bp.state_dict().keys()) == sorted(z["state_keys"].tolist())


def test_host_feed_refuses_a_cpu_device():
    import locov_b200
    from locov_b200._lib import LocoError
    with pytest.raises(LocoError):
        locov_b200.HostFeed("cpu")
    with pytest.raises(LocoError):
        locov_b200.HostFeed("cuda:0", depth=1)


def test_multi_token_box_predictor_is_built_from_a_cfg_like_the_reference():
    """box_emb_head.py:239-249 resolves "EmbeddingGroundingFastRCNNOutputLayers" too; its from_config (box_emb_grounding_head.py:349-381)
    builds the GroundingModule from the MMSS_HEAD.GROUNDING keys.  State-dict keys as in the reference (tests/golden/gbox_*.npz)."""
    import numpy as np
    cfg = M.get_cfg("stt")
    cfg.MODEL.ROI_BOX_HEAD.NAME = "EmbeddingGroundingFastRCNNOutputLayers"
    cfg.MODEL.ROI_BOX_HEAD.EMB_DIM = 16
    cfg.MODEL.MMSS_HEAD.GROUNDING.ALIGNMENT = "hardmax"
    bp = M.build_box_predictor(cfg, M.ShapeSpec(channels=24))
    assert isinstance(bp, M.EmbeddingGroundingFastRCNNOutputLayers) and isinstance(bp.cls_score, M.GroundingModule)
    assert bp.cls_score.alignment == "hardmax" and bp.cls_score.temperature == 10.0 and bp.num_classes is None
    assert bp.emb_pred.weight.requires_grad                      # (the reference never applies FREEZE_EMB_PRED in this class)
    bp.set_class_embeddings({0: torch.randn(2, 16), 1: torch.randn(1, 16), 2: np.random.randn(3, 16).astype("float32")})
    assert bp.num_classes == 3 and bp.cls_score.token_score.weight.shape == (7, 16)
    assert bp.cls_score.seg_off.tolist() == [0, 2, 3, 6, 7] and bp.cls_score.num_tok.tolist() == [2, 1, 3, 1]
    assert float(bp.cls_score.token_score.weight[-1].abs().sum()) == 0.0          # the background token
    z = np.load("tests/golden/gbox_soft_t10.npz")
    assert sorted(bp.state_dict().keys()) == list(z["state_keys"])
    with pytest.raises(NotImplementedError):
        M.GroundingModule(16, 3, 4, local_metric="cosine", normalize_emb=True)
