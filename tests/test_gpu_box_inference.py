"""The inference tail on the B200: loco_box_inference against the restated Detectron2 ``fast_rcnn_inference`` of the
oracle (oracle/d2_stubs.py, torchvision's CPU batched_nms underneath) on IDENTICAL probabilities / deltas / proposals.
Bar: detection counts, classes, kept-row indices and the order bit-exact; scores bit-exact (they are copies); boxes to 1e-5
(``expf`` of the two libms differs in the last place)."""
import numpy as np
import pytest
import torch

from locov_b200 import _lib, ops
from oracle import box_head, d2_stubs

pytestmark = pytest.mark.gpu
WEIGHTS = (10.0, 10.0, 5.0, 5.0)
IMAGE = (320, 480)


def _case(rows, K, seed, gain=3.0, zero_deltas=False, quantize=None):
    g = torch.Generator().manual_seed(seed)
    r = sum(rows)
    logits = torch.randn(r, K + 1, generator=g) * gain
    logits[:, K] = 0.0
    probs = torch.softmax(logits, 1)
    if quantize:                                   # many exactly equal scores: tie order matters
        probs = (probs * quantize).round() / quantize
    deltas = torch.zeros(r, 4) if zero_deltas else torch.randn(r, 4, generator=g) * 0.5
    props = torch.cat([p["proposal_boxes"] for p in box_head.make_proposals(len(rows), max(rows), K, seed=seed + 1, image_size=IMAGE)])
    idx = torch.cat([torch.arange(n) + i * max(rows) for i, n in enumerate(rows)]) if rows else torch.zeros(0, dtype=torch.long)
    return probs, deltas, props[idx]


def _reference(probs, deltas, props, rows, thresh, nms, topk):
    t = d2_stubs.Box2BoxTransform(weights=WEIGHTS)
    boxes = t.apply_deltas(deltas, props)
    out = []
    for b, s in zip(boxes.split(rows), probs.split(rows)):
        out.append(d2_stubs.fast_rcnn_inference_single_image(b, s, IMAGE, thresh, nms, topk))
    return out


def _canonical(scores, classes, kept, boxes):
    """Order within groups of EXACTLY equal scores is unspecified in the reference (torchvision's per-class batched NMS ends in an
    unstable ``sort(descending=True)``); ours is (RoI, class) ascending.  Compare such groups as sets: re-sort by (score desc, class, RoI)."""
    key = sorted(range(len(scores)), key=lambda j: (-float(scores[j]), int(classes[j]), int(kept[j])))
    key = torch.tensor(key, dtype=torch.long)
    return scores[key], classes[key], kept[key], boxes[key]


def _check(probs, deltas, props, rows, thresh, nms, topk, dev, exact_boxes=False, ties=False):
    ref = _reference(probs, deltas, props, rows, thresh, nms, topk)
    n0 = _lib.load().loco_launch_count()
    b, s, c, r, n = ops.box_inference(probs.to(dev), deltas.to(dev), props.to(dev), rows, [IMAGE] * len(rows), WEIGHTS, d2_stubs.Box2BoxTransform(weights=WEIGHTS).scale_clamp,
                                      thresh, nms, topk)
    assert _lib.load().loco_launch_count() - n0 == (4 if sum(rows) else 0)
    n = n.cpu().tolist()
    total = 0
    for i, (inst, kept) in enumerate(ref):
        assert n[i] == len(inst), f"image {i}: {n[i]} detections, reference {len(inst)}"
        k = n[i]
        total += k
        got = (s[i, :k].cpu(), c[i, :k].cpu(), r[i, :k].cpu(), b[i, :k].cpu())
        want = (inst.scores, inst.pred_classes, kept, inst.pred_boxes.tensor)
        if ties:
            assert bool((got[0][:-1] >= got[0][1:]).all())                       # still sorted by score
            got, want = _canonical(*got), _canonical(*want)
        assert torch.equal(got[1], want[1]), f"image {i}: classes / order differ"
        assert torch.equal(got[2], want[2]), f"image {i}: kept rows differ"
        assert torch.equal(got[0], want[0]), f"image {i}: scores differ"
        if exact_boxes:
            assert torch.equal(got[3], want[3])
        elif k:
            assert float((got[3] - want[3]).abs().max()) < 1e-4 * max(IMAGE)
    return total


@pytest.mark.parametrize("rows,K,thresh,nms,topk,gain", [
    ([300, 300, 300], 65, 0.001, 0.3, 50, 2.0),          # dense: every class has hundreds of candidates
    ([300, 300], 1203, 1e-4, 0.5, 300, 4.0),             # LVIS settings
    ([128, 0, 77], 17, 0.05, 0.5, 100, 3.0),             # an image without proposals in the middle
    ([1000], 80, 0.01, 0.5, 100, 3.0),                   # Detectron2's post-NMS proposal count
    ([40, 40], 5, 0.9999, 0.5, 100, 1.0),                # nothing above the threshold
    ([33, 65], 3, 0.0, 0.7, 1000, 1.0),                  # top-k larger than the survivors; K < 8 (partial class group)
])
@pytest.mark.parametrize("zero_deltas", [True, False])
def test_matches_fast_rcnn_inference(cuda_device, rows, K, thresh, nms, topk, gain, zero_deltas):
    probs, deltas, props = _case(rows, K, seed=K + len(rows), gain=gain, zero_deltas=zero_deltas)
    total = _check(probs, deltas, props, rows, thresh, nms, topk, cuda_device, exact_boxes=zero_deltas)
    if thresh < 0.9:
        assert total > 0


def test_equal_scores_resolve_in_candidate_order(cuda_device):
    rows = [200, 150]
    probs, deltas, props = _case(rows, 12, seed=5, gain=1.0, zero_deltas=True, quantize=64.0)
    assert len(torch.unique(probs)) < 70
    _check(probs, deltas, props, rows, 0.1, 0.5, 1000, cuda_device, exact_boxes=True, ties=True)   # (top-k beyond the survivors: no cut inside a tie)


def test_rows_with_non_finite_values_are_dropped_and_renumbered(cuda_device):
    rows = [90, 110]
    probs, deltas, props = _case(rows, 20, seed=9, gain=3.0)
    probs[5, 3] = float("nan")
    probs[17, 20] = float("inf")                      # the background column counts too
    deltas[40, 2] = float("nan")
    deltas[100, 0] = float("inf")
    probs[150, 0] = float("nan")
    _check(probs, deltas, props, rows, 0.01, 0.5, 100, cuda_device)


def test_module_routes_by_range_and_the_torchvision_path_agrees(cuda_device):
    """EmbeddingFastRCNNOutputLayers.inference uses the kernel inside its range (1 <= top-k <= 1024, <= 2048 RoIs per image) and the
    per-image torchvision path outside it (no top-k limit); on the same predictions both give the same detections."""
    import locov_b200.modeling as M
    from locov_b200._lib import LocoError
    cfg = M.get_cfg("stt")
    cfg.MODEL.ROI_BOX_HEAD.EMB_DIM = 32
    bp = M.build_box_predictor(cfg, 64).to(cuda_device).eval()
    g = torch.Generator().manual_seed(2)
    bp.set_class_embeddings(torch.cat([torch.randn(9, 32, generator=g) * 3, torch.zeros(1, 32)]))
    with torch.no_grad():
        bp.emb_pred.weight.mul_(30.0)
    props = box_head.make_proposals(2, 120, 9, seed=4, image_size=IMAGE)
    inst = box_head.instances_from(props, IMAGE, M.Instances, M.Boxes, device=cuda_device)
    x = torch.randn(240, 64, generator=g).to(cuda_device)
    lib = _lib.load()
    with torch.no_grad():
        pred = bp(x)
        bp.test_score_thresh, bp.test_topk_per_image = 0.02, 100
        n0 = lib.loco_launch_count()
        res_k, kept_k = bp.inference(pred, inst)
        n_kernel = lib.loco_launch_count() - n0
        bp.test_topk_per_image = -1                               # Detectron2: no limit -> outside the kernel's range
        n0 = lib.loco_launch_count()
        res_t, kept_t = bp.inference(pred, inst)
        n_torch = lib.loco_launch_count() - n0
    assert n_kernel >= 4 and n_torch == 0
    for a, b, ka, kb in zip(res_k, res_t, kept_k, kept_t):
        m = len(a)
        assert 0 < m <= 100 and len(b) >= m
        assert torch.equal(a.pred_classes, b.pred_classes[:m]) and torch.equal(ka, kb[:m])
        assert torch.equal(a.scores, b.scores[:m])
        assert float((a.pred_boxes.tensor - b.pred_boxes.tensor[:m]).abs().max()) < 1e-3
    with pytest.raises(LocoError):                                # the raw entry refuses what it cannot do
        ops.box_inference(torch.rand(2100, 4, device=cuda_device), torch.zeros(2100, 4, device=cuda_device), torch.rand(2100, 4, device=cuda_device) * 50,
                          [2100], [IMAGE], WEIGHTS, 4.0, 0.1, 0.5, 100)
