"""The benchmark's synthetic workloads (locov_b200/synthetic.py) are value-for-value the oracle's seeded inputs."""
import torch

from locov_b200 import synthetic
from oracle import box_head, lsm_head


def test_lsm_inputs_equal_the_oracle_generator():
    for kw in (dict(B=4, Rg=10, T=7, V=32, D=16, seed=3), dict(B=5, Rg=12, T=9, V=32, D=16, seed=4, ragged_regions=True, gain=3.0)):
        a, b = synthetic.lsm_inputs(**kw), lsm_head.make_lsm_inputs(**kw)
        for da, db in zip(a[:2], b[:2]):
            assert da.keys() == db.keys() and all(torch.equal(da[k], db[k]) for k in da)
        assert torch.equal(a[2], b[2]) and torch.equal(a[3], b[3])


def test_box_inputs_equal_the_oracle_generator():
    a, b = synthetic.box_inputs(33, 7, V=32, D=16, seed=5), box_head.make_box_inputs(33, 7, V=32, D=16, seed=5)
    assert all(torch.equal(x, y) for x, y in zip(a, b))


def test_coco_boxes_are_inside_the_image():
    r = synthetic.coco_boxes(3, 50, seed=1)
    assert r.shape == (150, 5) and set(r[:, 0].tolist()) == {0.0, 1.0, 2.0}
    assert bool((r[:, 1] <= r[:, 3]).all()) and bool((r[:, 2] <= r[:, 4]).all())
    assert float(r[:, 3].max()) <= 1216 and float(r[:, 4].max()) <= 800 and float(r[:, 1:].min()) >= 0
