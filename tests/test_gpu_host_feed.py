"""locov_b200.HostFeed: the pinned-host -> device staging ring in front of the drop-in heads."""
import pytest
import torch

import locov_b200
import locov_b200.modeling as M
from locov_b200._lib import LocoError
from oracle import lsm_head

pytestmark = pytest.mark.gpu


def _batch(seed):
    ii, ic, w, b = lsm_head.make_lsm_inputs(B=6, Rg=12, T=7, V=64, D=64, seed=seed, gain=4.0)
    pin = lambda d: {k: v.pin_memory() for k, v in d.items()}   # noqa: E731
    return (pin(ii), pin(ic)), w, b


def test_overlapped_feed_gives_the_results_of_synchronous_copies(cuda_device):
    batches = [_batch(s) for s in range(5)]
    cfg = M.get_cfg("lsm")
    head = M.GroundingHead(cfg, 64, 64).to(cuda_device)
    with torch.no_grad():
        head.v2l_projection.weight.copy_(batches[0][1]); head.v2l_projection.bias.copy_(batches[0][2])
    want = []
    with torch.no_grad():
        for (ii, ic), _, _ in batches:
            _, losses, d = head({k: v.to(cuda_device) for k, v in ii.items()}, {k: v.to(cuda_device) for k, v in ic.items()})
            want.append((torch.stack(list(losses.values())).cpu(), d["w2r"].cpu(), d["r2w"].cpu()))
    feed = locov_b200.HostFeed(cuda_device)
    got = []
    feed.submit(batches[0][0])
    with torch.no_grad():
        for i in range(len(batches)):
            if i + 1 < len(batches):
                feed.submit(batches[i + 1][0])
            ii_d, ic_d = feed.take()
            assert all(v.is_cuda for v in ii_d.values())
            _, losses, d = head(ii_d, ic_d)
            got.append((torch.stack(list(losses.values())), d["w2r"].clone(), d["r2w"].clone()))     # no host sync inside the loop
    torch.cuda.synchronize()
    for g, w in zip(got, want):
        for a, b in zip(g, w):
            assert torch.equal(a.cpu(), b)
    assert feed.bytes_copied == sum(v.numel() * v.element_size() for (ii, ic), _, _ in batches for dd in (ii, ic) for v in dd.values())


def test_misuse_is_reported(cuda_device):
    feed = locov_b200.HostFeed(cuda_device)
    with pytest.raises(LocoError):
        feed.take()
    with pytest.raises(LocoError):
        feed.submit({"x": torch.zeros(4)})                      # not pinned
    with pytest.raises(LocoError):
        feed.submit({"x": torch.zeros(4, device=cuda_device)})  # not a host tensor
    feed.submit({"x": torch.zeros(4).pin_memory()})
    feed.submit([torch.ones(3).pin_memory()])                   # another structure in the other slot
    with pytest.raises(LocoError):
        feed.submit({"x": torch.zeros(4).pin_memory()})         # ring full
    assert float(feed.take()["x"].sum()) == 0.0
    assert float(feed.take()[0].sum()) == 3.0
    feed.submit({"x": torch.full((5,), 2.0).pin_memory()})      # slot 0 again, new shape -> re-allocated
    assert float(feed.take()["x"].sum()) == 10.0
    with pytest.raises(LocoError):
        locov_b200.HostFeed("cpu")
