"""ORACLE — TEST INFRASTRUCTURE ONLY.

CPU restatements of the reference's region-text hot path (lmb-freiburg/locov).  Nothing under
``locov_b200/`` may import this package: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` use it, and only as the checker or the
CPU baseline — never as the thing shipped.

Pinning status:
  * roi_align  — pinned bit-for-bit against the compiled CPU op ``torch.ops.torchvision.roi_align``
                 (the third-party kernel the reference reaches through Detectron2).
  * lsm_head   — pinned against the REAL reference ``GroundingHead`` class imported from
                 /root/reference in the build container (oracle/ref_loader.py); outputs committed as
                 tests/golden/lsm_*.npz by tests/golden/make_golden.py.
  * box_head   — the reference class subclasses Detectron2's FastRCNNOutputLayers (absent here and
                 un-vendored), so it cannot be imported: restated from box_emb_head.py:179-236 and
                 Detectron2's published semantics, checked against torch's own F.linear /
                 F.cross_entropy / F.softmax.  The reference has no golden vectors for it
                 (SURVEY.md §4): "parity unpinned" beyond those torch primitives.
"""
