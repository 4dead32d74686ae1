"""CPU oracle for the BEV-projection hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import it, and only as the checker or as the timed CPU baseline.  The
product path (``mm_training_b200``) never imports this package and fails loudly
when its CUDA library is missing.

Pinning status
--------------
* voxel pooling (forward/backward/fused): pinned by the reference's own unit
  test recipe (``test/test_ops/test_voxel_pooling.py:15-37`` of the reference)
  -- ``tests/test_oracle_voxel_pool.py`` re-runs that python-loop golden -- and,
  on the GPU box, by the reference's own CUDA kernel compiled from its sources
  into ``oracle/_ref`` (``oracle/build_ref.sh``).
* geometry / index quantisation: restated from ``layers/backbones/lss_fpn.py``
  with the same torch ops; no reference test pins values (shape-only, stale).
* voxelizer / VFE / pillar scatter: **parity unpinned** -- the arithmetic lives in
  mmcv-full==1.7.0 / mmdet3d==1.0.0rc4 / spconv, none of which is vendored in
  the reference or installed here.  The restatement follows the published
  mmcv CPU algorithm and the hand-derived known-answer
  test of Appendix A.4.
"""
