# mirrors ops/voxel_pooling/__init__.py:1-3 of the reference
from .voxel_pooling import (voxel_pooling, voxel_pooling_fused, build_plan, PoolingPlan, pool_forward,
                            pool_backward, fused_forward, fused_backward, context_rows_nhwc, voxel_pooling_fused_concat,
                            voxel_pooling_fused_logits, fused_forward_cold)
from .rig import LiftSplatGeometry, rig_variant, voxel_pooling_rig

__all__ = ['voxel_pooling', 'voxel_pooling_fused', 'build_plan', 'PoolingPlan', 'LiftSplatGeometry', 'voxel_pooling_rig',
           'voxel_pooling_fused_concat', 'voxel_pooling_fused_logits']
