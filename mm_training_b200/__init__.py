"""B200-native BEV projection hot path of aimotive/mm_training (voxel pooling,
hard voxelizer, pillar scatter) behind the reference's operator API."""
__version__ = '0.1.0'
