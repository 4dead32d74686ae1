"""Which accumulation order of the rig plan kernel reproduces torch's geometry on this device?
Prints, per variant, the number of points whose (x, y, z) cell index differs from mm_training_b200.geometry
computed by torch on the GPU and on the CPU (run on the GPU box)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mm_training_b200 import _lib, synthetic
from mm_training_b200.configs import CFG_2, CFG_AIM
from mm_training_b200.ops.voxel_pooling.rig import LiftSplatGeometry, _random_rigs

dev = 'cuda'
gen = torch.Generator().manual_seed(7)
for name, cfg, rigs in (('cfg2 random rigs', CFG_2, _random_rigs(2, 4, gen)),
                        ('aim random rigs', CFG_AIM, _random_rigs(1, 2, gen))):
    lsg = LiftSplatGeometry.from_config(cfg, dev)
    s2e, k = rigs
    ref_gpu = lsg.geom_xyz(s2e.to(dev), k.to(dev))
    lsg_cpu = LiftSplatGeometry.from_config(cfg, 'cpu')
    ref_cpu = lsg_cpu.geom_xyz(s2e, k).to(dev)
    cmb = lsg.combine(s2e.to(dev), k.to(dev))
    cmb_cpu = lsg_cpu.combine(s2e, k).to(dev)
    print(name, 'points', ref_gpu.numel() // 3, 'gpu-vs-cpu torch mismatches', int((ref_gpu != ref_cpu).any(-1).sum()),
          'combine equal', bool(torch.equal(cmb, cmb_cpu)))
    for v in range(_lib.lib().bevpool_rig_num_variants()):
        g = lsg.rig_geom(cmb, v)
        gc = lsg.rig_geom(cmb_cpu, v)
        print(f'  variant {v}: vs torch-gpu {int((g != ref_gpu).any(-1).sum())}   vs torch-cpu (cpu combine) {int((gc != ref_cpu).any(-1).sum())}')
