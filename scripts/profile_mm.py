#!/usr/bin/env python
"""A few launches of the mass-matrix and inverse-dynamics kernels on 131 072 Kuka configurations, for ncu
(`ncu --set full -k regex:"mass_matrix|rnea_kernel" ... python scripts/profile_mm.py`)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import differentiable_robot_model_b200 as drm  # noqa: E402
from oracle import drm_oracle as O  # noqa: E402

DEV = "cuda:0"
B = 131072
m = drm.DifferentiableKUKAiiwa(device=DEV)
robot = O.load_robot(m.urdf_path, torch.float32)
q, qd, qdd = (t.to(DEV) for t in O.sample_inputs(robot, B, seed=0))
for _ in range(4):
    H = m.compute_lagrangian_inertia_matrix(q)
    tau = m.compute_inverse_dynamics(q, qd, qdd)
torch.cuda.synchronize()
print("ok", float(H.abs().max()), float(tau.abs().max()))
