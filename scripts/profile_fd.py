#!/usr/bin/env python
"""A few launches of the articulated-body kernel and its adjoint on a 131 072-row Kuka shard, for ncu
(`ncu --set full -k regex:aba ... python scripts/profile_fd.py`)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import differentiable_robot_model_b200 as drm  # noqa: E402
from differentiable_robot_model_b200.rigid_body_params import UnconstrainedTensor  # noqa: E402
from oracle import drm_oracle as O  # noqa: E402

DEV = "cuda:0"
B = 131072
m = drm.DifferentiableKUKAiiwa(device=DEV)
m.make_link_param_learnable("iiwa_link_1", "inertia_mat", UnconstrainedTensor(dim1=3, dim2=3))
robot = O.load_robot(m.urdf_path, torch.float32)
q, qd, _ = (t.to(DEV) for t in O.sample_inputs(robot, B, seed=0))
f = torch.randn(B, 7, device=DEV)
for _ in range(4):
    qg = q.clone().requires_grad_(True)
    qdd = m.compute_forward_dynamics(qg, qd, f, use_damping=True)
    qdd.square().mean().backward()
torch.cuda.synchronize()
print("ok", float(qdd.abs().max()))
