#!/usr/bin/env python
"""Launch each kernel of interest a few times at a representative size, for `ncu -k regex:<name>` captures
(scripts/gpu_profile.sh).  WHICH selects the group: fk_tree | fk_allegro | rnea | rnea_bwd | fk_bwd | aba."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import differentiable_robot_model_b200 as drm  # noqa: E402
from differentiable_robot_model_b200 import engine  # noqa: E402

DEV = torch.device("cuda", 0)
WHICH = os.environ.get("WHICH", "fk_tree")
REPS = int(os.environ.get("REPS", "4"))


def sample(m, batch, seed=0):
    gen = torch.Generator().manual_seed(seed)
    lim = m.get_joint_limits()
    lo = torch.tensor([float(l["lower"]) for l in lim]); hi = torch.tensor([float(l["upper"]) for l in lim])
    vel = torch.tensor([float(l["velocity"]) for l in lim])
    u = torch.rand(3, batch, len(lim), generator=gen)
    return ((lo + (hi - lo) * u[0]).to(DEV), ((2 * u[1] - 1) * 0.2 * vel).to(DEV), ((2 * u[2] - 1) * 0.4 * vel).to(DEV))


def allegro():
    return drm.DifferentiableRobotModel(os.path.join(drm.robot_model.robot_description_folder,
                                                     "allegro/urdf/allegro_hand_description_left.urdf"), device=DEV)


def main():
    if WHICH in ("fk_tree", "fk_allegro"):
        m = allegro()
        tips = ["link_3.0_tip", "link_7.0_tip", "link_11.0_tip", "link_15.0_tip"]
        ees = [m._name_to_idx_map[t] for t in tips]
        table, topo = m._link_table(), m._topology
        if WHICH == "fk_tree":
            q = sample(m, 1 << 20)[0]
            for _ in range(REPS):
                engine.fk_jacobian_multi_raw(topo, ees, table, q)
        else:
            q = sample(m, 1 << 21)[0]
            for _ in range(REPS):
                engine.fk_jacobian_raw(topo, ees[3], table, q)
                engine.fk_jacobian_multi_raw(topo, ees[3:], table, q)
    elif WHICH == "rnea":
        for cls, b in ((drm.DifferentiableFrankaPanda, 1 << 21), (drm.DifferentiableFrankaPanda, 65536)):
            m = cls(device=DEV)
            q, qd, qdd = sample(m, b)
            for _ in range(REPS):
                engine.inverse_dynamics_raw(m._topology, m._link_table(), q, qd, qdd, 3)
    elif WHICH in ("rnea_bwd", "fk_bwd"):
        m = drm.DifferentiableKUKAiiwa(device=DEV)
        b = 131072
        q, qd, qdd = sample(m, b)
        table, topo, ee = m._link_table(), m._topology, m._name_to_idx_map["iiwa_link_ee"]
        lib, P, S = engine.lib(), engine._ptr, engine._stream
        ws = engine._workspace(topo, b, DEV)
        qg, tg = torch.empty_like(q), torch.zeros_like(table)
        g_tau = torch.randn(b, 7, device=DEV)
        g = [torch.randn(b, 3, device=DEV), torch.randn(b, 4, device=DEV), torch.randn(b, 3, 7, device=DEV), torch.randn(b, 3, 7, device=DEV)]
        for _ in range(REPS):
            if WHICH == "rnea_bwd":
                assert lib.drmb200_inverse_dynamics_backward(ctypes.byref(topo), P(table), P(q), P(qd), P(qdd), b, 3, P(g_tau),
                                                             P(qg), P(qg), P(qg), P(tg), P(ws), S()) == 0
                assert lib.drmb200_inverse_dynamics_backward(ctypes.byref(topo), P(table), P(q), P(qd), P(qdd), b, 3 | 4, P(g_tau),
                                                             None, None, None, P(tg), P(ws), S()) == 0
            else:
                assert lib.drmb200_fk_jacobian_backward(ctypes.byref(topo), ee, P(table), P(q), b, P(g[0]), P(g[1]), P(g[2]), P(g[3]),
                                                        P(qg), P(tg), P(ws), S()) == 0
    elif WHICH == "aba":
        m = drm.DifferentiableKUKAiiwa(device=DEV)
        b = 131072
        q, qd, _ = sample(m, b)
        f = torch.randn(b, 7, device=DEV)
        q.requires_grad_(True)
        for _ in range(REPS):
            acc = m.compute_forward_dynamics(q, qd, f)
            acc.sum().backward()
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
