#!/usr/bin/env python
"""Per-body dynamic state of the REFERENCE after compute_inverse_dynamics: `_bodies[i].vel / .acc / .force`
(robot_model.py:183-193, 262-301), for three robots (chain, tree, arm + hand), 9 rows each -> tests/golden/state_*.npz.

Build-container only (needs /root/reference):    python tests/golden/make_golden_state.py"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(REPO, "oracle", "refshim"))
sys.path.insert(0, "/root/reference")

from differentiable_robot_model.robot_model import DifferentiableRobotModel  # noqa: E402

DATA = "/root/reference/diff_robot_data"
ROBOTS = {"state_iiwa7": "kuka_iiwa/urdf/iiwa7.urdf",
          "state_allegro_left": "allegro/urdf/allegro_hand_description_left.urdf",
          "state_iiwa7_allegro": "kuka_iiwa/urdf/iiwa7_allegro.urdf",
          "state_trifinger_edu": "trifinger_edu_description/trifinger_edu.urdf"}


def main():
    for stem, rel in ROBOTS.items():
        with contextlib.redirect_stdout(io.StringIO()):
            model = DifferentiableRobotModel(os.path.join(DATA, rel), stem)
        rng = np.random.RandomState(2)
        limits = model.get_joint_limits()
        lo = np.array([l["lower"] for l in limits]); hi = np.array([l["upper"] for l in limits])
        vel = np.array([l["velocity"] for l in limits])
        n = model._n_dofs
        f32 = lambda a: torch.tensor(a, dtype=torch.float32)  # noqa: E731
        q, qd, qdd = f32(rng.uniform(lo, hi, (9, n))), f32(rng.uniform(-0.2 * vel, 0.2 * vel, (9, n))), f32(rng.uniform(-0.4 * vel, 0.4 * vel, (9, n)))
        out = {"q": q.numpy(), "qd": qd.numpy(), "qdd": qdd.numpy()}
        for grav in (True, False):
            with torch.no_grad():
                tau = model.compute_inverse_dynamics(q, qd, qdd, include_gravity=grav, use_damping=True)
            tag = "g1" if grav else "g0"
            out[f"tau.{tag}"] = tau.numpy().copy()
            B = q.shape[0]
            exp = lambda t: t.expand(B, 3).numpy().copy()  # noqa: E731  (root state may be broadcastable)
            out[f"vel_ang.{tag}"] = np.stack([exp(b.vel.ang) for b in model._bodies])
            out[f"vel_lin.{tag}"] = np.stack([exp(b.vel.lin) for b in model._bodies])
            out[f"acc_ang.{tag}"] = np.stack([exp(b.acc.ang) for b in model._bodies])
            out[f"acc_lin.{tag}"] = np.stack([exp(b.acc.lin) for b in model._bodies])
            out[f"force_ang.{tag}"] = np.stack([exp(b.force.ang) for b in model._bodies])
            out[f"force_lin.{tag}"] = np.stack([exp(b.force.lin) for b in model._bodies])
        np.savez_compressed(os.path.join(HERE, stem + ".npz"), **out)
        print(stem, {k: v.shape for k, v in out.items() if k.startswith("force_lin")})


if __name__ == "__main__":
    main()
