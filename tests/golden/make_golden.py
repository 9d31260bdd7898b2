#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by RUNNING THE REFERENCE ITSELF.

Build-container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference (`/root/reference/differentiable_robot_model`, imported unmodified; the only shim is
`oracle/refshim/urdf_parser_py`, an XML reader standing in for the uninstalled third-party parser)
is evaluated in fp32 on CPU on seeded inputs for every well-formed shipped URDF.  One `.npz` per
robot stores inputs, the link parameters as the reference parsed them, all-link poses, FK pose +
quaternion and Jacobians of several links, inverse-dynamics torques for the four
(include_gravity, use_damping) combinations, and autograd gradients of fixed random linear losses
w.r.t. q / qd / qdd and every learnable link parameter.  The reference stores no expected values
of its own (its tests compare with live pybullet), so these files are what pins `oracle/` and the
CUDA engine to the reference.

Row 0 of every batch is the hand-checkable input q_k = 0.1k, qd_k = 0.05k, qdd_k = -0.02k of
SURVEY.md section 8(c); the other rows are drawn like tests/test_kinematics_dynamics.py:162-189
(np.random.uniform within the joint limits) with the velocity / acceleration ranges of BASELINE.md.
"""
import os
import sys
import io
import contextlib

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(REPO, "oracle", "refshim"))
sys.path.insert(0, "/root/reference")

from differentiable_robot_model.robot_model import DifferentiableRobotModel  # noqa: E402
from differentiable_robot_model.rigid_body_params import UnconstrainedScalar, UnconstrainedTensor  # noqa: E402

DATA = "/root/reference/diff_robot_data"
BATCH = 9

# (relative URDF path, links whose FK / Jacobian are stored)
ROBOTS = [
    ("2link_robot.urdf", ["endEffector", "arm2"]),
    ("kuka_iiwa/urdf/iiwa7.urdf", ["iiwa_link_ee", "iiwa_link_4", "iiwa_link_7"]),
    ("panda_description/urdf/panda_no_gripper.urdf", ["panda_virtual_ee_link", "panda_link5"]),
    ("panda_description/urdf/panda.urdf", ["panda_leftfinger", "panda_rightfinger", "panda_virtual_ee_link"]),
    ("allegro/urdf/allegro_hand_description_left.urdf",
     ["link_11.0_tip", "link_7.0_tip", "link_3.0_tip", "link_15.0_tip", "link_2.0"]),
    ("allegro/urdf/allegro_hand_description_left_small_damping.urdf", ["link_15.0_tip"]),
    ("trifinger_edu_description/trifinger_edu.urdf",
     ["finger_tip_link_0", "finger_tip_link_120", "finger_tip_link_240"]),
    ("kinova_description/urdf/jaco_clean.urdf", ["j2n6s300_link_ee", "j2n6s300_link_finger_tip_1"]),
    ("kinova_description/urdf/jaco.urdf", ["j2n6s300_end_effector", "j2n6s300_link_finger_tip_3"]),
    ("fetch_description/urdf/fetch_arm_no_gripper.urdf", ["virtual_ee_link"]),
    ("fetch_description/urdf/fetch_arm_no_gripper_small_damping.urdf", ["virtual_ee_link", "elbow_flex_link"]),
    ("kuka_iiwa/urdf/iiwa7_allegro.urdf", ["link_15.0_tip", "link_3.0_tip", "palm_link"]),
]


def quiet_model(path):
    with contextlib.redirect_stdout(io.StringIO()):
        return DifferentiableRobotModel(path, "golden")


def sample(model, rng):
    limits = model.get_joint_limits()
    n = model._n_dofs
    lo = np.array([l["lower"] for l in limits])
    hi = np.array([l["upper"] for l in limits])
    vel = np.array([l["velocity"] for l in limits])
    q = rng.uniform(lo, hi, size=(BATCH, n))
    qd = rng.uniform(-0.2 * vel, 0.2 * vel, size=(BATCH, n))
    qdd = rng.uniform(-0.4 * vel, 0.4 * vel, size=(BATCH, n))
    k = np.arange(1, n + 1)
    q[0], qd[0], qdd[0] = 0.1 * k, 0.05 * k, -0.02 * k
    f = lambda a: torch.tensor(a, dtype=torch.float32)  # noqa: E731
    return f(q), f(qd), f(qdd)


def make_all_learnable(model):
    """Swap every link parameter for an unconstrained module initialised at the URDF value."""
    learnable = {}
    for i, body in enumerate(model._bodies):
        if i == 0:
            continue
        name = body.name
        inits = {
            "mass": UnconstrainedScalar(init_val=body.inertia.mass().detach().clone()),
            "com": UnconstrainedTensor(1, 3, init_tensor=body.inertia.com().detach().clone().reshape(1, 3)),
            "inertia_mat": UnconstrainedTensor(3, 3, init_tensor=body.inertia.inertia_mat().detach().clone().reshape(3, 3)),
        }
        if body.joint_idx is not None:          # fixed-joint origins are frozen in the reference (quirk 4)
            inits["trans"] = UnconstrainedTensor(1, 3, init_tensor=body.trans().detach().clone().reshape(1, 3))
            inits["rot_angles"] = UnconstrainedTensor(1, 3, init_tensor=body.rot_angles().detach().clone().reshape(1, 3))
            inits["joint_damping"] = UnconstrainedScalar(init_val=body.joint_damping().detach().clone())
        for pname, module in inits.items():
            model.make_link_param_learnable(name, pname, module)
            learnable[(i, pname)] = module.param
    return learnable


def grads_of(loss, learnable, inputs):
    for p in learnable.values():
        p.grad = None
    for t in inputs:
        t.grad = None
    loss.backward()
    out = {}
    for (i, pname), p in learnable.items():
        if p.grad is not None:
            out[f"{pname}.{i}"] = p.grad.detach().numpy().copy()
    return out, [None if t.grad is None else t.grad.detach().numpy().copy() for t in inputs]


def main():
    torch.manual_seed(0)
    for rel, links in ROBOTS:
        path = os.path.join(DATA, rel)
        rng = np.random.RandomState(0)
        model = quiet_model(path)
        q, qd, qdd = sample(model, rng)
        n, N = model._n_dofs, len(model._bodies)
        out = {"q": q.numpy(), "qd": qd.numpy(), "qdd": qdd.numpy(),
               "link_names": np.array([b.name for b in model._bodies]),
               "fk_links": np.array(links)}

        # link parameters as parsed by the reference
        names = [b.name for b in model._bodies]
        out["parent"] = np.array([-1] + [names.index(model._urdf_model.get_name_of_parent_body(b.name))
                                        for b in model._bodies[1:]], dtype=np.int32)
        out["dof"] = np.array([-1 if b.joint_idx is None else b.joint_idx for b in model._bodies], dtype=np.int32)
        out["axis"] = np.stack([b.joint_axis.reshape(3).numpy() for b in model._bodies])
        out["trans"] = np.stack([b.trans().reshape(3).numpy() for b in model._bodies])
        out["rpy"] = np.stack([b.rot_angles().reshape(3).numpy() for b in model._bodies])
        out["mass"] = np.stack([b.inertia.mass().reshape(()).numpy() for b in model._bodies])
        out["com"] = np.stack([b.inertia.com().reshape(3).numpy() for b in model._bodies])
        out["inertia"] = np.stack([b.inertia.inertia_mat().reshape(3, 3).numpy() for b in model._bodies])
        out["damping"] = np.array([0.0 if b.joint_damping() is None else float(b.joint_damping())
                                   for b in model._bodies], dtype=np.float32)
        lim = model.get_joint_limits()
        out["limits"] = np.array([[l["lower"], l["upper"], l["velocity"], l["effort"]] for l in lim])

        with torch.no_grad():
            model.update_kinematic_state(q, qd)
            out["all_R"] = np.stack([np.broadcast_to(b.pose.rotation().numpy(), (BATCH, 3, 3)) for b in model._bodies])
            out["all_p"] = np.stack([np.broadcast_to(b.pose.translation().numpy(), (BATCH, 3)) for b in model._bodies])
            for link in links:
                pos, quat = quiet_model(path).compute_forward_kinematics(q, link)
                jl, ja = quiet_model(path).compute_endeffector_jacobian(q, link)
                out[f"pos.{link}"], out[f"quat.{link}"] = pos.numpy(), quat.numpy()
                out[f"jlin.{link}"], out[f"jang.{link}"] = jl.numpy(), ja.numpy()
            for grav in (0, 1):
                for damp in (0, 1):
                    tau = quiet_model(path).compute_inverse_dynamics(q, qd, qdd, include_gravity=bool(grav),
                                                                     use_damping=bool(damp))
                    out[f"tau.g{grav}d{damp}"] = tau.numpy()

        # ---- gradients through the reference's autograd graph -------------------------------------
        gen = torch.Generator().manual_seed(1234)
        G_pos = torch.randn(BATCH, 3, generator=gen)
        G_jl = torch.randn(BATCH, 3, n, generator=gen)
        G_ja = torch.randn(BATCH, 3, n, generator=gen)
        G_tau = torch.randn(BATCH, n, generator=gen)
        out["G_pos"], out["G_jl"], out["G_ja"], out["G_tau"] = G_pos.numpy(), G_jl.numpy(), G_ja.numpy(), G_tau.numpy()

        for link in links[:2]:
            model = quiet_model(path)
            learnable = make_all_learnable(model)
            qg = q.clone().requires_grad_(True)
            pos, _ = model.compute_forward_kinematics(qg, link)
            jl, ja = model.compute_endeffector_jacobian(qg, link)
            loss = (G_pos * pos).sum() + (G_jl * jl).sum() + (G_ja * ja).sum()
            pg, (dq,) = grads_of(loss, learnable, [qg])
            out[f"fkgrad.{link}.q"] = dq
            for k, v in pg.items():
                out[f"fkgrad.{link}.{k}"] = v

        model = quiet_model(path)
        learnable = make_all_learnable(model)
        qg, qdg, qddg = (t.clone().requires_grad_(True) for t in (q, qd, qdd))
        tau = model.compute_inverse_dynamics(qg, qdg, qddg, include_gravity=True, use_damping=True)
        loss = (G_tau * tau).sum()
        pg, (dq, dqd, dqdd) = grads_of(loss, learnable, [qg, qdg, qddg])
        out["idgrad.q"], out["idgrad.qd"], out["idgrad.qdd"] = dq, dqd, dqdd
        for k, v in pg.items():
            out[f"idgrad.{k}"] = v

        stem = os.path.splitext(os.path.basename(rel))[0]
        dst = os.path.join(HERE, f"{stem}.npz")
        np.savez_compressed(dst, **out)
        print(f"{stem}: N={N} n={n} keys={len(out)} -> {os.path.relpath(dst, REPO)} ({os.path.getsize(dst)} B)")


if __name__ == "__main__":
    main()
