"""GPU: the all-links kernel behind ``update_kinematic_state`` (per-body ``pose`` / ``vel``) against the reference's
golden all-link poses and the fp64 oracle's body-frame velocities, for every shipped URDF (chains and trees, all six
signed joint axes)."""
import numpy as np
import pytest
import torch

from conftest import assert_close, load_golden, urdf_path
import differentiable_robot_model_b200 as drm
from differentiable_robot_model_b200 import engine
from oracle import drm_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_update_kinematic_state_matches_reference_and_oracle(robot_stem):
    g = load_golden(robot_stem)
    m = drm.DifferentiableRobotModel(urdf_path(robot_stem), robot_stem, device=DEV)
    with pytest.raises(RuntimeError, match="update_kinematic_state"):
        m._bodies[1].pose
    q, qd = torch.tensor(g["q"], device=DEV), torch.tensor(g["qd"], device=DEV)
    m._link_table()
    launches = engine.launch_count()
    assert m.update_kinematic_state(q, qd) is None
    assert engine.launch_count() - launches == 1
    robot = O.load_robot(urdf_path(robot_stem), torch.float64)
    _, _, w, v, _ = O.kinematic_state(robot, q.cpu().double(), qd.cpu().double())
    for i, body in enumerate(m._bodies):
        pose = body.pose
        assert pose.rotation().shape == (q.shape[0], 3, 3) and pose.translation().shape == (q.shape[0], 3)
        assert_close(pose.rotation().cpu().numpy(), g["all_R"][i], what=f"R {body.name}")
        assert_close(pose.translation().cpu().numpy(), g["all_p"][i], what=f"p {body.name}")
        vel = body.vel
        scale = max(1.0, float(torch.stack(v).abs().max()))
        assert_close(vel.ang.cpu().numpy(), w[i].numpy(), atol=2e-6 * scale, what=f"ang vel {body.name}")
        assert_close(vel.lin.cpu().numpy(), v[i].numpy(), atol=2e-6 * scale, what=f"lin vel {body.name}")
    # the value types compose like the reference's: pose_i = pose_parent o joint_pose, checked through inverse()
    last = m._bodies[-1].pose
    ident = last.multiply_transform(last.inverse())
    assert float((ident.rotation() - torch.eye(3, device=DEV)).abs().max()) < 1e-5
    assert float(ident.translation().abs().max()) < 1e-5


def test_ragged_and_large_batches():
    m = drm.DifferentiableRobotModel(urdf_path("iiwa7_allegro"), "hand_arm", device=DEV)
    robot = O.load_robot(urdf_path("iiwa7_allegro"), torch.float64)
    for batch in (1, 127, 129, 4099):
        q, qd, _ = O.sample_inputs(robot, batch, seed=batch, dtype=torch.float64)
        m.update_kinematic_state(q.float().to(DEV), qd.float().to(DEV))
        R, p, w, v, _ = O.kinematic_state(robot, q.float().double(), qd.float().double())
        i = len(m._bodies) - 1
        assert_close(m._bodies[i].pose.translation().cpu().numpy(), p[i].numpy(), what="p")
        assert_close(m._bodies[i].pose.rotation().cpu().numpy(), R[i].numpy(), what="R")
        assert_close(m._bodies[i].vel.ang.cpu().numpy(), w[i].numpy(), atol=1e-5, what="w")
    # 1-D inputs are accepted like everywhere else
    m.update_kinematic_state(torch.zeros(23, device=DEV), torch.zeros(23, device=DEV))
    assert m._bodies[3].pose.translation().shape == (1, 3)
