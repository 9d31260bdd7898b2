"""GPU: the thin callers of the hot path: mass matrix, non-linear effects, forward dynamics,
all-links FK -- against the fp64 oracle and through round trips."""
import numpy as np
import pytest
import torch

from conftest import assert_close, canon_quat, load_golden, urdf_path
import differentiable_robot_model_b200 as drm
from oracle import drm_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def oracle_mass_matrix(robot, q):
    z = torch.zeros_like(q)
    g = O.inverse_dynamics(robot, q, z, z, True, False)
    cols = []
    for j in range(robot.n_dofs):
        e = z.clone()
        e[:, j] = 1
        cols.append(O.inverse_dynamics(robot, q, z, e, True, False) - g)
    return torch.stack(cols, dim=2)


# The hand is tested with the *small_damping* URDF, like the reference's own tests
# (tests/test_kinematics_dynamics.py:30-39): with damping 3..10 N m s the term d*qd (~10) swamps H*qdd (~1e-6)
# in fp32, so tau -> qdd is ill-posed at the input (the fp32 CPU oracle loses the same 0.15 rad/s^2).
@pytest.mark.parametrize("stem", ["iiwa7", "panda_no_gripper", "allegro_hand_description_left_small_damping",
                                  "trifinger_edu", "2link_robot"])
def test_mass_matrix_and_forward_dynamics(stem):
    m = drm.DifferentiableRobotModel(urdf_path(stem), stem, device=DEV)
    robot = O.load_robot(urdf_path(stem), torch.float64)
    q, qd, qdd = O.sample_inputs(robot, 257, seed=2, dtype=torch.float64)
    qg, qdg, qddg = (t.float().to(DEV) for t in (q, qd, qdd))
    H = m.compute_lagrangian_inertia_matrix(qg)
    Ho = oracle_mass_matrix(robot, qg.cpu().double())
    # reference tolerance for the mass matrix vs pybullet: rtol 1e-3, atol 1e-5 (tests/test_kinematics_dynamics.py:407-409)
    assert_close(H.cpu().numpy(), Ho.numpy(), rtol=1e-4, atol=2e-5 * float(Ho.abs().max()), what="H")
    assert H.shape == (257, robot.n_dofs, robot.n_dofs)
    for grav, damp in ((True, False), (True, True), (False, False)):
        tau = m.compute_inverse_dynamics(qg, qdg, qddg, include_gravity=grav, use_damping=damp)
        f = tau.clone()
        back = m.compute_forward_dynamics(qg, qdg, f, include_gravity=grav, use_damping=damp)
        assert torch.equal(f, tau)                              # inputs are not modified
        # reference tolerance for forward dynamics: rtol 1e-2, atol 1e-3 (tests/test_kinematics_dynamics.py:503)
        assert_close(back.cpu().numpy(), qddg.cpu().numpy(), rtol=1e-2, atol=1e-3 * max(1.0, float(qddg.abs().max())), what="fd round trip")
    # 1-D inputs squeeze
    assert m.compute_lagrangian_inertia_matrix(qg[0]).shape == (robot.n_dofs, robot.n_dofs)
    assert m.compute_forward_dynamics(qg[0], qdg[0], qddg[0]).shape == (robot.n_dofs,)


def test_all_links_fk_matches_reference_golden(robot_stem):
    g = load_golden(robot_stem)
    m = drm.DifferentiableRobotModel(urdf_path(robot_stem), robot_stem, device=DEV)
    q = torch.tensor(g["q"], device=DEV)
    poses = m.compute_forward_kinematics_all_links(q)
    names = g["link_names"].tolist()
    assert list(poses) == names
    robot = O.load_robot(urdf_path(robot_stem), torch.float64)
    for i, name in enumerate(names):
        pos, quat = poses[name]
        assert_close(pos.cpu().numpy(), g["all_p"][i], what=f"pos {name}")
        want = O.quaternion(torch.tensor(g["all_R"][i], dtype=torch.float64)).numpy()
        assert_close(canon_quat(quat.cpu().numpy()), canon_quat(want), atol=2e-6, what=f"quat {name}")
    # recursive=True gives the (correct) non-recursive answer
    last = names[-1]
    a = m.compute_forward_kinematics(q, last, recursive=True)
    b = m.compute_forward_kinematics(q, last)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])


def test_forward_dynamics_is_differentiable():
    m = drm.DifferentiableKUKAiiwa(device=DEV)
    robot = O.load_robot(m.urdf_path, torch.float32)
    q, qd, _ = (t.to(DEV) for t in O.sample_inputs(robot, 64, seed=4))
    f = torch.randn(64, 7, device=DEV, requires_grad=True)
    qdd = m.compute_forward_dynamics(q, qd, f)
    qdd.sum().backward()
    # d qdd / d f = H^-1 (symmetric positive definite): row sums of H^-1
    H = m.compute_lagrangian_inertia_matrix(q)
    want = torch.linalg.inv(H.double()).sum(1).float()
    assert float((f.grad - want).abs().max()) < 1e-2 * float(want.abs().max())
