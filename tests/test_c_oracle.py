"""CPU: pin the C oracle (oracle/drm_oracle.c, float and double builds) against the golden vectors
generated from the reference itself, and against the torch oracle on a larger seeded batch."""
import numpy as np
import pytest
import torch

from conftest import assert_close, canon_quat, load_golden, urdf_path
from oracle import drm_oracle as O
from oracle.c_oracle import CRobot


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_c_oracle_matches_reference_golden(robot_stem, dtype):
    g = load_golden(robot_stem)
    robot = O.load_robot(urdf_path(robot_stem), torch.float64)
    c = CRobot(robot, dtype)
    for link in g["fk_links"].tolist():
        pos, quat, jl, ja = c.fk_jacobian(robot.index(link), g["q"])
        assert_close(pos, g[f"pos.{link}"], what=f"pos {link}")
        assert_close(canon_quat(quat), canon_quat(g[f"quat.{link}"]), what=f"quat {link}")
        assert_close(jl, g[f"jlin.{link}"], what=f"jlin {link}")
        assert_close(ja, g[f"jang.{link}"], what=f"jang {link}")
    for grav in (0, 1):
        for damp in (0, 1):
            tau = c.inverse_dynamics(g["q"], g["qd"], g["qdd"], grav, damp)
            scale = float(np.abs(g[f"tau.g{grav}d{damp}"]).max())
            assert_close(tau, g[f"tau.g{grav}d{damp}"], rtol=1e-5, atol=max(1e-5, 2e-6 * scale), what=f"tau g{grav}d{damp}")


def test_c_oracle_matches_torch_oracle_and_is_thread_count_independent():
    robot = O.load_robot(urdf_path("allegro_hand_description_left"), torch.float64)
    q, qd, qdd = O.sample_inputs(robot, 3001, seed=5, dtype=torch.float64)
    c = CRobot(robot, np.float64)
    e = robot.index("link_3.0_tip")
    outs1 = c.fk_jacobian(e, q.numpy(), n_threads=1)
    outs8 = c.fk_jacobian(e, q.numpy(), n_threads=8)
    for a, b in zip(outs1, outs8):
        np.testing.assert_array_equal(a, b)
    pos, quat = O.forward_kinematics(robot, q, "link_3.0_tip")
    jl, ja = O.jacobian(robot, q, "link_3.0_tip")
    assert_close(outs1[0], pos.numpy(), rtol=1e-10, atol=1e-12, what="pos")
    assert_close(canon_quat(outs1[1]), canon_quat(quat.numpy()), rtol=1e-10, atol=1e-12, what="quat")
    assert_close(outs1[2], jl.numpy(), rtol=1e-10, atol=1e-12, what="jlin")
    assert_close(outs1[3], ja.numpy(), rtol=1e-10, atol=1e-12, what="jang")
    tau = c.inverse_dynamics(q.numpy(), qd.numpy(), qdd.numpy())
    assert_close(tau, O.inverse_dynamics(robot, q, qd, qdd).numpy(), rtol=1e-9, atol=1e-10, what="tau")


def test_per_element_quaternion_loop_matches_vectorised_and_reference_golden():
    """oracle.quaternion_per_element restates the reference's shipped Python loop; it must agree with the vectorised
    oracle (same branches, same signs) and with the reference's own outputs."""
    g = load_golden("iiwa7_allegro")
    robot = O.load_robot(urdf_path("iiwa7_allegro"), torch.float32)
    q = torch.tensor(g["q"])
    R, _, _, _, _ = O.kinematic_state(robot, q)
    for link in g["fk_links"].tolist():
        Ri = R[robot.index(link)]
        loop, vec = O.quaternion_per_element(Ri), O.quaternion(Ri)
        assert_close(loop.numpy(), vec.numpy(), rtol=1e-6, atol=1e-7, what="loop vs vectorised")
        assert_close(canon_quat(loop.numpy()), canon_quat(g[f"quat.{link}"]), what="loop vs reference")
    Rrand = torch.linalg.qr(torch.randn(257, 3, 3, generator=torch.Generator().manual_seed(0)))[0]
    Rrand = Rrand * torch.sign(torch.linalg.det(Rrand))[:, None, None]            # proper rotations, all four branches
    assert_close(O.quaternion_per_element(Rrand).numpy(), O.quaternion(Rrand).numpy(), rtol=1e-5, atol=1e-6, what="random R")


@pytest.mark.parametrize("dtype,tol", [(np.float32, 2e-4), (np.float64, 2e-4)])
def test_c_forward_dynamics_matches_reference_golden(robot_stem, dtype, tol):
    """C restatement of the articulated-body algorithm vs the reference's own fp32 outputs (tests/golden/*.fd.npz),
    per configuration relative to its largest acceleration (the hand models divide by inertias ~1e-7)."""
    import os
    from conftest import GOLDEN_DIR
    g = np.load(os.path.join(GOLDEN_DIR, robot_stem + ".fd.npz"))
    robot = O.load_robot(urdf_path(robot_stem), torch.float64)
    cr = CRobot(robot, dtype)
    for grav in (0, 1):
        for damp in (0, 1):
            got = cr.forward_dynamics(g["q"], g["qd"], g["f"], bool(grav), bool(damp), n_threads=2)
            want = g[f"qdd.g{grav}d{damp}"]
            assert np.all(np.abs(got - want) <= tol * np.abs(want).max(axis=1, keepdims=True) + 1e-6), (grav, damp)


def test_c_forward_dynamics_matches_torch_oracle_nonsymmetric():
    robot = O.load_robot(urdf_path("iiwa7"), torch.float64)
    gen = torch.Generator().manual_seed(5)
    robot.inertia = robot.inertia + 0.05 * robot.inertia.abs().amax(dim=(1, 2), keepdim=True) * torch.randn(
        robot.inertia.shape, generator=gen, dtype=torch.float64)
    q, qd, _ = O.sample_inputs(robot, 64, seed=8, dtype=torch.float64)
    f = torch.randn(64, 7, generator=gen, dtype=torch.float64)
    want = O.forward_dynamics(robot, q, qd, f, True, True).numpy()
    got = CRobot(robot, np.float64).forward_dynamics(q.numpy(), qd.numpy(), f.numpy(), True, True)
    np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-9 * np.abs(want).max())
