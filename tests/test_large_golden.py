"""CPU: the oracle against the 1024-row (Allegro: 512-row) reference outputs of tests/golden/large_*.npz, including the
RAW quaternion sign on every row that is clear of a branch boundary -- all four branches of the reference's
get_quaternion occur >= 170 times at the Kuka and Panda end effectors."""
import numpy as np
import pytest
import torch

from conftest import LARGE_GOLDEN, assert_close, load_golden, quat_branch_margin, urdf_path
from oracle import drm_oracle as O


@pytest.mark.parametrize("stem", sorted(LARGE_GOLDEN))
def test_oracle_matches_large_reference_batches(stem):
    g = load_golden(stem)
    robot = O.load_robot(urdf_path(LARGE_GOLDEN[stem]), torch.float32)
    q, qd, qdd = (torch.tensor(g[k]) for k in ("q", "qd", "qdd"))
    for link in g["links"].tolist():
        pos, quat = O.forward_kinematics(robot, q, link)
        jl, ja = O.jacobian(robot, q, link)
        assert_close(pos.numpy(), g[f"pos.{link}"], what=f"{stem} pos {link}")
        assert_close(jl.numpy(), g[f"jlin.{link}"], what=f"{stem} jlin {link}")
        assert_close(ja.numpy(), g[f"jang.{link}"], what=f"{stem} jang {link}")
        clear = quat_branch_margin(g[f"R.{link}"]) > 1e-3
        assert clear.mean() > 0.9
        assert_close(quat.numpy()[clear], g[f"quat.{link}"][clear], atol=2e-6, what=f"{stem} raw quat {link}")
    tau = O.inverse_dynamics(robot, q, qd, qdd, True, True)
    assert_close(tau.numpy(), g["tau"], atol=1e-5, what=f"{stem} tau")
    if stem != "large_allegro_left":
        assert np.bincount(g["branch"], minlength=4).min() >= 48        # every quaternion branch is exercised
