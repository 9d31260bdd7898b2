"""GPU: the parameter-learning / trajectory-optimisation examples run and make progress (the reference only checks
"does not raise", tests/test_examples.py:22-35; its dynamics example even trains on an all-zero trajectory at
n_data=250 because of int(n_data * dt) == 1)."""
import os
import sys

import pytest

from conftest import REPO

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(REPO, "examples"))


def test_learn_kinematics_of_iiwa():
    import learn_kinematics_of_iiwa as ex
    hist = ex.run(n_epochs=300, n_data=100, device="cuda:0")
    assert hist[-1] < 0.5 * hist[0]


def test_learn_dynamics_iiwa():
    import learn_dynamics_iiwa as ex
    hist = ex.run(n_epochs=3, n_data=1000, device="cuda:0")     # n_data * dt = 4 s of sine motion
    assert len(hist) == 3 and all(h == h for h in hist) and hist[-1] < hist[0]


def test_kinematic_trajectory_opt():
    import run_kinematic_trajectory_opt as ex
    hist = ex.run(n_iters=150, n_targets=2048, device="cuda:0")
    assert hist[-1] < 0.05 * hist[0]


def test_learn_forward_dynamics_iiwa():
    import learn_forward_dynamics_iiwa as ex
    hist = ex.run(n_epochs=4, n_data=2000, device="cuda:0")     # n_data * dt = 8 s of sine motion
    assert len(hist) == 4 and all(h == h for h in hist) and hist[-1] < hist[0]


def test_learn_kinematics_of_toy():
    import learn_kinematics_of_toy as ex
    hist = ex.run(n_epochs=600, n_data=100, device="cuda:0")
    assert hist[-1] < 0.5 * hist[0]


def test_learn_dynamics_iiwa_sharded_single_rank():
    """The sharded-batch example (fused flat parameter, fused all-reduce + Adam kernel, one CUDA graph per iteration) on one
    rank; tests/test_sharded_gpu.py covers the N >= 2 exchange."""
    import learn_dynamics_iiwa_sharded as ex
    hist = ex.run(n_iters=150, n_data=8192, log=lambda *_: None)
    assert all(h == h for h in hist) and hist[-1] < 0.5 * hist[0]
