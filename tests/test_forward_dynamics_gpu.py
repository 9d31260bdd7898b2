"""GPU: the articulated-body forward-dynamics kernel (drmb200_forward_dynamics, csrc/aba.cu) against

  * golden vectors produced by the reference's compute_forward_dynamics (tests/golden/*.fd.npz,
    robot_model.py:488-624), all four (include_gravity, use_damping) combinations, every shipped URDF;
  * the CPU oracle (oracle/drm_oracle.py: forward_dynamics) on seeded batches, with symmetric and
    NON-symmetric inertia matrices (the reference never symmetrises);
  * the inverse-dynamics kernel: ID(q, qd, FD(q, qd, f)) == f for symmetric inertias.

Tolerance: forward dynamics divides by articulated inertias that are tiny for the hand models, so errors are
judged per configuration relative to that configuration's largest acceleration: 2e-4 against the reference's own
fp32 output (two different fp32 evaluation orders), 1e-4 against the fp64 oracle for the arm models.
"""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR, urdf_path
import differentiable_robot_model_b200 as drm
from differentiable_robot_model_b200 import engine
from oracle import drm_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ARMS = ["iiwa7", "panda_no_gripper", "panda", "fetch_arm_no_gripper", "2link_robot", "trifinger_edu"]


def cuda(a):
    return torch.tensor(np.asarray(a), dtype=torch.float32, device=DEV)


def rowwise_err(got, want):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    return (np.abs(got - want) / (np.abs(want).max(axis=1, keepdims=True) + 1e-30)).max()


def load_fd(stem):
    return np.load(os.path.join(GOLDEN_DIR, stem + ".fd.npz"), allow_pickle=False)


def test_forward_dynamics_matches_reference_golden(robot_stem):
    g = load_fd(robot_stem)
    m = drm.DifferentiableRobotModel(urdf_path(robot_stem), robot_stem, device=DEV)
    q, qd, f = cuda(g["q"]), cuda(g["qd"]), cuda(g["f"])
    for grav in (0, 1):
        for damp in (0, 1):
            f_in = f.clone()
            qdd = m.compute_forward_dynamics(q, qd, f_in, include_gravity=bool(grav), use_damping=bool(damp))
            assert torch.equal(f_in, f), "the caller's f must not be modified"
            assert rowwise_err(qdd.cpu().numpy(), g[f"qdd.g{grav}d{damp}"]) < 2e-4, (robot_stem, grav, damp)


def oracle_case(stem, batch, nonsym, seed=3):
    robot = O.load_robot(urdf_path(stem), torch.float64)
    if nonsym:
        gen = torch.Generator().manual_seed(11)
        scale = robot.inertia.abs().amax(dim=(1, 2), keepdim=True).clamp_min(1e-6)
        robot.inertia = robot.inertia + 0.05 * scale * torch.randn(robot.inertia.shape, generator=gen, dtype=torch.float64)
    q, qd, _ = O.sample_inputs(robot, batch, seed=seed, dtype=torch.float64)
    f = torch.randn(batch, robot.n_dofs, generator=torch.Generator().manual_seed(seed), dtype=torch.float64)
    return robot, q, qd, f


@pytest.mark.parametrize("nonsym", [False, True], ids=["symmetric", "nonsymmetric"])
@pytest.mark.parametrize("stem", ARMS + ["iiwa7_allegro", "allegro_hand_description_left_small_damping"])
def test_forward_dynamics_matches_oracle(stem, nonsym):
    batch = 257                                           # ragged tile: exercises the cooperative-copy tail
    robot, q, qd, f = oracle_case(stem, batch, nonsym)
    topo = drm.DifferentiableRobotModel(urdf_path(stem), stem, device=DEV)._topology
    table = O.link_table(robot).to(torch.float32).to(DEV)
    for flags, (grav, damp) in ((engine.GRAVITY | engine.DAMPING, (True, True)), (0, (False, False))):
        want = O.forward_dynamics(robot, q, qd, f, grav, damp).numpy()
        got = engine.forward_dynamics_raw(topo, table, cuda(q), cuda(qd), cuda(f), flags).cpu().numpy()
        tol = 1e-4 if stem in ARMS else 2e-3              # hand models: articulated inertias ~1e-7, fp32 conditioning
        assert rowwise_err(got, want) < tol, (stem, nonsym, flags, rowwise_err(got, want))


def test_forward_then_inverse_dynamics_round_trip():
    m = drm.DifferentiableKUKAiiwa(device=DEV)
    robot = O.load_robot(m.urdf_path, torch.float32)
    q, qd, _ = O.sample_inputs(robot, 65536 + 3, seed=5)
    f = 5.0 * torch.randn(q.shape, generator=torch.Generator().manual_seed(1))
    q, qd, f = q.to(DEV), qd.to(DEV), f.to(DEV)
    qdd = m.compute_forward_dynamics(q, qd, f, include_gravity=True, use_damping=True)
    tau = m.compute_inverse_dynamics(q, qd, qdd, include_gravity=True, use_damping=True)
    err = (tau - f).abs().max().item()
    assert err < 2e-3 * max(1.0, f.abs().max().item()), err
    # and against the previous-generation path H^-1 (f - nle) built from the RNEA kernel
    qdd_solve = m.compute_forward_dynamics_crba(q, qd, f, include_gravity=True, use_damping=True)
    assert rowwise_err(qdd.cpu().numpy(), qdd_solve.cpu().numpy()) < 2e-3


def test_empty_batch_and_errors():
    m = drm.DifferentiableKUKAiiwa(device=DEV)
    z = torch.zeros(0, 7, device=DEV)
    assert m.compute_forward_dynamics(z, z, z).shape == (0, 7)
    with pytest.raises(Exception):
        m.compute_forward_dynamics(torch.zeros(4, 6, device=DEV), torch.zeros(4, 6, device=DEV), torch.zeros(4, 6, device=DEV))
