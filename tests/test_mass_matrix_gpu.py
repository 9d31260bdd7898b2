"""GPU: the one-launch mass-matrix kernel (drmb200_mass_matrix, csrc/mass_matrix.cu) against

  * the reference's construction evaluated with the fp64 oracle (column j = ID(q, 0, e_j) - ID(q, 0, 0),
    robot_model.py:403-450), all robot families incl. trees, fixed links and non-symmetric inertias;
  * the same construction through the RNEA kernel (compute_lagrangian_inertia_matrix_stacked);
and its gradients (RNEA adjoint over the stacked columns) against autograd through that stacked path.
Tolerance: the reference's own for the mass matrix vs pybullet is rtol 1e-3 / atol 1e-5
(tests/test_kinematics_dynamics.py:407-409); here rtol 1e-4 and 2e-5 x the largest entry.
"""
import numpy as np
import pytest
import torch

from conftest import assert_close, urdf_path
import differentiable_robot_model_b200 as drm
from differentiable_robot_model_b200 import engine
from test_backward_gpu import learnable_model
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


@pytest.mark.parametrize("stem", ["iiwa7", "panda", "allegro_hand_description_left", "trifinger_edu", "2link_robot",
                                  "jaco_clean", "iiwa7_allegro", "fetch_arm_no_gripper"])
def test_mass_matrix_matches_oracle_and_stacked_rnea(stem):
    m = drm.DifferentiableRobotModel(urdf_path(stem), stem, device=DEV)
    robot = O.load_robot(urdf_path(stem), torch.float64)
    q = O.sample_inputs(robot, 259, seed=6, dtype=torch.float64)[0]          # ragged tile
    qg = q.float().to(DEV)
    m._link_table()
    launches = engine.launch_count()
    H = m.compute_lagrangian_inertia_matrix(qg)
    assert engine.launch_count() - launches == 1
    assert H.shape == (259, robot.n_dofs, robot.n_dofs)
    Ho = oracle_mass_matrix(robot, qg.cpu().double())
    scale = float(Ho.abs().max())
    assert_close(H.cpu().numpy(), Ho.numpy(), rtol=1e-4, atol=2e-5 * scale, what="H vs oracle")
    for grav, damp in ((True, True), (False, False)):
        Hs = m.compute_lagrangian_inertia_matrix_stacked(qg, include_gravity=grav, use_damping=damp)
        assert_close(H.cpu().numpy(), Hs.cpu().numpy(), rtol=1e-4, atol=2e-5 * scale, what="H vs stacked RNEA")
    # symmetric (symmetric link inertias) and positive definite
    assert float((H - H.transpose(1, 2)).abs().max()) < 1e-5 * scale
    assert float(torch.linalg.eigvalsh(H.double().cpu()).min()) > 0


def test_mass_matrix_nonsymmetric_inertia_and_edge_cases():
    stem = "iiwa7"
    robot = O.load_robot(urdf_path(stem), torch.float64)
    gen = torch.Generator().manual_seed(3)
    scale = robot.inertia.abs().amax(dim=(1, 2), keepdim=True)
    robot.inertia = (robot.inertia + 0.05 * scale * torch.randn(robot.inertia.shape, generator=gen, dtype=torch.float64)).float().double()
    m = drm.DifferentiableRobotModel(urdf_path(stem), stem, device=DEV)
    q = O.sample_inputs(robot, 64, seed=1, dtype=torch.float64)[0]
    H = engine.mass_matrix_raw(m._topology, O.link_table(robot).float().to(DEV), q.float().to(DEV))
    Ho = oracle_mass_matrix(robot, q.float().double())
    assert_close(H.cpu().numpy(), Ho.numpy(), rtol=1e-4, atol=2e-5 * float(Ho.abs().max()), what="H nonsymmetric")
    assert float((Ho - Ho.transpose(1, 2)).abs().max()) > 1e-4           # the case really is non-symmetric
    # empty batch, 1-D input
    assert m.compute_lagrangian_inertia_matrix(torch.zeros(0, 7, device=DEV)).shape == (0, 7, 7)
    assert m.compute_lagrangian_inertia_matrix(q[0].float().to(DEV)).shape == (7, 7)


@pytest.mark.parametrize("stem", ["iiwa7", "trifinger_edu"])
def test_mass_matrix_gradients_match_stacked_autograd(stem):
    robot = O.load_robot(urdf_path(stem), torch.float32)
    q = O.sample_inputs(robot, 130, seed=8)[0]
    G = torch.randn(130, robot.n_dofs, robot.n_dofs, generator=torch.Generator().manual_seed(4)).to(DEV)
    grads = []
    for method in ("compute_lagrangian_inertia_matrix", "compute_lagrangian_inertia_matrix_stacked"):
        m, params = learnable_model(stem)
        qg = q.to(DEV).requires_grad_(True)
        H = getattr(m, method)(qg)
        (G * H).sum().backward()
        grads.append((qg.grad.clone(), {k: (torch.zeros_like(p) if p.grad is None else p.grad.clone()) for k, p in params.items()}))
    (dq_a, pa), (dq_b, pb) = grads
    scale = max(float(dq_b.abs().max()), max(float(v.abs().max()) for v in pb.values()))
    np.testing.assert_allclose(dq_a.cpu().numpy(), dq_b.cpu().numpy(), rtol=2e-3, atol=2e-5 * scale)
    for k in pb:
        np.testing.assert_allclose(pa[k].cpu().numpy(), pb[k].cpu().numpy(), rtol=2e-3, atol=2e-5 * scale, err_msg=str(k))
