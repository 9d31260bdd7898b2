"""GPU: the fused link-table build (csrc/table.cu) against the batched-torch assembly used on the CPU, values and
gradients, for every shipped URDF."""
import pytest
import torch

from conftest import assert_close, urdf_path
import differentiable_robot_model_b200 as drm
from differentiable_robot_model_b200 import engine, link_table

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_fused_table_matches_torch_assembly(robot_stem):
    gpu = drm.DifferentiableRobotModel(urdf_path(robot_stem), robot_stem, device=DEV)
    cpu = drm.DifferentiableRobotModel(urdf_path(robot_stem), robot_stem, device="cpu")
    launches = engine.launch_count()
    t_gpu = gpu._link_table()
    assert engine.launch_count() - launches == 1                 # one kernel, not ~60 torch launches
    assert_close(t_gpu.cpu().numpy(), cpu._link_table().numpy(), rtol=1e-6, atol=1e-7, what="table")

    # gradient: random perturbation of every raw parameter, random upstream gradient
    torch.manual_seed(0)
    raw = link_table.gather_raw_parameters(gpu._bodies, gpu._device)
    raw = (raw + 0.1 * torch.randn_like(raw)).requires_grad_(True)
    G = torch.randn(raw.shape[0], 28, device=DEV)
    (engine.BuildLinkTableFunction.apply(raw) * G).sum().backward()

    r64 = raw.detach().cpu().double().requires_grad_(True)
    rpy, trans, mass, com, inertia, damping = r64[:, 0:3], r64[:, 3:6], r64[:, 6], r64[:, 7:10], r64[:, 10:19], r64[:, 19]
    F = link_table._rpy_to_matrix(rpy)
    cx, cy, cz = com[:, 0], com[:, 1], com[:, 2]
    ssT = torch.stack([cy * cy + cz * cz, -cx * cy, -cx * cz, -cx * cy, cx * cx + cz * cz, -cy * cz,
                       -cx * cz, -cy * cz, cx * cx + cy * cy], dim=1)
    table = torch.cat([F, trans, inertia + mass[:, None] * ssT, mass[:, None] * com, mass[:, None], damping[:, None],
                       torch.zeros(r64.shape[0], 2, dtype=torch.float64)], dim=1)
    assert_close(engine.BuildLinkTableFunction.apply(raw).detach().cpu().numpy(), table.detach().numpy(), rtol=1e-6,
                 atol=1e-6, what="table from perturbed raw")
    (table * G.cpu().double()).sum().backward()
    assert_close(raw.grad.cpu().numpy(), r64.grad.numpy(), rtol=1e-5, atol=1e-5, what="raw grad")
