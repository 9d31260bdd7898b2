"""Inertia / scalar parametrisations against golden vectors produced by the reference's own classes
(tests/golden/make_params_golden.py; reference: rigid_body_params.py:14-403).  CPU only."""
import os

import numpy as np
import pytest
import torch

from differentiable_robot_model_b200 import rigid_body_params as P

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "params.npz"))
NETS = ("SymmPosDef3DInertiaMatrixNet", "CovParameterized3DInertiaMatrixNet", "Symm3DInertiaMatrixNet")


@pytest.mark.parametrize("name", NETS)
def test_init_from_inertia_matrix(name):
    net = getattr(P, name)(init_param=torch.tensor(G["init_inertia"]))
    np.testing.assert_allclose(net.l.detach().numpy(), G[f"{name}.init_l"], rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(net().detach().numpy(), G[f"{name}.init_out"], rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(net().detach().numpy(), G["init_inertia"][0], rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("name", NETS)
def test_forward_and_gradient(name):
    net = getattr(P, name)()
    with torch.no_grad():
        net.l.copy_(torch.tensor(G[f"{name}.l"]))
    val = net()
    (val * torch.tensor(G["weight"])).sum().backward()
    np.testing.assert_allclose(val.detach().numpy(), G[f"{name}.out"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(net.l.grad.numpy(), G[f"{name}.grad"], rtol=1e-5, atol=1e-6)


def test_batched_nets():
    x = torch.tensor(G["rows"])
    np.testing.assert_allclose(P.SymmMatNet(3)(x).numpy(), G["SymmMatNet.out"], rtol=1e-6, atol=1e-7)
    spsd, l = P.CholeskyNet(3, 0.25).get_symm_pos_semi_def_matrix_and_l(x)
    np.testing.assert_allclose(spsd.numpy(), G["CholeskyNet.out"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(l.numpy(), G["CholeskyNet.l"], rtol=1e-6, atol=1e-7)


def test_constructions_are_physical():
    torch.manual_seed(3)
    for _ in range(5):
        spd = P.SymmPosDef3DInertiaMatrixNet(init_param_std=0.5)()
        assert torch.linalg.eigvalsh(spd).min() > 0
        cov = P.CovParameterized3DInertiaMatrixNet(init_param_std=0.5)()
        ev = torch.linalg.eigvalsh(cov)
        assert ev.min() > 0 and ev[0] + ev[1] >= ev[2] * (1 - 1e-5)      # triangle inequality of principal moments
        tri = P.TriangParam3DInertiaMatrixNet(bias=1e-4, init_param_std=0.5)()
        ev = torch.linalg.eigvalsh(tri)
        assert ev.min() > 0 and ev[0] + ev[1] >= ev[2] * (1 - 1e-5)
        np.testing.assert_allclose(tri.detach().numpy(), tri.detach().numpy().T, atol=1e-7)


def test_triangular_net_round_trip():
    inertia = torch.tensor(G["init_inertia"])
    net = P.TriangParam3DInertiaMatrixNet(bias=1e-4, init_param=inertia)
    np.testing.assert_allclose(net().detach().numpy(), inertia[0].numpy(), rtol=1e-4, atol=1e-6)
    net().sum().backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in net.parameters())


def test_exp_map_matches_matrix_exponential():
    w = torch.tensor([0.3, -0.7, 0.5])
    hat = torch.tensor([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    np.testing.assert_allclose(P.exp_map_so3(w).numpy(), torch.linalg.matrix_exp(hat).numpy(), atol=1e-6)


def test_utils_helpers_match_reference_semantics():
    from differentiable_robot_model_b200 import utils as U
    a, b = torch.tensor([[1.0, 2.0, 3.0], [0.5, -1.0, 2.0]]), torch.tensor([[0.3, -0.2, 0.9], [1.0, 1.0, 1.0]])
    np.testing.assert_allclose(U.cross_product(a, b).numpy(), torch.linalg.cross(a, b).numpy(), atol=1e-6)
    S = U.vector3_to_skew_symm_matrix(a[0])
    assert S.shape == (1, 3, 3) and torch.equal(S, -S.transpose(1, 2))
    A = U.bfill_diagonal(U.bfill_lowertriangle(torch.zeros(2, 3, 3), torch.tensor([1.0, 2.0, 3.0])), torch.tensor([7.0, 8.0, 9.0]))
    assert A[1, 1, 0] == 1 and A[1, 2, 0] == 2 and A[1, 2, 1] == 3 and A[0, 2, 2] == 9 and A[0, 0, 1] == 0
    assert U.convert_into_at_least_2d_pytorch_tensor([1.0, 2.0]).shape == (1, 2)
    w = torch.tensor([0.3, -0.7, 0.5])
    np.testing.assert_allclose(U.exp_map_so3(w).numpy(), P.exp_map_so3(w).numpy(), atol=1e-7)
