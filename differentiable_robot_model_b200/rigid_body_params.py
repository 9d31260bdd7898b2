"""
Link-parameter parametrisations
====================================
The small ``torch.nn.Module``s callers plug into ``make_link_param_learnable``.  They are host-side
producers of scalars / ``[1,3]`` / ``[3,3]`` tensors that ``link_table.build_link_table`` consumes; the
engine does not care which module produced a value.  All of the reference's parametrisation classes
(``rigid_body_params.py:14-403``) are mirrored: the three unconstrained / positive ones its examples use
(``examples/learn_dynamics_iiwa.py:57-65``, ``examples/learn_kinematics_of_iiwa.py:33-38``) and the
constrained 3-D inertia-matrix families (symmetric, Cholesky SPD, density-covariance, triangular principal
moments).  The inertia families evaluate with a handful of vectorised device ops (index gathers instead of the
reference's element-by-element fills) so that they stay cheap next to a microsecond-scale kernel.
"""
import math

import numpy as np
import torch

# position of (row, col) of the 6-vector [diag(3) | strictly-lower(3)] used by the reference
# (np.diag_indices + np.tril_indices(k=-1): (0,0) (1,1) (2,2) (1,0) (2,0) (2,1))
_ROWS = (0, 1, 2, 1, 2, 2)
_COLS = (0, 1, 2, 0, 0, 1)


class UnconstrainedScalar(torch.nn.Module):
    """A free scalar (reference: rigid_body_params.py:14-23)."""

    def __init__(self, init_val=None):
        super().__init__()
        self.param = torch.nn.Parameter(torch.rand(1) if init_val is None else init_val)

    def forward(self):
        return self.param


class PositiveScalar(torch.nn.Module):
    """``l^2 + min_val`` (reference: rigid_body_params.py:26-43)."""

    def __init__(self, min_val=0.0, init_param_std=1.0, init_param=None):
        super().__init__()
        self._min_val = min_val
        if init_param is None:
            start = torch.empty(1, 1).normal_(mean=0.0, std=init_param_std)
        else:
            start = torch.sqrt(init_param - self._min_val)
        self.l = torch.nn.Parameter(start.squeeze())

    def forward(self):
        return (self.l * self.l + self._min_val).squeeze()


class UnconstrainedTensor(torch.nn.Module):
    """A free ``[dim1, dim2]`` tensor (reference: rigid_body_params.py:46-56)."""

    def __init__(self, dim1, dim2, init_tensor=None, init_std=0.1):
        super().__init__()
        self._dim1, self._dim2 = dim1, dim2
        if init_tensor is None:
            init_tensor = torch.empty(dim1, dim2).normal_(mean=0.0, std=init_std)
        self.param = torch.nn.Parameter(init_tensor)

    def forward(self):
        return self.param


def _vec6_from_matrix(mat):
    m = torch.as_tensor(mat, dtype=torch.float32).reshape(3, 3)
    return m[list(_ROWS), list(_COLS)].clone()


def _lower_from_vec6(l):
    """[6] -> lower-triangular [3,3] (diagonal first, then (1,0) (2,0) (2,1))."""
    L = l.new_zeros(3, 3)
    return L.index_put((torch.tensor(_ROWS, device=l.device), torch.tensor(_COLS, device=l.device)), l)


class SymmMatNet(torch.nn.Module):
    """Symmetric ``[B,q,q]`` matrix from ``[B, q(q+1)/2]`` rows (reference: rigid_body_params.py:59-83)."""

    def __init__(self, qdim):
        self._qdim = qdim
        super().__init__()

    def forward(self, l):
        q = self._qdim
        ii, jj = np.tril_indices(q, k=-1)
        lower = l.new_zeros(l.shape[0], q, q)
        if q > 1:
            lower[:, ii, jj] = l[:, q:]
        return torch.diag_embed(l[:, :q]) + lower + lower.transpose(-2, -1)


class CholeskyNet(torch.nn.Module):
    """Symmetric positive semi-definite ``L L^T`` from the entries of L (reference: rigid_body_params.py:86-132)."""

    def __init__(self, qdim, bias):
        self._qdim = qdim
        self._bias = bias
        super().__init__()

    def get_raw_l(self, raw_l_input):
        return raw_l_input

    def get_l(self, raw_l_input):
        raw_l = self.get_raw_l(raw_l_input)
        shift = raw_l.new_zeros(raw_l.shape[-1])
        shift[: self._qdim] = self._bias              # positive bias on the diagonal of L
        return raw_l + shift

    def get_L(self, l):
        q = self._qdim
        ii, jj = np.tril_indices(q, k=-1)
        L = torch.diag_embed(l[:, :q])
        if q > 1:
            L[:, ii, jj] = l[:, q:]
        return L

    def get_symm_pos_semi_def_matrix_and_l(self, raw_l_input):
        l = self.get_l(raw_l_input)
        L = self.get_L(l)
        return L @ L.transpose(-2, -1), l


def _cholesky_init(mat, bias):
    L = np.linalg.cholesky(np.asarray(mat, dtype=np.float64).reshape(3, 3) - bias * np.eye(3))
    return torch.tensor(L[list(_ROWS), list(_COLS)], dtype=torch.float32)


class SymmPosDef3DInertiaMatrixNet(CholeskyNet):
    """``L L^T + bias * 1`` (reference: rigid_body_params.py:341-383)."""

    def __init__(self, bias=1e-7, init_param_std=0.01, init_param=None, is_initializing_params=True):
        super().__init__(qdim=3, bias=0)
        self.spd_3d_inertia_mat_diag_bias = bias
        if init_param is None or not is_initializing_params:
            start = torch.empty(6).normal_(mean=0.0, std=init_param_std)
        else:
            start = _cholesky_init(torch.as_tensor(init_param).detach().cpu().numpy(), bias)
        self.l = torch.nn.Parameter(start)

    def forward(self):
        L = _lower_from_vec6(self.l)
        return L @ L.t() + self.spd_3d_inertia_mat_diag_bias * torch.eye(3, device=self.l.device)


class CovParameterized3DInertiaMatrixNet(CholeskyNet):
    """Inertia matrix from the density-weighted covariance ``Sigma = L L^T + bias * 1`` of the body,
    ``I = tr(Sigma) 1 - Sigma`` (Wensing et al. 2017; reference: rigid_body_params.py:245-338)."""

    def __init__(self, bias=1.0e-7, init_param_std=0.01, init_param=None, is_initializing_params=True):
        super().__init__(qdim=3, bias=0)
        self.spd_3d_cov_inertia_mat_diag_bias = bias
        if init_param is None or not is_initializing_params:
            start = torch.empty(6).normal_(mean=0.0, std=init_param_std)
        else:
            inertia = torch.as_tensor(init_param).detach().cpu().double().reshape(3, 3)
            lower = torch.tril(inertia)                      # the reference reads the lower triangle only
            sym = lower + torch.tril(inertia, -1).t()
            cov = 0.5 * torch.trace(sym) * torch.eye(3, dtype=torch.float64) - sym
            start = _cholesky_init(cov.numpy(), bias)
        self.l = torch.nn.Parameter(start)

    def forward(self):
        L = _lower_from_vec6(self.l)
        eye = torch.eye(3, device=self.l.device)
        cov = L @ L.t() + self.spd_3d_cov_inertia_mat_diag_bias * eye
        return torch.trace(cov) * eye - cov


class Symm3DInertiaMatrixNet(SymmMatNet):
    """Free symmetric 3x3 (reference: rigid_body_params.py:386-403)."""

    def __init__(self, init_param_std=0.01, init_param=None, is_initializing_params=True):
        super().__init__(qdim=3)
        if init_param is None or not is_initializing_params:
            start = torch.empty(6).normal_(mean=0.0, std=init_param_std)
        else:
            start = _vec6_from_matrix(torch.as_tensor(init_param).detach().cpu())
        self.l = torch.nn.Parameter(start)

    def forward(self):
        lower = _lower_from_vec6(self.l)
        return lower + torch.tril(lower, -1).t()


def exp_map_so3(omega, epsilon=1.0e-14):
    """Rodrigues formula with the reference's regulariser (utils.py:57-69), on omega's device."""
    zero = omega.new_zeros(())
    hat = torch.stack([torch.stack([zero, -omega[2], omega[1]]),
                       torch.stack([omega[2], zero, -omega[0]]),
                       torch.stack([-omega[1], omega[0], zero])])
    th = torch.norm(omega, p=2)
    return (torch.eye(3, device=omega.device) + (torch.sin(th) / (th + epsilon)) * hat
            + ((1.0 - torch.cos(th)) / ((th + epsilon) * (th + epsilon))) * (hat @ hat))


def _log_map_so3(R, epsilon=1.0e-14):
    """numpy restatement of se3_so3_util.logMapSO3 + getVec3FromSkewSymMat (se3_so3_util.py:73-88,148-164)."""
    c = min(1.0, max(-1.0, (np.trace(R) - 1.0) / 2.0))
    theta = math.acos(c)
    hat = theta * (R - R.T) / (2.0 * math.sin(theta) + epsilon)
    return np.array([0.5 * (hat[2, 1] - hat[1, 2]), 0.5 * (hat[0, 2] - hat[2, 0]), 0.5 * (hat[1, 0] - hat[0, 1])])


class TriangParam3DInertiaMatrixNet(torch.nn.Module):
    """``R diag(J1, J2, J3) R^T`` with ``J3 = sqrt(J1^2 + J2^2 - 2 J1 J2 cos(alpha))``, ``0 < alpha < pi`` -- the principal
    moments satisfy the triangle inequalities by construction (reference: rigid_body_params.py:135-242).

    The reference class cannot be instantiated as shipped (it hands ``init_param=`` to ``UnconstrainedTensor``,
    which has no such argument, rigid_body_params.py:216-221, and feeds numpy arrays to the torch-only
    ``logMapSO3``, :166); this class implements the construction those lines intend."""

    def __init__(self, bias, init_param_std=0.01, init_param=None, is_initializing_params=True):
        self._qdim = 3
        self._bias = bias
        super().__init__()
        j1 = j2 = alpha_logit = None
        if init_param is None or not is_initializing_params:
            axis_angle = torch.empty(3).normal_(mean=0.0, std=init_param_std)
        else:
            mat = torch.as_tensor(init_param).detach().cpu().double().reshape(3, 3).numpy()
            R, J, _ = np.linalg.svd(mat, full_matrices=True)
            if np.linalg.det(R) < 0.0:           # a member of SO(3), not just O(3)
                R[:, 0] = -R[:, 0]
            axis_angle = torch.tensor(_log_map_so3(R), dtype=torch.float32)
            assert J[0] > bias and J[1] > bias, "Please set bias value smaller, such that this condition is satisfied!"
            cos_alpha = (J[0] * J[0] + J[1] * J[1] - J[2] * J[2]) / (2.0 * J[0] * J[1])
            frac = math.acos(min(1.0, max(-1.0, cos_alpha))) / math.pi
            frac = min(1.0 - 1e-6, max(1e-6, frac))
            j1, j2 = torch.tensor(J[0], dtype=torch.float32), torch.tensor(J[1], dtype=torch.float32)
            alpha_logit = torch.tensor([[math.log(frac / (1.0 - frac))]], dtype=torch.float32)
        self.inertia_ori_axis_angle = torch.nn.Parameter(axis_angle)
        self.J1net = PositiveScalar(min_val=bias, init_param_std=0.1, init_param=j1)
        self.J2net = PositiveScalar(min_val=bias, init_param_std=0.1, init_param=j2)
        self.alpha_param_net = UnconstrainedTensor(dim1=1, dim2=1, init_tensor=alpha_logit, init_std=init_param_std)
        self.J = None
        self.R = None
        self.inertia_mat = None

    def forward(self):
        alpha = math.pi * torch.sigmoid(self.alpha_param_net().squeeze())
        J1 = self.J1net().squeeze()
        J2 = self.J2net().squeeze()
        J3 = torch.sqrt(J1 * J1 + J2 * J2 - 2.0 * J1 * J2 * torch.cos(alpha))
        self.J = torch.diag(torch.stack([J1, J2, J3]))
        self.R = exp_map_so3(self.inertia_ori_axis_angle)
        self.inertia_mat = self.R @ (self.J @ self.R.t())
        return self.inertia_mat
