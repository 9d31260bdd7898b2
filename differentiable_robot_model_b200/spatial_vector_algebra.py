"""
Spatial vector algebra value types
====================================
Host-side mirror of the value types of the reference's ``spatial_vector_algebra.py``
(``CoordinateTransform`` :56-172, ``SpatialMotionVec`` :175-250, ``SpatialForceVec`` :253-305, ``x_rot`` /
``y_rot`` / ``z_rot`` :14-53): same constructors, method names, argument meaning and ``[ang, lin]``
(Featherstone) vector order, so code written against the reference's types keeps working.

Inside the engine these objects do not exist: the CUDA kernels fuse every one of these operations
for a whole kinematic tree into registers (``csrc/``).  The classes below are small batched-torch
conveniences for callers that want to manipulate individual transforms / spatial vectors; they work on
any device and are differentiable.  Unlike the reference, ``get_quaternion`` is vectorised over the
batch (the reference loops over batch elements in Python, ``spatial_vector_algebra.py:116-135``) while
keeping its branch structure and xyzw sign convention.
"""
from __future__ import annotations

from typing import Optional

import torch


def _angle_1d(angle: torch.Tensor) -> torch.Tensor:
    return angle.reshape(-1)


def _stack_rows(rows, batch):
    return torch.stack(rows, dim=-1).reshape(batch, 3, 3)


def x_rot(angle: torch.Tensor) -> torch.Tensor:
    a = _angle_1d(angle)
    c, s, one, zero = torch.cos(a), torch.sin(a), torch.ones_like(a), torch.zeros_like(a)
    return _stack_rows([one, zero, zero, zero, c, -s, zero, s, c], a.shape[0])


def y_rot(angle: torch.Tensor) -> torch.Tensor:
    a = _angle_1d(angle)
    c, s, one, zero = torch.cos(a), torch.sin(a), torch.ones_like(a), torch.zeros_like(a)
    return _stack_rows([c, zero, s, zero, one, zero, -s, zero, c], a.shape[0])


def z_rot(angle: torch.Tensor) -> torch.Tensor:
    a = _angle_1d(angle)
    c, s, one, zero = torch.cos(a), torch.sin(a), torch.ones_like(a), torch.zeros_like(a)
    return _stack_rows([c, -s, zero, s, c, zero, zero, zero, one], a.shape[0])


def _skew(v: torch.Tensor) -> torch.Tensor:
    z = torch.zeros_like(v[:, 0])
    return torch.stack([z, -v[:, 2], v[:, 1], v[:, 2], z, -v[:, 0], -v[:, 1], v[:, 0], z], dim=-1).reshape(-1, 3, 3)


def _apply(mat: torch.Tensor, vec: torch.Tensor) -> torch.Tensor:
    return (mat @ vec.unsqueeze(2)).squeeze(2)


class CoordinateTransform(object):
    """SE(3) element as ``(rot [B,3,3], trans [B,3])``."""

    def __init__(self, rot=None, trans=None, device="cpu"):
        self._device = torch.device(device)
        self._rot = torch.eye(3, device=self._device) if rot is None else rot
        self._trans = torch.zeros(3, device=self._device) if trans is None else trans
        if self._rot.ndim == 2:
            self._rot = self._rot.unsqueeze(0)
        if self._trans.ndim == 1:
            self._trans = self._trans.unsqueeze(0)

    def set_translation(self, t):
        self._trans = t.unsqueeze(0) if t.ndim == 1 else t

    def set_rotation(self, rot):
        self._rot = rot.unsqueeze(0) if rot.ndim == 2 else rot

    def rotation(self):
        return self._rot

    def translation(self):
        return self._trans

    def inverse(self):
        rt = self._rot.transpose(-2, -1)
        return CoordinateTransform(rt, -_apply(rt, self._trans))

    def multiply_transform(self, other: "CoordinateTransform"):
        return CoordinateTransform(self._rot @ other.rotation(), _apply(self._rot, other.translation()) + self._trans)

    def trans_cross_rot(self):
        return _skew(self._trans) @ self._rot

    def get_quaternion(self):
        R = self._rot
        d0, d1, d2 = R[:, 0, 0], R[:, 1, 1], R[:, 2, 2]
        t_a = d0 + d1 + d2 + 1
        case_a = t_a > 1
        i2 = (~case_a) & (d2 > torch.maximum(d0, d1))
        i1 = (~case_a) & (~i2) & (d1 > d0)
        t0, t1, t2 = d0 - (d1 + d2) + 1, d1 - (d2 + d0) + 1, d2 - (d0 + d1) + 1
        qa = torch.stack([R[:, 2, 1] - R[:, 1, 2], R[:, 0, 2] - R[:, 2, 0], R[:, 1, 0] - R[:, 0, 1], t_a], dim=1)
        q0 = torch.stack([t0, R[:, 0, 1] + R[:, 1, 0], R[:, 2, 0] + R[:, 0, 2], R[:, 2, 1] - R[:, 1, 2]], dim=1)
        q1 = torch.stack([R[:, 0, 1] + R[:, 1, 0], t1, R[:, 1, 2] + R[:, 2, 1], R[:, 0, 2] - R[:, 2, 0]], dim=1)
        q2 = torch.stack([R[:, 2, 0] + R[:, 0, 2], R[:, 1, 2] + R[:, 2, 1], t2, R[:, 1, 0] - R[:, 0, 1]], dim=1)
        t = torch.where(case_a, t_a, torch.where(i2, t2, torch.where(i1, t1, t0)))
        q = torch.where(case_a[:, None], qa, torch.where(i2[:, None], q2, torch.where(i1[:, None], q1, q0)))
        return q * (0.5 / torch.sqrt(t))[:, None]

    def _plucker(self):
        B = self._rot.shape[0]
        rt = self._rot.transpose(-2, -1)
        mat = torch.zeros((B, 6, 6), device=self._rot.device, dtype=self._rot.dtype)
        mat[:, :3, :3] = rt
        mat[:, 3:, 3:] = rt
        mat[:, 3:, :3] = -(rt @ _skew(self._trans.expand(B, 3)))
        return mat

    def to_matrix(self):
        """6x6 Pluecker motion transform ``[[E, 0], [-E r^x, E]]`` with ``E = rot^T`` (reference :138-154)."""
        return self._plucker()

    def to_matrix_transpose(self):
        """The variant the reference calls ``to_matrix_transpose`` (:156-172): lower-left block
        ``-(rot r^x)^T``."""
        B = self._rot.shape[0]
        rt = self._rot.transpose(-2, -1)
        mat = torch.zeros((B, 6, 6), device=self._rot.device, dtype=self._rot.dtype)
        mat[:, :3, :3] = rt
        mat[:, 3:, 3:] = rt
        mat[:, 3:, :3] = -(self._rot @ _skew(self._trans.expand(B, 3))).transpose(-1, -2)
        return mat


class _SpatialVec(object):
    def __init__(self, lin=None, ang=None, device=None):
        if lin is None or ang is None:
            assert device is not None, "Cannot initialize with default values without specifying device."
            device = torch.device(device)
        self.lin = lin if lin is not None else torch.zeros((1, 3), device=device)
        self.ang = ang if ang is not None else torch.zeros((1, 3), device=device)

    def get_vector(self):
        return torch.cat([self.ang, self.lin], dim=1)

    def multiply(self, v):
        b = self.lin.shape[0]
        return SpatialForceVec(self.lin * v.view(b, 1), self.ang * v.view(b, 1))

    def dot(self, other):
        return torch.sum(self.ang * other.ang, dim=-1) + torch.sum(self.lin * other.lin, dim=-1)


class SpatialMotionVec(_SpatialVec):
    def __init__(self, lin_motion: Optional[torch.Tensor] = None, ang_motion: Optional[torch.Tensor] = None, device=None):
        super().__init__(lin_motion, ang_motion, device)

    def add_motion_vec(self, smv: "SpatialMotionVec") -> "SpatialMotionVec":
        return SpatialMotionVec(self.lin + smv.lin, self.ang + smv.ang)

    def cross_motion_vec(self, smv: "SpatialMotionVec") -> "SpatialMotionVec":
        ang = torch.cross(self.ang, smv.ang, dim=-1)
        lin = torch.cross(self.ang, smv.lin, dim=-1) + torch.cross(self.lin, smv.ang, dim=-1)
        return SpatialMotionVec(lin, ang)

    def cross_force_vec(self, sfv: "SpatialForceVec") -> "SpatialForceVec":
        ang = torch.cross(self.ang, sfv.ang, dim=-1) + torch.cross(self.lin, sfv.lin, dim=-1)
        lin = torch.cross(self.ang, sfv.lin, dim=-1)
        return SpatialForceVec(lin, ang)

    def transform(self, transform: CoordinateTransform) -> "SpatialMotionVec":
        ang = _apply(transform.rotation(), self.ang)
        lin = _apply(transform.trans_cross_rot(), self.ang) + _apply(transform.rotation(), self.lin)
        return SpatialMotionVec(lin, ang)


class SpatialForceVec(_SpatialVec):
    def __init__(self, lin_force: Optional[torch.Tensor] = None, ang_force: Optional[torch.Tensor] = None, device=None):
        super().__init__(lin_force, ang_force, device)

    def add_force_vec(self, sfv: "SpatialForceVec") -> "SpatialForceVec":
        return SpatialForceVec(self.lin + sfv.lin, self.ang + sfv.ang)

    def transform(self, transform: CoordinateTransform) -> "SpatialForceVec":
        lin = _apply(transform.rotation(), self.lin)
        ang = _apply(transform.trans_cross_rot(), self.lin) + _apply(transform.rotation(), self.ang)
        return SpatialForceVec(lin, ang)


# the reference defines the inertia type in this module (spatial_vector_algebra.py:308); it lives next to the body here
from .rigid_body import DifferentiableSpatialRigidBodyInertia  # noqa: E402,F401
