"""
Per-link host objects
====================================
Host-side mirror of the reference's ``DifferentiableRigidBody`` (``rigid_body.py:24-171``) and
``DifferentiableSpatialRigidBodyInertia`` (``spatial_vector_algebra.py:308-372``).

In the reference these objects *compute*: every call of ``update_joint_state`` builds ``[B,3,3]``
tensors per link.  Here they only *hold parameters*: the six attributes that can be made learnable
(``trans``, ``rot_angles``, ``joint_damping`` on the body; ``mass``, ``com``, ``inertia_mat`` on
``body.inertia``) are zero-argument callables -- plain lambdas over the URDF constants, or the
``torch.nn.Module`` a caller swapped in through ``make_link_param_learnable``
(``robot_model.py:682-689``).  ``link_table.build_link_table`` evaluates them once per call into the
flat device table the CUDA kernels read; all per-configuration arithmetic happens in the kernels.
"""
from typing import List, Optional

import torch


class DifferentiableSpatialRigidBodyInertia(torch.nn.Module):
    def __init__(self, rigid_body_params, device="cpu"):
        super().__init__()
        self.mass = lambda: rigid_body_params["mass"]
        self.com = lambda: rigid_body_params["com"]
        self.inertia_mat = lambda: rigid_body_params["inertia_mat"]
        self._device = torch.device(device)

    def _get_parameter_values(self):
        return self.mass(), self.com(), self.inertia_mat()

    # The two value-level operations of the reference class (``spatial_vector_algebra.py:321-372``), for callers that
    # use the type directly.  The engine never calls them: the kernels read (I_o, m c, m) from the link table.
    def _origin_inertia(self):
        mass, com, inertia_mat = self._get_parameter_values()
        mass = torch.as_tensor(mass).reshape(())
        c = torch.as_tensor(com).reshape(3)
        zero = torch.zeros((), dtype=c.dtype, device=c.device)
        skew = torch.stack([torch.stack([zero, -c[2], c[1]]), torch.stack([c[2], zero, -c[0]]), torch.stack([-c[1], c[0], zero])])
        return mass, mass * c, torch.as_tensor(inertia_mat).reshape(3, 3) + mass * (skew @ skew.t())

    def multiply_motion_vec(self, smv):
        """``I v`` for a spatial motion vector (``spatial_vector_algebra.py:321-338``): ``lin = m v.lin - (m c) x v.ang``,
        ``ang = I_o v.ang + (m c) x v.lin`` with ``I_o = I_c + m S(c) S(c)^T`` (``inertia_mat`` used as given)."""
        from .spatial_vector_algebra import SpatialForceVec
        mass, mcom, inertia = self._origin_inertia()
        mc = mcom.expand_as(smv.ang)
        lin = mass * smv.lin - torch.linalg.cross(mc, smv.ang)
        ang = smv.ang @ inertia.t() + torch.linalg.cross(mc, smv.lin)
        return SpatialForceVec(lin, ang)

    def get_spatial_mat(self):
        """6x6 spatial inertia in [ang; lin] order, ``[[I_o, (m c)^], [((m c)^)^T, m 1]]`` (``:340-372``)."""
        mass, mcom, inertia = self._origin_inertia()
        zero = torch.zeros((), dtype=mcom.dtype, device=mcom.device)
        skew = torch.stack([torch.stack([zero, -mcom[2], mcom[1]]), torch.stack([mcom[2], zero, -mcom[0]]),
                            torch.stack([-mcom[1], mcom[0], zero])])
        top = torch.cat([inertia, skew], dim=1)
        bot = torch.cat([skew.t(), mass * torch.eye(3, dtype=mcom.dtype, device=mcom.device)], dim=1)
        return torch.cat([top, bot], dim=0)


class DifferentiableRigidBody(torch.nn.Module):
    """One link plus the joint that connects it to its parent (joint at the start of the link)."""

    _children: List["DifferentiableRigidBody"]

    def __init__(self, rigid_body_params, device="cpu"):
        super().__init__()
        # plain (unregistered) references: registering the parent as a sub-module would create
        # module cycles and duplicate state_dict entries
        object.__setattr__(self, "_parent", None)
        object.__setattr__(self, "_children", [])
        self._device = torch.device(device)
        self.joint_id = rigid_body_params["joint_id"]
        self.name = rigid_body_params["link_name"]
        self.joint_type = rigid_body_params["joint_type"]
        self.joint_idx = None          # DoF column, assigned by the model for movable joints

        # parameters that can be made learnable
        self.inertia = DifferentiableSpatialRigidBodyInertia(rigid_body_params, device=self._device)
        self.joint_damping = lambda: rigid_body_params["joint_damping"]
        self.trans = lambda: rigid_body_params["trans"].reshape(1, 3)
        self.rot_angles = lambda: rigid_body_params["rot_angles"].reshape(1, 3)

        self.joint_axis = rigid_body_params["joint_axis"]
        self.joint_limits = rigid_body_params["joint_limits"]

        # The reference builds the joint pose of a *fixed* joint once, in the constructor
        # (rigid_body.py:64-67); making its trans / rot_angles learnable later has no effect.
        # Keep the construction-time values so the link table reproduces that behaviour.
        self._ctor_trans = rigid_body_params["trans"].reshape(1, 3).detach().clone()
        self._ctor_rot_angles = rigid_body_params["rot_angles"].reshape(1, 3).detach().clone()

    # Per-joint value helpers with the reference's names and semantics (rigid_body.py:130-165): joint pose
    # R_fix(rpy) R_axis(sign q) / trans, joint velocity and acceleration about the signed axis.  The model's compute_*
    # entry points never call them (the kernels do this per link in registers); they exist for callers that walk the
    # bodies themselves.
    def update_joint_state(self, q, qd):
        from .spatial_vector_algebra import CoordinateTransform, SpatialMotionVec, x_rot, y_rot, z_rot
        batch = q.shape[0]
        axis = self.joint_axis.reshape(1, 3).to(q.device)
        ang_vel = qd.reshape(batch, 1) @ axis
        self.joint_vel = SpatialMotionVec(torch.zeros_like(ang_vel), ang_vel)
        rpy = self.rot_angles().reshape(3).to(q.device)
        fixed = (z_rot(rpy[2]) @ y_rot(rpy[1])) @ x_rot(rpy[0])
        if torch.abs(axis[0, 0]) == 1:
            rot = x_rot(torch.sign(axis[0, 0]) * q)
        elif torch.abs(axis[0, 1]) == 1:
            rot = y_rot(torch.sign(axis[0, 1]) * q)
        else:
            rot = z_rot(torch.sign(axis[0, 2]) * q)
        self.joint_pose = CoordinateTransform(rot=fixed.expand(batch, 3, 3) @ rot,
                                              trans=self.trans().reshape(1, 3).to(q.device).expand(batch, 3), device=q.device)

    def update_joint_acc(self, qdd):
        from .spatial_vector_algebra import SpatialMotionVec
        ang_acc = qdd.reshape(qdd.shape[0], 1) @ self.joint_axis.reshape(1, 3).to(qdd.device)
        self.joint_acc = SpatialMotionVec(torch.zeros_like(ang_acc), ang_acc)

    # kinematic tree construction (robot_model.py:133-137)
    def set_parent(self, link: "DifferentiableRigidBody"):
        object.__setattr__(self, "_parent", link)

    def add_child(self, link: "DifferentiableRigidBody"):
        self._children.append(link)

    # per-link kinematic state left behind by model.update_kinematic_state (reference: rigid_body.py:70-73)
    def _bind_model(self, model):
        import weakref
        object.__setattr__(self, "_model_ref", weakref.ref(model))

    @property
    def pose(self):
        """World pose of this link (``CoordinateTransform``) for the last ``update_kinematic_state`` call."""
        return self._model_ref()._body_pose(self.joint_id)

    @property
    def vel(self):
        """Body-frame spatial velocity (``SpatialMotionVec``) for the last ``update_kinematic_state`` call."""
        return self._model_ref()._body_vel(self.joint_id)

    @property
    def acc(self):
        """Body-frame spatial acceleration (``SpatialMotionVec``; base acceleration = gravity folded in like the
        reference, ``robot_model.py:262-277``) for the inputs of the last ``compute_inverse_dynamics`` call."""
        return self._model_ref()._body_acc(self.joint_id)

    @property
    def force(self):
        """Wrench of this link plus everything it carries (``SpatialForceVec``, ``robot_model.py:284-301``) for the
        inputs of the last ``compute_inverse_dynamics`` call."""
        return self._model_ref()._body_force(self.joint_id)

    def get_joint_limits(self):
        return self.joint_limits

    def get_joint_damping_const(self):
        return self.joint_damping()
