"""
Synthetic datasets for the parameter-learning examples
========================================================
Mirror of the reference's ``data_utils.py`` (``InverseDynamicsDataset`` :13, ``ForwardDynamicsDataset`` :31,
``generate_random_forward_kinematics_data`` :49, ``generate_random_inverse_dynamics_data`` :70,
``generate_sine_motion_inverse_dynamics_data`` :112, ``generate_sine_motion_forward_dynamics_data`` :148): same names,
arguments and returned containers, so the reference's examples run against this package unchanged.  The
generators are thin callers of the hot path (one FK or RNEA launch per dataset) and build the trajectories directly
on the model's device.  Two reference limitations are not reproduced: the random inverse-dynamics generator hard-codes
7 DoF (``data_utils.py:81,87,95``; here ``robot_model._n_dofs``), everything else -- including the sine generator's
``T = int(n_data * dt)`` quirk that yields an all-zero trajectory when ``n_data * dt < 2`` (``:118``) -- is kept.
"""
import math

import numpy as np
import torch
from torch.utils.data import Dataset


class _DynamicsDataset(Dataset):
    def __init__(self, data):
        self.data = data

    def __getitem__(self, index):
        d = self.data
        return [d["q"][index], d["qd"][index], d["qdd_des"][index], d["tau"][index]]

    def __len__(self):
        return self.data["q"].shape[0]


class InverseDynamicsDataset(_DynamicsDataset):
    """Items ``[q, qd, qdd_des, tau]``; ``var()`` is the per-joint variance of the torques."""

    def var(self):
        return self.data["tau"].var(dim=0)


class ForwardDynamicsDataset(_DynamicsDataset):
    """Items ``[q, qd, qdd, tau]``; ``var()`` is the per-joint variance of the accelerations."""

    def var(self):
        return self.data["qdd_des"].var(dim=0)


def _limits(robot_model):
    lim = robot_model.get_joint_limits()
    lower = np.asarray([j["lower"] for j in lim])
    upper = np.asarray([j["upper"] for j in lim])
    vel = np.asarray([j["velocity"] for j in lim])
    return lower, upper, vel


def _uniform(low, high, n_data, device):
    return torch.tensor(np.random.uniform(low=low, high=high, size=(n_data, len(low))), dtype=torch.float32, device=device)


def generate_random_forward_kinematics_data(robot_model, n_data, ee_name):
    lower, upper, _ = _limits(robot_model)
    q = _uniform(lower, upper, n_data, robot_model._device)
    ee_pos, _ = robot_model.compute_forward_kinematics(q=q, link_name=ee_name)
    return {"q": q, "ee_pos": ee_pos}


def generate_random_inverse_dynamics_data(robot_model, n_data):
    device = robot_model._device
    lower, upper, vel = _limits(robot_model)
    vel = 0.2 * vel
    q = _uniform(lower, upper, n_data, device)
    qd = _uniform(-vel, vel, n_data, device)
    qdd_des = _uniform(-2.0 * vel, 2.0 * vel, n_data, device)
    tau = robot_model.compute_inverse_dynamics(q=q, qd=qd, qdd_des=qdd_des, include_gravity=True)
    return InverseDynamicsDataset(data={"q": q, "qd": qd, "qdd_des": qdd_des, "tau": tau})


def _sine_motion(robot_model, n_data, dt, freq):
    device, n_dofs = robot_model._device, robot_model._n_dofs
    amplitude = 0.7
    horizon = int(n_data * dt)
    t = torch.linspace(0.0, horizon - 1, n_data, device=device)
    w = 2.0 * math.pi * freq
    phase = (w * t).reshape(n_data, 1).expand(n_data, n_dofs)
    q = (amplitude * torch.sin(phase)).contiguous()
    qd = (w * amplitude * torch.cos(phase)).contiguous()
    qdd_des = (-(w ** 2) * amplitude * torch.sin(phase)).contiguous()
    tau = robot_model.compute_inverse_dynamics(q=q, qd=qd, qdd_des=qdd_des, include_gravity=True)
    return {"q": q, "qd": qd, "qdd_des": qdd_des, "tau": tau}


def generate_sine_motion_inverse_dynamics_data(robot_model, n_data, dt, freq):
    return InverseDynamicsDataset(data=_sine_motion(robot_model, n_data, dt, freq))


def generate_sine_motion_forward_dynamics_data(robot_model, n_data, dt, freq):
    return ForwardDynamicsDataset(data=_sine_motion(robot_model, n_data, dt, freq))
