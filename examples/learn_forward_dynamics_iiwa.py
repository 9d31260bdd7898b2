"""Learn link-1 mass / centre of mass / inertia of a Kuka iiwa from joint accelerations (B200 engine).

Same experiment, entry point and signature as the reference's ``examples/learn_forward_dynamics_iiwa.py:50`` --
``run(n_epochs, n_data, device)``: every ``compute_forward_dynamics`` call is ONE articulated-body kernel launch
and ``loss.backward()`` runs its analytic adjoint kernel (the learned ``inertia_mat`` is an unconstrained, hence
non-symmetric, 3x3 -- the kernels carry general 6x6 articulated inertias for exactly this case).
"""
import numpy as np
import torch
from torch.utils.data import DataLoader

from differentiable_robot_model_b200 import DifferentiableKUKAiiwa, DifferentiableRobotModel
from differentiable_robot_model_b200.data_utils import generate_sine_motion_forward_dynamics_data
from differentiable_robot_model_b200.rigid_body_params import PositiveScalar, UnconstrainedTensor


class NMSELoss(torch.nn.Module):
    def __init__(self, var):
        super().__init__()
        self.var = var

    def forward(self, yp, yt):
        return (((yp - yt) ** 2) / self.var).mean()


def run(n_epochs=100, n_data=10000, device="cuda"):
    gt_robot_model = DifferentiableKUKAiiwa(device=device)
    learnable_robot_model = DifferentiableRobotModel(gt_robot_model.urdf_path, name="kuka_iiwa", device=device)
    learnable_robot_model.make_link_param_learnable("iiwa_link_1", "mass", PositiveScalar())
    learnable_robot_model.make_link_param_learnable("iiwa_link_1", "com", UnconstrainedTensor(dim1=1, dim2=3))
    learnable_robot_model.make_link_param_learnable("iiwa_link_1", "inertia_mat", UnconstrainedTensor(dim1=3, dim2=3))

    train_data = generate_sine_motion_forward_dynamics_data(gt_robot_model, n_data=n_data, dt=1.0 / 250.0, freq=0.1)
    train_loader = DataLoader(dataset=train_data, batch_size=100, shuffle=False)
    optimizer = torch.optim.Adam(learnable_robot_model.parameters(), lr=1e-2)
    loss_fn = NMSELoss(train_data.var())
    history = []
    for epoch in range(n_epochs):
        losses = []
        for q, qd, qdd, tau in train_loader:
            optimizer.zero_grad()
            qdd_pred = learnable_robot_model.compute_forward_dynamics(q=q, qd=qd, f=tau, include_gravity=True,
                                                                      use_damping=True)
            loss = loss_fn(qdd_pred, qdd)
            loss.backward()
            optimizer.step()
            losses.append(loss.item())
        history.append(float(np.mean(losses)))
        print(f"i: {epoch} loss: {history[-1]}")
    learnable_robot_model.print_learnable_params()
    return history


if __name__ == "__main__":
    run()
