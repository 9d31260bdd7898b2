"""Learn link-1 mass / inertia and link-2 joint origin of a Kuka iiwa from torques (B200 engine).

Same experiment, entry point and signature as the reference's ``examples/learn_dynamics_iiwa.py:49`` --
``run(n_epochs, n_data, device)`` -- written against this package: every ``compute_inverse_dynamics`` call is one
RNEA kernel launch and ``loss.backward()`` runs the analytic adjoint kernel.
"""
import numpy as np
import torch
from torch.utils.data import DataLoader

from differentiable_robot_model_b200 import DifferentiableKUKAiiwa, DifferentiableRobotModel
from differentiable_robot_model_b200.data_utils import generate_sine_motion_inverse_dynamics_data
from differentiable_robot_model_b200.rigid_body_params import PositiveScalar, UnconstrainedTensor


class NMSELoss(torch.nn.Module):
    def __init__(self, var):
        super().__init__()
        self.var = var

    def forward(self, yp, yt):
        return (((yp - yt) ** 2) / self.var).mean()


def run(n_epochs=10, n_data=1000, device="cuda"):
    gt_robot_model = DifferentiableKUKAiiwa(device=device)
    learnable_robot_model = DifferentiableRobotModel(gt_robot_model.urdf_path, name="kuka_iiwa", device=device)
    learnable_robot_model.make_link_param_learnable("iiwa_link_1", "mass", PositiveScalar())
    learnable_robot_model.make_link_param_learnable("iiwa_link_1", "inertia_mat", UnconstrainedTensor(dim1=3, dim2=3))
    learnable_robot_model.make_link_param_learnable("iiwa_link_2", "trans", UnconstrainedTensor(dim1=1, dim2=3))

    train_data = generate_sine_motion_inverse_dynamics_data(gt_robot_model, n_data=n_data, dt=1.0 / 250.0, freq=0.05)
    train_loader = DataLoader(dataset=train_data, batch_size=100, shuffle=False)
    optimizer = torch.optim.Adam(learnable_robot_model.parameters(), lr=1e-2)
    loss_fn = NMSELoss(train_data.var())
    history = []
    for epoch in range(n_epochs):
        losses = []
        for q, qd, qdd_des, gt_tau in train_loader:
            optimizer.zero_grad()
            tau_pred = learnable_robot_model.compute_inverse_dynamics(q=q, qd=qd, qdd_des=qdd_des, include_gravity=True)
            loss = loss_fn(tau_pred, gt_tau)
            loss.backward()
            optimizer.step()
            losses.append(loss.item())
        history.append(float(np.mean(losses)))
        print(f"i: {epoch} loss: {history[-1]}")
    learnable_robot_model.print_learnable_params()
    return history


if __name__ == "__main__":
    run()
