"""Learn a joint origin (trans, rot_angles of link 1) of a Kuka iiwa from end-effector positions (B200 engine).

Same experiment and ``run(n_epochs, n_data, device)`` entry point as the reference's
``examples/learn_kinematics_of_iiwa.py:26``: one FK kernel launch forward, the analytic FK adjoint kernel backward.
"""
import torch

from differentiable_robot_model_b200 import DifferentiableKUKAiiwa, DifferentiableRobotModel
from differentiable_robot_model_b200.data_utils import generate_random_forward_kinematics_data
from differentiable_robot_model_b200.rigid_body_params import UnconstrainedTensor


def run(n_epochs=3000, n_data=100, device="cuda"):
    gt_robot_model = DifferentiableKUKAiiwa(device=device)
    learnable_robot_model = DifferentiableRobotModel(gt_robot_model.urdf_path, "kuka_iiwa", device=device)
    learnable_robot_model.make_link_param_learnable("iiwa_link_1", "trans", UnconstrainedTensor(dim1=1, dim2=3))
    learnable_robot_model.make_link_param_learnable("iiwa_link_1", "rot_angles", UnconstrainedTensor(dim1=1, dim2=3))

    train_data = generate_random_forward_kinematics_data(gt_robot_model, n_data=n_data, ee_name="iiwa_link_ee")
    q, gt_ee_pos = train_data["q"], train_data["ee_pos"]
    optimizer = torch.optim.Adam(learnable_robot_model.parameters(), lr=1e-3)
    loss_fn = torch.nn.MSELoss()
    history = []
    for i in range(n_epochs):
        optimizer.zero_grad()
        ee_pos_pred, _ = learnable_robot_model.compute_forward_kinematics(q=q, link_name="iiwa_link_ee")
        loss = loss_fn(ee_pos_pred, gt_ee_pos)
        loss.backward()
        optimizer.step()
        history.append(float(loss))
        if i % 100 == 0:
            print(f"i: {i}, loss: {history[-1]}")
    print("gt trans:", gt_robot_model._bodies[1].trans())
    learnable_robot_model.print_learnable_params()
    return history


if __name__ == "__main__":
    run()
