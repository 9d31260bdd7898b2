"""Learn the two joint origins of the planar 2-link toy robot from end-effector positions, freezing and
un-freezing one of them on the way (B200 engine).

Same experiment as the reference's ``examples/learn_kinematics_of_toy.py:27`` -- ``run(n_epochs, n_data, device)``.
Every ``compute_forward_kinematics`` call is one kernel launch; ``loss.backward()`` runs the analytic adjoint.
"""
import torch

from differentiable_robot_model_b200 import DifferentiableRobotModel, DifferentiableTwoLinkRobot
from differentiable_robot_model_b200.data_utils import generate_random_forward_kinematics_data
from differentiable_robot_model_b200.rigid_body_params import UnconstrainedTensor


def run(n_epochs=3000, n_data=100, device="cuda"):
    gt_robot_model = DifferentiableTwoLinkRobot(device=device)
    learnable_robot_model = DifferentiableRobotModel(gt_robot_model.urdf_path, name="2link", device=device)
    learnable_robot_model.make_link_param_learnable("arm1", "trans", UnconstrainedTensor(dim1=1, dim2=3))
    learnable_robot_model.make_link_param_learnable("arm2", "trans", UnconstrainedTensor(dim1=1, dim2=3))

    train_data = generate_random_forward_kinematics_data(gt_robot_model, n_data=n_data, ee_name="endEffector")
    q, gt_ee_pos = train_data["q"], train_data["ee_pos"]

    optimizer = torch.optim.Adam(learnable_robot_model.parameters(), lr=1e-3)
    loss_fn = torch.nn.MSELoss()
    history = []
    for i in range(n_epochs):
        optimizer.zero_grad()
        ee_pos_pred, _ = learnable_robot_model.compute_forward_kinematics(q=q, link_name="endEffector")
        loss = loss_fn(ee_pos_pred, gt_ee_pos)
        history.append(float(loss.detach()))
        if i % 500 == 0:
            print(f"i: {i}, loss: {history[-1]}")
        if i == 10:
            learnable_robot_model.freeze_learnable_link_param(link_name="arm1", parameter_name="trans")
        if i == 100:
            learnable_robot_model.unfreeze_learnable_link_param(link_name="arm1", parameter_name="trans")
        loss.backward()
        optimizer.step()
    print("ground-truth joint origins:", gt_robot_model._bodies[1].trans(), gt_robot_model._bodies[2].trans())
    learnable_robot_model.print_learnable_params()
    return history


if __name__ == "__main__":
    run()
