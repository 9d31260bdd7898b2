"""Gradient-based inverse kinematics: optimise joint angles so that the end effector reaches target positions.

Counterpart of the reference's ``examples/run_kinematic_trajectory_opt.py`` (gradients w.r.t. **q**, :44): the backward
pass is the FK adjoint kernel without any table gradient.
"""
import torch

from differentiable_robot_model_b200 import DifferentiableKUKAiiwa


def run(n_iters=200, n_targets=4096, device="cuda"):
    torch.manual_seed(0)
    robot = DifferentiableKUKAiiwa(device=device)
    limits = robot.get_joint_limits()
    lo = torch.tensor([l["lower"] for l in limits], device=device)
    hi = torch.tensor([l["upper"] for l in limits], device=device)
    q_goal = lo + (hi - lo) * torch.rand(n_targets, 7, device=device)
    target, _ = robot.compute_forward_kinematics(q_goal, "iiwa_link_ee")       # reachable targets
    q = (q_goal + 0.3 * torch.randn_like(q_goal)).clamp(lo, hi).requires_grad_(True)
    opt = torch.optim.Adam([q], lr=2e-2)
    history = []
    for i in range(n_iters):
        opt.zero_grad()
        pos, _ = robot.compute_forward_kinematics(q, "iiwa_link_ee")
        loss = (pos - target).square().sum(dim=1).mean()
        loss.backward()
        opt.step()
        history.append(float(loss))
    print(f"mean squared ee error: {history[0]:.5f} -> {history[-1]:.7f}")
    return history


if __name__ == "__main__":
    run()
