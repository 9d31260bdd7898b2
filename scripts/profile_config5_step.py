import sys, os, json, torch
sys.path.insert(0, "."); sys.path.insert(0, "scripts")
import bench_sharded as bs
import differentiable_robot_model_b200 as drm
from differentiable_robot_model_b200.rigid_body_params import UnconstrainedScalar, UnconstrainedTensor
dev = torch.device("cuda", 0)
m = drm.DifferentiableKUKAiiwa(device=dev)
for i in range(1, 8):
    b = m._bodies[i]
    m.make_link_param_learnable(b.name, "mass", UnconstrainedScalar(init_val=b.inertia.mass().detach().clone()))
    m.make_link_param_learnable(b.name, "com", UnconstrainedTensor(1, 3, init_tensor=b.inertia.com().detach().clone()))
    m.make_link_param_learnable(b.name, "inertia_mat", UnconstrainedTensor(3, 3, init_tensor=b.inertia.inertia_mat().detach().clone().reshape(3, 3)))
flat = m.fuse_learnable_parameters()
q, qd, qdd = bs.sample(m, 131072, 0, dev)
target = torch.randn(131072, 7, device=dev)
from differentiable_robot_model_b200.parallel import PeerAllReduceAdam
opt = PeerAllReduceAdam(flat, lr=1e-3)
flat.grad = torch.zeros_like(flat)
def step():
    flat.grad.zero_()
    with m.shared_link_table():
        pos, quat, jl, ja = m.compute_fk_and_jacobian(q, "iiwa_link_ee")
        tau = m.compute_inverse_dynamics(q, qd, qdd)
    loss = (tau - target).square().mean() + pos.square().mean()
    loss.backward()
    opt.step()
for _ in range(5): step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(5): step()
    torch.cuda.synchronize()
rows = [(k.key[:90], k.count // 5, round(k.device_time_total / 5, 2)) for k in prof.key_averages() if k.device_time_total > 0 and k.cpu_time_total == 0]
rows.sort(key=lambda r: -r[2])
for r in rows: print(r)
print("total kernel us per step", round(sum(r[2] for r in rows), 1), "kernels per step", sum(r[1] for r in rows))
