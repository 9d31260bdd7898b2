#!/usr/bin/env python
"""Where does a parameter-learning step (BASELINE config 5, one 131 072-row shard) spend its time?
torch.profiler table of CUDA kernels + a coarse CUDA-event breakdown of the phases."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import differentiable_robot_model_b200 as drm  # noqa: E402
from differentiable_robot_model_b200 import engine  # noqa: E402
from differentiable_robot_model_b200.rigid_body_params import UnconstrainedScalar, UnconstrainedTensor  # noqa: E402
from oracle import drm_oracle as O  # noqa: E402

DEV = torch.device("cuda", 0)
B = int(os.environ.get("BATCH", 131072))


def learnable_kuka():
    m = drm.DifferentiableKUKAiiwa(device=DEV)
    for i in range(1, 8):
        b = m._bodies[i]
        m.make_link_param_learnable(b.name, "mass", UnconstrainedScalar(init_val=b.inertia.mass().detach().clone()))
        m.make_link_param_learnable(b.name, "com", UnconstrainedTensor(1, 3, init_tensor=b.inertia.com().detach().clone()))
        m.make_link_param_learnable(b.name, "inertia_mat", UnconstrainedTensor(
            3, 3, init_tensor=b.inertia.inertia_mat().detach().clone().reshape(3, 3)))
    return m


def ev():
    return torch.cuda.Event(enable_timing=True)


def main():
    m = learnable_kuka()
    robot = O.load_robot(m.urdf_path, torch.float32)
    q, qd, qdd = (t.to(DEV) for t in O.sample_inputs(robot, B, seed=0))
    target = torch.randn(B, 7, device=DEV)

    def step():
        for p in m.parameters():
            p.grad = None
        tau = m.compute_inverse_dynamics(q, qd, qdd)
        loss = (tau - target).square().mean()
        loss.backward()

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    import time
    t0 = time.perf_counter()
    for _ in range(20):
        step()
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - t0) / 20 * 1e3

    # phase breakdown with events (forward only pieces)
    out = {"batch": B, "wall_ms_per_step": wall_ms}
    e = [ev() for _ in range(6)]
    e[0].record()
    table = m._link_table()
    e[1].record()
    tau = engine.InverseDynamicsFunction.apply(table, q, qd, qdd, m._topology, 3)
    e[2].record()
    loss = (tau - target).square().mean()
    e[3].record()
    loss.backward()
    e[4].record()
    torch.cuda.synchronize()
    out["gpu_ms"] = {"table_build": e[0].elapsed_time(e[1]), "rnea_forward": e[1].elapsed_time(e[2]),
                     "loss": e[2].elapsed_time(e[3]), "backward_total": e[3].elapsed_time(e[4])}
    # raw kernels alone
    g = torch.randn(B, 7, device=DEV)
    tg = torch.zeros_like(table.detach())
    ws = engine._workspace(m._topology, B, DEV)
    import ctypes
    lib = engine.lib()

    def raw_bwd(need_inputs):
        qg = torch.empty_like(q) if need_inputs else None
        rc = lib.drmb200_inverse_dynamics_backward(
            ctypes.byref(m._topology), engine._ptr(table.detach()), engine._ptr(q), engine._ptr(qd), engine._ptr(qdd), B, 3,
            engine._ptr(g), engine._ptr(qg), engine._ptr(qg), engine._ptr(qg), engine._ptr(tg), engine._ptr(ws),
            engine._stream())
        assert rc == 0

    for need in (False, True):
        for _ in range(3):
            raw_bwd(need)
        a, b = ev(), ev()
        a.record()
        for _ in range(10):
            raw_bwd(need)
        b.record()
        torch.cuda.synchronize()
        out["gpu_ms"]["rnea_backward_kernel" + ("_with_input_grads" if need else "_table_only")] = a.elapsed_time(b) / 10

    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            step()
        torch.cuda.synchronize()
    rows = []
    for k in sorted(prof.key_averages(), key=lambda k: -k.device_time_total)[:12]:
        rows.append({"name": k.key[:70], "calls": k.count, "cuda_us_total": k.device_time_total, "cpu_us_total": k.cpu_time_total})
    out["top_cuda"] = rows
    out["n_cuda_kernel_launches_per_step"] = sum(k.count for k in prof.key_averages() if k.device_time_total > 0 and k.cpu_time_total == 0) / 3
    print(json.dumps(out))


if __name__ == "__main__":
    main()
