#!/usr/bin/env python
"""The two SHARDED BASELINE configs on this rank's shard (called by bench.py on every rank; also runnable alone on 1 GPU):

  config 4  Allegro hand FK + Jacobian of the four fingertips, 262 144 configurations over 8 GPUs = 32 768 per GPU:
            one fused tree-walk launch per step (and, beside it, four single-fingertip launches)
  config 5  Kuka iiwa FK+Jacobian + RNEA forward, scalar loss, backward to the inertial parameters of links 1..7
            (21 tensors fused into one flat Parameter), SUM all-reduce of that gradient over the ranks (the only
            collective of the data path), fused Adam step; 1 048 576 over 8 GPUs = 131 072 per GPU

Timing: CUDA events on the launching stream, `reps` repetitions of a `steps`-step region, median per rank; the caller
takes the max over ranks.  Weak scaling: the per-GPU shard is fixed."""
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import differentiable_robot_model_b200 as drm  # noqa: E402
from differentiable_robot_model_b200 import engine  # noqa: E402
from differentiable_robot_model_b200.rigid_body_params import UnconstrainedScalar, UnconstrainedTensor  # noqa: E402

ALLEGRO = "allegro/urdf/allegro_hand_description_left.urdf"
TIPS = ["link_3.0_tip", "link_7.0_tip", "link_11.0_tip", "link_15.0_tip"]


def sample(model, batch, seed, dev):
    gen = torch.Generator().manual_seed(seed)
    lim = model.get_joint_limits()
    lo = torch.tensor([float(l["lower"]) for l in lim]); hi = torch.tensor([float(l["upper"]) for l in lim])
    vel = torch.tensor([float(l["velocity"]) for l in lim])
    u = torch.rand(3, batch, len(lim), generator=gen)
    return ((lo + (hi - lo) * u[0]).to(dev), ((2 * u[1] - 1) * 0.2 * vel).to(dev), ((2 * u[2] - 1) * 0.4 * vel).to(dev))


def _median_region_ms(stream, replay, reps, barrier):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    out = []
    for _ in range(reps):
        barrier()
        torch.cuda._sleep(400_000)
        e0.record(stream)
        replay()
        e1.record(stream)
        stream.synchronize()
        out.append(e0.elapsed_time(e1))
    return statistics.median(out)


def config3(dev, barrier, batch=65536, steps=64, reps=5):
    """BASELINE config 3 (single GPU): Franka Panda inverse dynamics (RNEA, gravity + damping), 65 536 per launch, four
    independent launches in flight (graph branches), inputs rotated over buffer sets larger than L2.  The model is constant,
    so its table is folded once (drmb200_fold_link_table) like model.compute_inverse_dynamics does by itself."""
    m = drm.DifferentiableFrankaPanda(device=dev)
    table, topo = m._link_table(), m._topology
    folded = engine.fold_link_table(topo, table)
    n = m._n_dofs
    R = 24                                                 # 24 x 7.3 MB > L2
    sets = [sample(m, batch, 300 + r, dev) for r in range(R)]
    outs = [torch.empty(batch, n, device=dev) for _ in range(R)]
    stream = torch.cuda.Stream(device=dev)
    side = [torch.cuda.Stream(device=dev) for _ in range(3)]

    def step(i):
        engine.inverse_dynamics_raw(topo, table, *sets[i % R], 3, out=outs[i % R], folded=folded)

    with torch.cuda.stream(stream):
        for i in range(4):
            step(i)
        stream.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream):
            fork = torch.cuda.Event()
            fork.record(stream)
            for s in side:
                s.wait_event(fork)
            for i in range(steps):
                if i % 4 == 0:
                    step(i)
                else:
                    with torch.cuda.stream(side[i % 4 - 1]):
                        step(i)
            for s in side:
                j = torch.cuda.Event()
                j.record(s)
                stream.wait_event(j)
        g.replay()
        stream.synchronize()
        ms = _median_region_ms(stream, g.replay, reps, barrier) / steps
    return {"batch_per_launch": batch, "ms_per_launch": ms, "configs_per_s": batch / (ms * 1e-3), "algorithmic_bytes_per_config": 16 * n,
            "table": "folded once" if folded is not None else "as given", "launch": "CUDA graph, 4 branches", "steps_per_region": steps,
            "reps": reps}


def config4(dev, rank, barrier, per_gpu=32768, steps=64, reps=5):
    """-> {fused_ms_per_step, per_tip_ms_per_step (4 launches)} on this rank's shard."""
    m = drm.DifferentiableRobotModel(os.path.join(drm.robot_model.robot_description_folder, ALLEGRO), "allegro", device=dev)
    table, topo = m._link_table(), m._topology
    ees = [m._name_to_idx_map[t] for t in TIPS]
    n = m._n_dofs
    R = 6                                                  # 6 x 56 MB of outputs > L2
    qs = [sample(m, per_gpu, 100 * rank + r, dev)[0] for r in range(R)]
    outs = [(torch.empty(4, per_gpu, 3, device=dev), torch.empty(4, per_gpu, 4, device=dev), torch.empty(4, per_gpu, 3, n, device=dev),
             torch.empty(4, per_gpu, 3, n, device=dev)) for _ in range(R)]
    stream = torch.cuda.Stream(device=dev)
    res = {}
    engine.set_option("fk_pdl", 2)                         # stream of independent batches: launches overlap (drm_b200.h "fk_pdl")
    with torch.cuda.stream(stream):
        for name in ("fused", "per_tip"):
            def step(i):
                if name == "fused":
                    engine.fk_jacobian_multi_raw(topo, ees, table, qs[i % R], out=outs[i % R])
                else:
                    o = outs[i % R]
                    for e, ee in enumerate(ees):
                        engine.fk_jacobian_raw(topo, ee, table, qs[i % R], out=(o[0][e], o[1][e], o[2][e], o[3][e]))
            for i in range(3):
                step(i)
            stream.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=stream):
                for i in range(steps):
                    step(i)
            g.replay()
            stream.synchronize()
            res[f"{name}_ms_per_step"] = _median_region_ms(stream, g.replay, reps, barrier) / steps
            del g
    engine.set_option("fk_pdl", 0)
    res.update({"per_gpu_batch": per_gpu, "algorithmic_bytes_per_config": 4 * n + 4 * (28 + 24 * n), "steps_per_region": steps, "reps": reps})
    return res


def config5(dev, rank, world, dist, barrier, per_gpu=131072, steps=20, reps=5):
    """-> {ms_per_step, allreduce_scalars}: graph A (zero grad, forward, loss, backward) -> all-reduce of the flat gradient
    -> graph B (fused Adam), per step."""
    torch.manual_seed(7)                                   # identical initial parameters on every rank
    m = drm.DifferentiableKUKAiiwa(device=dev)
    for i in range(1, 8):
        b = m._bodies[i]
        m.make_link_param_learnable(b.name, "mass", UnconstrainedScalar(init_val=b.inertia.mass().detach().clone()))
        m.make_link_param_learnable(b.name, "com", UnconstrainedTensor(1, 3, init_tensor=b.inertia.com().detach().clone()))
        m.make_link_param_learnable(b.name, "inertia_mat", UnconstrainedTensor(3, 3, init_tensor=b.inertia.inertia_mat().detach().clone().reshape(3, 3)))
    flat = m.fuse_learnable_parameters()
    q, qd, qdd = sample(m, per_gpu, 1000 + rank, dev)
    target = torch.randn(per_gpu, 7, device=dev)
    from differentiable_robot_model_b200.parallel import PeerAllReduceAdam
    peer_opt = PeerAllReduceAdam(flat, lr=1e-3)           # NVLink peer-memory all-reduce fused with Adam: one kernel
    opt = torch.optim.Adam([flat], lr=1e-3, capturable=True, fused=True)      # the NCCL + torch baseline beside it
    flat.grad = torch.zeros_like(flat)
    stream = torch.cuda.Stream(device=dev)

    def fwd_bwd():
        flat.grad.zero_()
        with m.shared_link_table():
            pos, quat, jl, ja = m.compute_fk_and_jacobian(q, "iiwa_link_ee")
            tau = m.compute_inverse_dynamics(q, qd, qdd)
        loss = torch.nn.functional.mse_loss(tau, target) + pos.square().mean()     # fused forward / backward kernels of torch
        loss.backward()
        return loss

    with torch.cuda.stream(stream):
        for _ in range(3):
            fwd_bwd()
            if world > 1:
                dist.all_reduce(flat.grad)
            opt.step()
        stream.synchronize()
        mode = ("one CUDA graph per step: zero grad, FK+Jacobian, RNEA, loss, backward to one flat gradient, then ONE kernel: "
                "peer-memory SUM all-reduce of the gradient over NVLink fused with the Adam update (csrc/comm.cu)")
        g1 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g1, stream=stream):
            loss = fwd_bwd()
            peer_opt.step()

        def region():
            for _ in range(steps):
                g1.replay()

        nccl_ms = None
        if world > 1:                                      # the baseline: same step with an NCCL all-reduce + torch's fused Adam
            try:
                g2 = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g2, stream=stream):
                    fwd_bwd()
                    dist.all_reduce(flat.grad)
                    opt.step()

                def region_nccl():
                    for _ in range(steps):
                        g2.replay()
                region_nccl()
                stream.synchronize()
                nccl_ms = _median_region_ms(stream, region_nccl, reps, barrier) / steps
            except Exception as exc:
                nccl_ms = f"capture failed: {type(exc).__name__}"

        region()
        stream.synchronize()
        ms = _median_region_ms(stream, region, reps, barrier) / steps
    assert not peer_opt.peer_timeout(), "a peer did not arrive in the fused all-reduce"
    return {"per_gpu_batch": per_gpu, "ms_per_step": ms, "ms_per_step_nccl_allreduce_torch_adam": nccl_ms,
            "allreduce_scalars": int(flat.numel()) if world > 1 else 0,
            "final_loss": float(loss), "algorithmic_bytes_per_config": 420, "steps_per_region": steps, "reps": reps,
            "step": mode}


if __name__ == "__main__":
    import json
    d = torch.device("cuda", 0)
    torch.cuda.set_device(d)
    print(json.dumps({"config3": config3(d, torch.cuda.synchronize), "config4": config4(d, 0, torch.cuda.synchronize),
                      "config5": config5(d, 0, 1, None, torch.cuda.synchronize)}))
