#!/usr/bin/env python
"""BASELINE config 1: per-call latency of the Python API at batch 1 (2-link toy robot, compute_forward_kinematics), where
the kernel is ~2 us and everything else is host overhead: tensor_check, link-table lookup, ctypes, launch.

  api_us_async      wall time per call, calls issued back to back without synchronising (dispatch cost)
  api_us_sync       wall time per call including a device synchronise after every call (what a control loop sees)
  raw_us_async      the same through engine.fk_jacobian_raw with preallocated outputs (no tensor_check / allocation)
  reference_cpu_us  the unmodified reference (baseline/_ref) on this host's CPU, same call, if installed
Prints one JSON object."""
import json
import os
import sys
import time

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import differentiable_robot_model_b200 as drm  # noqa: E402
from differentiable_robot_model_b200 import engine  # noqa: E402


def per_call_us(fn, n, sync_each=False):
    for _ in range(50):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
        if sync_each:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6


def main():
    out = {}
    for name, cls, link, n in (("2link_robot", drm.DifferentiableTwoLinkRobot, "endEffector", 2), ("kuka_iiwa", drm.DifferentiableKUKAiiwa, "iiwa_link_ee", 7)):
        m = cls(device="cuda:0")
        q = torch.zeros(1, n, device="cuda:0") + 0.3
        q1 = torch.zeros(n, device="cuda:0") + 0.3
        table, topo, ee = m._link_table(), m._topology, m._name_to_idx_map[link]
        outs = (torch.empty(1, 3, device="cuda:0"), torch.empty(1, 4, device="cuda:0"), None, None)
        with torch.no_grad():
            res = {"api_us_async": per_call_us(lambda: m.compute_forward_kinematics(q, link), 5000),
                   "api_us_async_1d_input": per_call_us(lambda: m.compute_forward_kinematics(q1, link), 5000),
                   "api_us_sync": per_call_us(lambda: m.compute_forward_kinematics(q, link), 2000, sync_each=True),
                   "raw_us_async": per_call_us(lambda: engine.fk_jacobian_raw(topo, ee, table, q, want_jac=False, out=outs), 5000)}
        out[name] = res
    ref_dir = os.path.join(REPO, "baseline", "_ref")
    if os.path.isdir(os.path.join(ref_dir, "differentiable_robot_model")):
        sys.path.insert(0, os.path.join(REPO, "oracle", "refshim"))
        sys.path.insert(0, ref_dir)
        import contextlib
        import io
        from differentiable_robot_model.robot_model import DifferentiableTwoLinkRobot as RefToy
        with contextlib.redirect_stdout(io.StringIO()):
            ref = RefToy()
        qc = torch.zeros(1, 2) + 0.3
        torch.set_num_threads(1)
        with torch.no_grad():
            for _ in range(20):
                ref.compute_forward_kinematics(qc, "endEffector")
            t0 = time.perf_counter()
            for _ in range(200):
                ref.compute_forward_kinematics(qc, "endEffector")
            out["2link_robot"]["reference_cpu_us"] = (time.perf_counter() - t0) / 200 * 1e6
    print(json.dumps(out))


if __name__ == "__main__":
    main()
