#!/usr/bin/env python
"""End-to-end FK+Jacobian through drmb200_fk_jacobian_host with page-locked host buffers: fused (the kernel's TMA copies
read / write host memory directly, one launch) vs staged (H2D -> kernel -> D2H on three streams).  Prints one JSON object."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import differentiable_robot_model_b200 as drm  # noqa: E402
from differentiable_robot_model_b200 import engine  # noqa: E402
from oracle import drm_oracle as O  # noqa: E402

DEV = torch.device("cuda", 0)


def main():
    m = drm.DifferentiableKUKAiiwa(device=DEV)
    robot = O.load_robot(m.urdf_path, torch.float32)
    table, topo, ee = m._link_table(), m._topology, m._name_to_idx_map["iiwa_link_ee"]
    rows = []
    for batch, steps in ((65536, 200), (1 << 20, 12)):
        q_host = [O.sample_inputs(robot, batch, seed=r)[0].pin_memory() for r in range(2)]
        outs = [(torch.empty(batch, 3).pin_memory(), torch.empty(batch, 4).pin_memory(), torch.empty(batch, 3, 7).pin_memory(),
                 torch.empty(batch, 3, 7).pin_memory()) for _ in range(2)]
        for fused in (1, 0, 1, 0):
            engine.set_option("host_fused", fused)
            for i in range(3):
                engine.fk_jacobian_host(topo, ee, 0, table, q_host[i % 2], *outs[i % 2])
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for i in range(steps):
                engine.fk_jacobian_host(topo, ee, 0, table, q_host[i % 2], *outs[i % 2])
            dt = time.perf_counter() - t0
            rows.append({"batch": batch, "mode": "fused" if fused else "staged", "configs_per_s": steps * batch / dt,
                         "us_per_call": dt / steps * 1e6, "pcie_d2h_GBps": steps * batch * 196 / dt / 1e9,
                         "pcie_h2d_GBps": steps * batch * 28 / dt / 1e9})
    engine.set_option("host_fused", 1)
    print(json.dumps({"gpu": torch.cuda.get_device_name(0), "rows": rows}))


if __name__ == "__main__":
    main()
