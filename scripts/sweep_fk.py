#!/usr/bin/env python
"""A/B sweep of the FK+Jacobian kernel's tuning knobs (staging variant, tile size, unrolled / rolled) on the
Kuka iiwa, at the BASELINE batch (65 536, graph-replayed with several launches in flight) and at 2^22 per launch.
Prints one JSON object; copy to profiles/ to have it judged."""
import itertools
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import differentiable_robot_model_b200 as drm  # noqa: E402
from differentiable_robot_model_b200 import engine  # noqa: E402
from oracle import drm_oracle as O  # noqa: E402

DEV = torch.device("cuda", 0)
BYTES = 224


def main():
    m = drm.DifferentiableKUKAiiwa(device=DEV)
    robot = O.load_robot(m.urdf_path, torch.float32)
    table, topo, ee = m._link_table(), m._topology, m._name_to_idx_map["iiwa_link_ee"]
    small, big = 65536, 1 << 22
    R = 16
    qs = [O.sample_inputs(robot, small, seed=r)[0].to(DEV) for r in range(R)]
    outs = [(torch.empty(small, 3, device=DEV), torch.empty(small, 4, device=DEV), torch.empty(small, 3, 7, device=DEV),
             torch.empty(small, 3, 7, device=DEV)) for _ in range(R)]
    q_big = torch.cat(qs * (big // small // R))
    out_big = (torch.empty(big, 3, device=DEV), torch.empty(big, 4, device=DEV), torch.empty(big, 3, 7, device=DEV),
               torch.empty(big, 3, 7, device=DEV))
    stream = torch.cuda.Stream(device=DEV)
    rows = []
    combos = [(1, 0, 1), (1, 0, 2), (1, 0, 0), (1, 1, 1), (0, 0, 1)]   # (staging, unrolled, packed: 1 row pairs, 2 two configs/thread)
    if os.environ.get("SWEEP_ONLY_PACKED"):
        combos = [(1, 0, 1), (1, 0, 2)]
    for (variant, unroll, packed), tile in itertools.product(combos, (64, 128, 256)):
        engine.set_option("fk_variant", variant)
        engine.set_option("fk_unroll", unroll)
        engine.set_option("fk_packed", packed)
        engine.set_option("fk_tile", tile)
        row = {"staging": "tma_bulk" if variant else "coop", "unrolled": bool(unroll), "packed_f32x2": int(packed),
               "tile": tile}
        with torch.cuda.stream(stream):
            # large batch
            for _ in range(3):
                engine.fk_jacobian_raw(topo, ee, table, q_big, out=out_big)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(10):
                engine.fk_jacobian_raw(topo, ee, table, q_big, out=out_big)
            e1.record(stream)
            stream.synchronize()
            ms = e0.elapsed_time(e1) / 10
            row["big_Gcfg_s"] = big / ms / 1e6
            row["big_hbm_GBps"] = big * BYTES / ms / 1e6
            # BASELINE batch: graph, inflight lanes
            for inflight in (1, 4):
                side = [torch.cuda.Stream(device=DEV) for _ in range(inflight - 1)]
                for i in range(R):
                    engine.fk_jacobian_raw(topo, ee, table, qs[i], out=outs[i])
                stream.synchronize()
                g = torch.cuda.CUDAGraph()
                nodes = 256
                with torch.cuda.graph(g, stream=stream):
                    fork = torch.cuda.Event()
                    fork.record(stream)
                    for s in side:
                        s.wait_event(fork)
                    for i in range(nodes):
                        lane = i % inflight
                        if lane == 0:
                            engine.fk_jacobian_raw(topo, ee, table, qs[i % R], out=outs[i % R])
                        else:
                            with torch.cuda.stream(side[lane - 1]):
                                engine.fk_jacobian_raw(topo, ee, table, qs[i % R], out=outs[i % R])
                    for s in side:
                        j = torch.cuda.Event()
                        j.record(s)
                        stream.wait_event(j)
                for _ in range(3):
                    g.replay()
                stream.synchronize()
                e0.record(stream)
                reps = 40
                for _ in range(reps):
                    g.replay()
                e1.record(stream)
                stream.synchronize()
                us = e0.elapsed_time(e1) * 1e3 / (reps * nodes)
                row[f"small_us_per_launch_inflight{inflight}"] = us
                row[f"small_Gcfg_s_inflight{inflight}"] = small / us / 1e3
        rows.append(row)
    engine.set_option("fk_variant", 1); engine.set_option("fk_unroll", 2); engine.set_option("fk_packed", 1); engine.set_option("fk_tile", 0)
    print(json.dumps({"gpu": torch.cuda.get_device_name(0), "rows": rows}))


if __name__ == "__main__":
    main()
