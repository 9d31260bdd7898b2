#!/usr/bin/env python
"""Where do the ~6 us go that a 20-launch region costs beyond 6.8 + 19 x 2.78 us?  Times the same 20-node PDL graph
(a) with events around the graph launch (bench.py), (b) with EXTERNAL event-record nodes captured as the first and last
nodes of the graph (device time from the moment the graph starts executing), for 1 / 2 / 4 chains.  Prints JSON."""
import json
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import differentiable_robot_model_b200 as drm  # noqa: E402
from differentiable_robot_model_b200 import engine  # noqa: E402

DEV = torch.device("cuda", 0)


def main():
    m = drm.DifferentiableKUKAiiwa(device=DEV)
    table, topo, ee = m._link_table(), m._topology, m._name_to_idx_map["iiwa_link_ee"]
    B, R, K = 65536, 16, 20
    gen = torch.Generator().manual_seed(0)
    qs = [((torch.rand(B, 7, generator=gen) * 2 - 1) * 2.9).to(DEV) for _ in range(R)]
    outs = [(torch.empty(B, 3, device=DEV), torch.empty(B, 4, device=DEV), torch.empty(B, 3, 7, device=DEV), torch.empty(B, 3, 7, device=DEV)) for _ in range(R)]
    stream = torch.cuda.Stream(device=DEV)
    side = [torch.cuda.Stream(device=DEV) for _ in range(3)]
    engine.set_option("fk_pdl", 2)
    res = {}
    with torch.cuda.stream(stream):
        for i in range(R):
            engine.fk_jacobian_raw(topo, ee, table, qs[i], out=outs[i])
        stream.synchronize()
        for branches in (1, 2, 4):
            for inside in (False, True):
                e0 = torch.cuda.Event(enable_timing=True, external=inside)
                e1 = torch.cuda.Event(enable_timing=True, external=inside)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=stream):
                    if inside:
                        e0.record(stream)
                    if branches > 1:
                        fork = torch.cuda.Event(); fork.record(stream)
                        for s in side[:branches - 1]:
                            s.wait_event(fork)
                    for i in range(K):
                        lane = i % branches
                        if lane == 0:
                            engine.fk_jacobian_raw(topo, ee, table, qs[i % R], out=outs[i % R])
                        else:
                            with torch.cuda.stream(side[lane - 1]):
                                engine.fk_jacobian_raw(topo, ee, table, qs[i % R], out=outs[i % R])
                    if branches > 1:
                        for s in side[:branches - 1]:
                            j = torch.cuda.Event(); j.record(s); stream.wait_event(j)
                    if inside:
                        e1.record(stream)
                for _ in range(3):
                    g.replay()
                stream.synchronize()
                t = []
                for _ in range(15):
                    torch.cuda.synchronize()
                    torch.cuda._sleep(600_000)
                    if not inside:
                        e0.record(stream)
                    g.replay()
                    if not inside:
                        e1.record(stream)
                    stream.synchronize()
                    t.append(e0.elapsed_time(e1) * 1e3)
                res[f"branches{branches}_{'events_in_graph' if inside else 'events_around_launch'}_us"] = statistics.median(t)
                del g
    print(json.dumps(res))


if __name__ == "__main__":
    main()
