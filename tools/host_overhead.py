"""Profiling helper: host (Python + ctypes) time per stage of one forward, launches asynchronous, GPU otherwise idle.
    python tools/host_overhead.py [workload]      (GPU box)"""
import os, sys, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from dagnn_b200 import runtime as rt

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
dev = torch.device("cuda:0")
B = bench.build_workload(wl, 1)
m = bench.build_module(wl).to(dev)
G = B.to(dev)
with torch.no_grad():
    for _ in range(5):
        bench.hot_path(m, G, wl)
    torch.cuda.synchronize()
    acc = {}
    for rep in range(20):
        torch.cuda.synchronize()
        t = [time.perf_counter()]
        X = m.encoder(G.x, G.node_depth.view(-1, )); t.append(time.perf_counter())
        sched = m.build_schedule(G, 256); t.append(time.perf_counter())
        packed = m._pack(dev); t.append(time.perf_counter())
        Hs = rt.sweep(sched, X, packed, m.emb_dim, m.hidden_dim, m.num_layers, 0, m.w_edge_attr); t.append(time.perf_counter())
        out = m.readout(G, X, Hs, sched); t.append(time.perf_counter())
        sched.finalize(); t.append(time.perf_counter())
        for k, (a, b) in zip(["encoder", "schedule", "pack", "sweep launch", "readout", "finalize (incl. waiting for the GPU)"], zip(t, t[1:])):
            acc.setdefault(k, []).append((b - a) * 1e6)
    for k, v in acc.items():
        v.sort()
        print("%-40s median %7.1f us   min %7.1f" % (k, v[len(v) // 2], v[0]))
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(20):
        bench.hot_path(m, G, wl)
    pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(18)
