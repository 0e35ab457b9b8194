"""Fine-grained stamps of the projection tile of tail steps (needs a -DDAGNN_TRACE_FINE build). GPU box."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from dagnn_b200 import runtime as rt, _lib
wl = bench.WORKLOADS["c2"]
_lib.build_library()
dev = torch.device("cuda:0")
B = bench.build_workload(wl, 1); m = bench.build_module(wl).to(dev); G = B.to(dev)
with torch.no_grad():
    X, Hs, sched = m.node_states(G); packed = m._pack(dev)
    L = sched.num_levels[0]; steps = L + 1
    for _ in range(3):
        Hs, tr = rt.sweep(sched, X, packed, 256, 256, 2, 0, True, trace_steps=steps)
    torch.cuda.synchronize()
tr = tr.cpu().numpy()[:, :148, :].astype(np.float64) / 1965.0
for s in (45, 55, 64):
    t = tr[s + 1]
    for c in (0, 1, 40, 90):
        if t[c, 6] * 1965 > 0:
            print("step %d cta %d (us from phase start): gate-done %.2f | bar1 %.2f || issuer: start %.2f A-issued %.2f B0-landed %.2f A0-landed %.2f all-MMA-issued %.2f || workers: at-epilogue-wait %.2f acc-ready %.2f stored %.2f | bar2 %.2f" % (
                s, c, t[c, 8] - t[c, 0], t[c, 9] - t[c, 0], t[c, 10], t[c, 11], t[c, 12], t[c, 13], t[c, 14],
                t[c, 1] - t[c, 0], t[c, 2] - t[c, 0], t[c, 3] - t[c, 0], t[c, 5] - t[c, 0]))
