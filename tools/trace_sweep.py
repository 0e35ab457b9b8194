"""Profiling helper: per-wavefront-step timeline of the persistent sweep kernel (clock64 stamps written by the kernel).
    python tools/trace_sweep.py [workload]      (GPU box)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from dagnn_b200 import runtime as rt, _lib

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
_lib.build_library()
dev = torch.device("cuda:0")
B = bench.build_workload(wl, 1)
m = bench.build_module(wl).to(dev)
G = B.to(dev)
with torch.no_grad():
    X, Hs, sched = m.node_states(G)
    packed = m._pack(dev) if wl["kind"] == "code2" else m._packed
    nvid = wl["emb"] if wl["kind"] == "NA" else 0
    L = sched.num_levels[0]
    steps = L + wl["layers"] - 1
    for _ in range(3):
        Hs, tr = rt.sweep(sched, X, packed, wl["emb"], wl["hid"], wl["layers"], nvid, wl["kind"] == "code2", trace_steps=steps)
    torch.cuda.synchronize()
tr = tr.cpu().numpy()[:, :148, :]
MHZ = 1965.0
lo = sched.lvl_off_host
print("step | level sizes d0/d1 | flags | tiles/CTA max | phaseG  barrier  phaseM  tiles (us, max over CTAs) | step us")
tot = 0
for s in range(steps):
    t = tr[s].astype(np.float64)
    n0 = [int(lo[d][s + 1] - lo[d][s]) if s < L else 0 for d in range(len(lo))]
    has_small = t[:, 1].max() > 0
    if has_small:
        g = (t[:, 1] - t[:, 0]) / MHZ; bw = (t[:, 2] - t[:, 1]) / MHZ; pm = (t[:, 3] - t[:, 2]) / MHZ; tl = (t[:, 4] - t[:, 3]) / MHZ
    else:
        g = bw = pm = np.zeros(148); tl = (t[:, 4] - t[:, 0]) / MHZ
    end = t[:, 5] if s + 1 < steps else t[:, 4]
    stepdur = ((end - t[:, 0]) / MHZ).max()
    tot += stepdur
    print("%3d | %12s | %3d | %2d | %6.1f %6.1f %6.1f %7.1f | %7.1f" % (s, n0, int(t[0, 7]), int(t[:, 6].max()), g.max(), bw.min(), pm.max(), tl.max(), stepdur))
print("sum of step durations: %.1f us" % tot)
