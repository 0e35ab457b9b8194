"""Profiling helper (library built with -DDAGNN_GATE_TRACE): stage latencies of the first node warp 0 of every CTA
handles in a gate phase. Slots: 0 phase start, 10 node start, 11 row pointers there, 12 scores there, 13 first rows there,
14 Gi + sums there, 15 node done, 8 gate phase done (whole CTA).
    python tools/trace_gate_stages.py [workload] [step ...]      (GPU box)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from dagnn_b200 import runtime as rt, _lib

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
want = [int(a) for a in sys.argv[2:]] or [3, 8, 17, 20, 30]
dev = torch.device("cuda:0")
B = bench.build_workload(wl, 1)
m = bench.build_module(wl).to(dev)
G = B.to(dev)
with torch.no_grad():
    X, Hs, sched = m.node_states(G)
    packed = m._pack(dev) if wl["kind"] == "code2" else m._packed
    nvid = wl["emb"] if wl["kind"] == "NA" else 0
    steps = sched.num_levels[0] + wl["layers"] - 1
    for _ in range(3):
        Hs, tr = rt.sweep(sched, X, packed, wl["emb"], wl["hid"], wl["layers"], nvid, wl["kind"] == "code2", trace_steps=steps)
    torch.cuda.synchronize()
tr = tr.cpu().numpy()[:, :148, :].astype(np.float64)
MHZ = 1965.0
names = ["sync->node", "rowptr", "scores", "rows", "gi+softmax", "finish", "node->phase end"]
for s in want:
    if s >= steps:
        continue
    t = tr[s + 1]
    ok = t[:, 15] > 0
    cols = [t[:, 10] - t[:, 0], t[:, 11] - t[:, 10], t[:, 12] - t[:, 11], t[:, 13] - t[:, 12], t[:, 14] - t[:, 13],
            t[:, 15] - t[:, 14], t[:, 8] - t[:, 15]]
    ok &= (t[:, 10] > 0) & (t[:, 11] >= t[:, 10]) & (t[:, 15] >= t[:, 14])
    print("step %d (%d CTAs stamped): median / p90 in us" % (s, ok.sum()))
    if ok.sum() == 0:
        continue
    if ok.sum() < 40:      # the selective trace modes: list the nodes one by one
        for c in np.nonzero(ok)[0][:10]:
            print("   cta %3d: " % c + "  ".join("%s %.2f" % (n, col[c] / MHZ) for n, col in zip(names, cols)))
    for n, c in zip(names, cols):
        c = c[ok] / MHZ
        print("   %-16s %6.2f %6.2f" % (n, np.median(c), np.percentile(c, 90)))
