"""Profiling helper: per-wavefront-step timeline of the persistent sweep kernel (clock64 stamps written by the kernel,
layout in include/dagnn_b200.h, DagnnSweepArgs.trace).
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
tr = tr.cpu().numpy()[:, :148, :].astype(np.float64)
MHZ = 1965.0
lo = sched.lvl_off_host
print("all times in us, max over CTAs unless noted. row s = gate phase of step s, then projection of the rows it produced (row -1: X)")
print("step | level sizes | gate(max) gate(med) bar1 | cols rows tiles/CTA | build  acc  epi  rest | bar2(min) | step | issuer, first tile: start->operands->MMAs issued->acc seen")
tot = 0
for k in range(steps + 1):
    s = k - 1
    t = tr[k]
    n0 = [int(lo[d][s + 1] - lo[d][s]) if 0 <= s < L else 0 for d in range(len(lo))]
    act = t[:, 6] > 0
    gate = (t[:, 8] - t[:, 0]) / MHZ if s >= 0 else np.zeros(148)
    bar1 = (t[:, 9] - t[:, 8]) / MHZ if s >= 0 else np.zeros(148)
    p0 = t[:, 9] if s >= 0 else t[:, 0]
    g = np.where(act, (t[:, 1] - p0) / MHZ, 0); bw = np.where(act, (t[:, 2] - t[:, 1]) / MHZ, 0)
    pm = np.where(act, (t[:, 3] - t[:, 2]) / MHZ, 0); tl = np.where(act, (t[:, 4] - t[:, 3]) / MHZ, 0)
    end = np.maximum(t[:, 5], t[:, 4])
    bar2 = (t[:, 5] - t[:, 4]) / MHZ
    stepdur = ((end - t[:, 0]) / MHZ).max()
    tot += stepdur
    f = int(t[0, 7])
    i0 = np.where(act, (t[:, 10] - p0) / MHZ, 0); i1 = np.where(act, (t[:, 11] - t[:, 10]) / MHZ, 0)
    i2 = np.where(act, (t[:, 12] - t[:, 11]) / MHZ, 0); i3 = np.where(act, (t[:, 2] - t[:, 12]) / MHZ, 0)
    i4 = np.where(act, (t[:, 13] - t[:, 11]) / MHZ, 0); i5 = np.where(act, (t[:, 14] - t[:, 13]) / MHZ, 0)
    med = lambda x: np.median(x[act]) if act.any() else 0
    iss = " | %4.1f %4.1f %4.1f %4.1f (chunk 0 %4.2f, chunk 1 %4.2f)" % (med(i0), med(i1), med(i2), med(i3), med(i4), med(i5))
    print("%3d | %12s | %6.1f %6.1f %5.1f | %3d %3d %2d | %6.1f %6.1f %6.1f %6.1f | %5.1f | %7.1f" % (
        s, n0, gate.max(), np.median(gate), bar1.min() if s >= 0 else 0, f & 4095, f >> 12, int(t[:, 6].max()), g.max(), bw.max(), pm.max(), tl.max(),
        bar2.min(), stepdur) + iss)
print("sum of step durations: %.1f us" % tot)
