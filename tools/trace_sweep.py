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
tr = tr.cpu().numpy()[:, :148, :]
MHZ = 1965.0
lo = sched.lvl_off_host
print("all times in us. first tile of CTA 0: pre = rowptr + softmax weights, wait = free operand stage, build = gather+split+store, "
      "hand = proxy fence + arrive; issuer: wB = wait weights, wA = wait operands, iss = MMA issue")
print("step | level sizes | U rows | tiles/CTA | build  acc  epi  rest (max over CTAs) | step || slowest CTA: pre wait slow+store hand comb pref | wA iss")
tot = 0
for s in range(steps):
    t = tr[s].astype(np.float64)
    n0 = [int(lo[d][s + 1] - lo[d][s]) if s < L else 0 for d in range(len(lo))]
    act = t[:, 6] > 0
    g = np.where(act, (t[:, 1] - t[:, 0]) / MHZ, 0); bw = np.where(act, (t[:, 2] - t[:, 1]) / MHZ, 0)
    pm = np.where(act, (t[:, 3] - t[:, 2]) / MHZ, 0); tl = np.where(act, (t[:, 4] - t[:, 3]) / MHZ, 0)
    end = t[:, 5] if s + 1 < steps else t[:, 4]
    stepdur = ((end - t[:, 0]) / MHZ).max()
    tot += stepdur
    k = int(np.argmax(np.where(act, t[:, 4] - t[:, 0], -1))) if act.any() else 0     # slowest CTA of the step
    f = int(t[0, 7])
    print("%3d | %12s | %2d %3d | %2d | %6.1f %6.1f %6.1f %7.1f | %7.1f || %5.1f %5.1f %5.1f %5.1f %5.1f %5.1f | %5.1f %5.1f" % (
        s, n0, f & 255, f >> 8, int(t[:, 6].max()), g.max(), bw.max(), pm.max(), tl.max(), stepdur,
        t[k, 8] / MHZ, t[k, 9] / MHZ, t[k, 10] / MHZ, t[k, 11] / MHZ, t[k, 12] / MHZ, t[k, 13] / MHZ, t[k, 14] / MHZ, t[k, 15] / MHZ))
    if len(sys.argv) > 2 and s == int(sys.argv[2]):       # per-CTA dump of one step
        for c in range(148):
            print("   cta %3d tiles %d | build %.1f acc %.1f epi %.1f rest %.1f | pre %.1f wait %.1f build %.1f hand %.1f | wB %.1f wA %.1f iss %.1f stg %d" % (
                c, int(t[c, 6]), g[c], bw[c], pm[c], tl[c], t[c, 8] / MHZ, t[c, 9] / MHZ, t[c, 10] / MHZ, t[c, 11] / MHZ, t[c, 12] / MHZ,
                t[c, 13] / MHZ, t[c, 14] / MHZ, int(t[c, 15])))
print("sum of step durations: %.1f us" % tot)
