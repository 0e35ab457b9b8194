"""Profiling helper: which CTAs are the slowest in the gate phase of a step, and what rows (in-degrees) they were dealt.
The row listing reproduces the plain dealing rule (every CTA takes rows); in steps where long in-edge lists get CTAs of
their own (<= 4096 rows, see scan_heavy in csrc/sweep.cu) the first CTAs hold those nodes instead — the times are right
either way.
    python tools/trace_gate_balance.py [workload] [step ...]      (GPU box)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from dagnn_b200 import runtime as rt, _lib

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
want = [int(a) for a in sys.argv[2:]] or [8, 12, 17, 20]
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
    layers = wl["layers"]
    steps = L + layers - 1
    for _ in range(3):
        Hs, tr = rt.sweep(sched, X, packed, wl["emb"], wl["hid"], layers, nvid, wl["kind"] == "code2", trace_steps=steps)
    torch.cuda.synchronize()
tr = tr.cpu().numpy()[:, :148, :].astype(np.float64)
MHZ = 1965.0
lo = sched.lvl_off_host
dirs = len(lo)
rp = [sched.rowptr[d].cpu().numpy() for d in range(dirs)]
NG, NW = 148, 8
W = NG * NW
for s in want:
    t = tr[s + 1]
    gate = (t[:, 8] - t[:, 0]) / MHZ
    rows = [[] for _ in range(NG)]
    rbase = 0
    for q in range(dirs * layers):
        d, i = divmod(q, layers)
        l = s - i
        if l < 0 or l >= sched.num_levels[d]:
            continue
        p0, p1 = int(lo[d][l]), int(lo[d][l + 1])
        for r in range(p1 - p0):
            slot = (r + rbase) % W
            rows[slot % NG].append((slot // NG, q, int(rp[d][p0 + r + 1] - rp[d][p0 + r]) if l > 0 else 0))
        rbase += p1 - p0
    order = np.argsort(-gate)
    print("step %d: gate max %.1f med %.1f min %.1f us" % (s, gate.max(), np.median(gate), gate.min()))
    for c in list(order[:6]) + list(order[72:75]) + list(order[-2:]):
        print("   cta %3d  %5.1f us  rows (warp, seg, in-edges): %s" % (c, gate[c], rows[c]))
