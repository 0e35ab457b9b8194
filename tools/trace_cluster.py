"""Profiling helper: per-level timeline of the cluster-resident sweep (clock64 stamps written by k_sweep_cluster; slots in
dagnn_b200/csrc/sweep_cluster.cu: 0 level start, 1 gathered, 2 exchanged, 3 projected + cells done, 4 level closed, 5 / 7 / 8
copy warp done / MMA warp done / first accumulators complete).
    python tools/trace_cluster.py [workload]      (GPU box)"""
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
    for _ in range(3):
        Hs, tr = rt.sweep(sched, X, packed, wl["emb"], wl["hid"], wl["layers"], nvid, wl["kind"] == "code2", trace_steps=L)
    torch.cuda.synchronize()
tr = tr.cpu().numpy().astype(np.float64)          # [L + 1][256][16]
MHZ = 1965.0
dirs = 2 if wl["bidir"] else 1
items = dirs * wl["layers"]
ncta = int((tr[0, :, 0] > 0).sum())
G_ = max(1, ncta // (8 * items))
t00 = tr[:L, :ncta, 0][tr[:L, :ncta, 0] > 0].min()
print("clusters: %d items x %d groups, %d CTAs; times in us; per cluster (CTA 0 of it): level | rows? | gather exch proj close | level total | start offset" % (items, G_, ncta))
for c in range(items * G_):
    cta = 8 * c
    g, di = c % G_, c // G_
    i, d = di // dirs, di % dirs
    tot = 0.0
    print("--- cluster %d: layer %d dir %d group %d" % (c, i, d, g))
    for l in range(L):
        t = tr[l, cta]
        if t[0] == 0:
            continue
        if t[4] == 0 or t[4] < t[0]:
            continue
        # max over the cluster's CTAs of each phase end
        tc = tr[l, cta:cta + 8]
        ga = (tc[:, 1] - tc[:, 0]).max() / MHZ
        ex = (tc[:, 2] - tc[:, 1]).max() / MHZ
        pr = (tc[:, 3] - tc[:, 2]).max() / MHZ
        cl = (tc[:, 4] - tc[:, 3]).max() / MHZ
        lv = (tc[:, 4].max() - tc[:, 0].min()) / MHZ
        tot += lv
        iss = ""
        if t[5] > 0 and t[7] > 0 and t[8] > 0:
            iss = " | after barrier: copies started %5.2f, MMAs issued %5.2f, first accumulators seen %5.2f, last cells stored %5.2f" % (
                (t[5] - t[2]) / MHZ, (t[7] - t[2]) / MHZ, (t[8] - t[2]) / MHZ, (t[3] - t[2]) / MHZ)
        print("%3d | %6.2f %6.2f %6.2f %6.2f | %7.2f | %8.1f%s" % (l, ga, ex, pr, cl, lv, (t[0] - t00) / MHZ, iss))
    print("    sum of level durations %.1f us" % tot)
end = tr[:L, :ncta, 4].max()
print("first start -> last close: %.1f us" % ((end - t00) / MHZ))
k0 = tr[L, :ncta, 0]
print("kernel entry (earliest CTA) -> last close: %.1f us; prologue (weights on chip, tables, level-0 input rows) per cluster, us:" % ((end - k0.min()) / MHZ))
for c in range(items * G_):
    cta = 8 * c
    g, di = c % G_, c // G_
    i, d = di // dirs, di % dirs
    lv = tr[:L, cta, :]
    done = lv[:, 4].max()
    print("  cluster %2d (layer %d dir %d group %d): entry +%.1f, prologue %.1f, first level starts +%.1f, last level closes +%.1f" % (
        c, i, d, g, (tr[L, cta, 0] - k0.min()) / MHZ, (tr[L, cta:cta + 8, 1].max() - tr[L, cta, 0]) / MHZ,
        (lv[0, 0] - k0.min()) / MHZ, (done - k0.min()) / MHZ))
