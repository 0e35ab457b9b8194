#!/usr/bin/env python
"""bench.py — graphs/s of the DAGNN layer-wise forward (encoder -> schedule -> level sweeps -> readout).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3|c5|na|bn|big] [--train-step]

Workload at N GPUs (weak scaling): BASELINE.json configs[1] — synthetic ogbg-code2-shaped batch, DAGNN 2 layers,
emb_dim = hidden = 256, bidirectional, attn_h, max-pool readout, 128 graphs PER GPU (global batch 128·N, built
with one seed on every rank and split with the reference's node-balanced rule; no data-path collective).
A step = one forward of the hot path over one batch.

  value : graphs/s, inputs resident in HBM, device-timed (CUDA events per step, L2 flushed between steps,
          max over ranks).
  e2e   : the same through the public module API with the batch in PINNED HOST memory: H2D of every input
          tensor + forward + D2H of the readout inside the timed region.
  roofline : the level-sweep kernel (k_sweep): algorithmic bytes of one sweep / device time of
          the sweep's launches, against the measured HBM peak (MEASURED_PEAKS.json, else the recipe's fallback).
  cpu_baseline : the oracle port of the reference's CPU path (oracle/dagnn_oracle.py — the reference is Python and
          needs PyG, which does not exist on the box) on the same batch, host cores of this box.
`--impl reference` times that CPU port alone (rank 0 only) and prints the same line with "impl": "reference".
`--train-step` times a whole training step instead (forward + multi-head cross entropy + backward + ONE NCCL all-reduce of the
flat fp32 gradient buffer + Adam step; BASELINE configs[4] with --workload c5): same line plus a "train" object.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

WORKLOADS = {
    # name: (graphs per GPU, emb, hidden, layers, bidirectional, kind, description)
    "c2": dict(graphs=128, emb=256, hid=256, layers=2, bidir=True, kind="code2",
               desc="ogbg-code2-shaped synthetic, DAGNN 2-layer emb_dim=hidden=256 bidirectional attn_h, batch=128/GPU"),
    "c3": dict(graphs=256, emb=300, hid=300, layers=5, bidir=True, kind="code2",
               desc="ogbg-code2-shaped synthetic, DAGNN 5-layer emb_dim=hidden=300 bidirectional attn_h, batch=256/GPU"),
    "c5": dict(graphs=128, emb=300, hid=300, layers=5, bidir=True, kind="code2",
               desc="ogbg-code2-shaped synthetic, DAGNN 5-layer emb_dim=hidden=300 bidirectional, batch=128/GPU (1024 over 8)"),
    "big": dict(graphs=4096, emb=256, hid=256, layers=2, bidir=True, kind="code2",
                desc="ogbg-code2-shaped synthetic, config-2 model at batch=4096/GPU (per-level kernel HBM evidence)"),
    "na": dict(graphs=32, emb=8, hid=501, layers=2, bidir=False, kind="NA", rows="na_real_hs501",
               desc="NA data set rows 1000..1031 of dvae/data/final_structures6.txt (8-node ENAS DAGs), D-VAE DAGNN hs=501 2-layer "
                    "unidirectional, batch=32"),
    "bn": dict(graphs=128, emb=10, hid=501, layers=2, bidir=True, kind="BN", rows="bn_real_hs501",
               desc="BN data set rows 0..127 of dvae/data/asia_200k.txt (10-node Bayesian networks), D-VAE DAGNN_BN hs=501 2-layer "
                    "bidirectional, batch=128"),
}
SEED = 20262
FALLBACK_HBM_GBS = 6650.0      # /opt/skills/guides/B200_PROFILING.md fallback


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-flush", action="store_true", help="do not flush L2 between steps (profiling runs)")
    ap.add_argument("--train-step", action="store_true", help="time forward + loss + backward + gradient all-reduce + optimizer step")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------ workload
def build_workload(wl, world):
    from dagnn_b200 import data as D
    if wl["kind"] == "code2":
        return D.make_code2_batch(wl["graphs"] * world, SEED)
    # real rows of the reference's data files, carried by the golden fixtures (tests/golden/*.npz, key "rows"); more ranks
    # than one repeat them
    z = np.load(os.path.join(ROOT, "tests", "golden", wl["rows"] + ".npz"), allow_pickle=False)
    rows = json.loads(str(z["rows"]))
    dec = D.decode_enas_row if wl["kind"] == "NA" else D.decode_bn_row
    need = wl["graphs"] * world
    return D.collate_dvae([dec(rows[k % len(rows)]) for k in range(need)])


def build_module(wl):
    from dagnn_b200 import data as D, dvae, ogb
    if wl["kind"] == "code2":
        enc = ogb.ASTNodeEncoder(wl["emb"], D.CODE2_NUM_NODETYPES, D.CODE2_NUM_NODEATTRS, D.CODE2_MAX_DEPTH)
        m = ogb.DAGNN(D.CODE2_NUM_VOCAB, D.CODE2_MAX_SEQ_LEN, wl["emb"], wl["hid"], None, encoder=enc,
                      num_layers=wl["layers"], bidirectional=wl["bidir"], out_wx=False, out_pool_all=False)
    else:
        nvt = wl["emb"]
        cls = dvae.DAGNN if wl["kind"] == "NA" else dvae.DAGNN_BN
        m = cls(nvt, wl["hid"], wl["hid"], nvt, nvt, 0, 1, hs=wl["hid"], nz=56, num_nodes=nvt, num_layers=wl["layers"],
                bidirectional=wl["bidir"])
    D.deterministic_init_(m, 1)
    return m.eval()


def hot_path(m, G, wl):
    """the path under test: node encoder -> integer schedule -> level sweeps (all directions) -> pooled readout."""
    if wl["kind"] == "code2":
        return m.forward_readout(G)
    return m(G)


def oracle_forward(p, B, wl):
    from oracle import dagnn_oracle as O       # the CPU baseline leg is one of the places allowed to execute oracle/
    if wl["kind"] == "code2":
        return O.ogb_forward(p, B, num_layers=wl["layers"], bidirectional=wl["bidir"], heads=False)[1]
    return O.dvae_forward(p, B, num_layers=wl["layers"], bidirectional=wl["bidir"], num_nodes=wl["emb"],
                          vid=(wl["kind"] == "NA"))[0]


def sweep_algorithmic_bytes(B, wl):
    """DESIGN.md §5 / SURVEY.md §8d: per direction d and layer i, fp32 rows + int32 indices, weights excluded:
    4·N·D_i (input rows) + 4·N·H (state rows written) + E'_d·(4·H + 4 + 1) (gathered predecessor rows + col index
    + type/valid) + 8·N (row pointer + node id); E'_d = in-edges (direction d) of nodes with level_d > 0."""
    if wl["kind"] == "code2":
        lv = [B._bi_layer_idx0.numpy(), B._bi_layer_idx1.numpy()]
    else:
        lv = [B.bi_layer_index[0][0].numpy(), B.bi_layer_index[1][0].numpy()]
    ei = B.edge_index.numpy()
    N, H = int(B.x.shape[0]), wl["hid"]
    total, edges = 0, []
    for d in range(2 if wl["bidir"] else 1):
        tgt = ei[1 - d]
        Ed = int((lv[d][tgt] > 0).sum())
        edges.append(Ed)
        for i in range(wl["layers"]):
            Di = wl["emb"] if i == 0 else H
            total += 4 * N * Di + 4 * N * H + Ed * (4 * H + 5) + 8 * N
    return total, edges


# ------------------------------------------------------------------------------------------ helpers
class ClockSampler(object):
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for k, nm in enumerate(names):
                    if r[5 + k].lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def ncu_traffic(workload, kernel):
    """(dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel`, where the number comes from) — from the committed
    `ncu --set full` capture (profiles/traffic.json: bytes per launch keyed by workload, tagged with the kernel and the commit it
    was captured at), or (None, reason). A capture of another kernel than the one this run launches is not reported."""
    try:
        j = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[workload]
        if j.get("kernel") != kernel:
            return None, "no ncu capture of %s for this workload (profiles/traffic.json holds %s)" % (kernel, j.get("kernel"))
        return float(j["dram_bytes_per_launch"]), "ncu --set full, %s at commit %s (profiles/traffic.json; 1 GPU, whole batch)" % (
            kernel, j.get("commit", "?"))
    except Exception as e:
        return None, "profiles/traffic.json: %s" % (e,)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            j = json.load(open(path))
            for k in ("hbm_gbs", "hbm_gb_s", "hbm_GBps"):
                if k in j:
                    return float(j[k]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md, MEASURED_PEAKS.json absent)"


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host core this process may run on."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    torch.set_num_threads(max(1, n))
    return torch.get_num_threads()


def cpu_baseline(p, B, wl, budget_s=25.0):
    """oracle port on this box's host cores: one untimed warm-up on a small slice, then full-batch forwards until
    ~budget_s is spent (at least 1, at most 5); graphs/s = graphs / median."""
    from dagnn_b200 import data as D
    ng = int(B.num_graphs)
    use_all_host_threads()
    with torch.no_grad():
        oracle_forward(p, D.select_graphs(B, range(min(4, ng))), wl)
        ts, t_all = [], time.perf_counter()
        while len(ts) < 5 and (not ts or time.perf_counter() - t_all + ts[-1] < budget_s):
            t0 = time.perf_counter()
            oracle_forward(p, B, wl)
            ts.append(time.perf_counter() - t0)
    med = float(np.median(ts))
    return {"value": ng / med, "unit": "graphs/s", "cores": int(torch.get_num_threads()), "host_cpus": os.cpu_count(),
            "kind": "port (reference-on-shim unavailable on the box)", "sample": "%d full-batch forwards of the same %d-graph batch (oracle/dagnn_oracle.py, torch CPU "
            "fp32, eval/no_grad), median %.3f s" % (len(ts), ng, med)}


# ------------------------------------------------------------------------------------------ reference arm
def run_reference(args, wl, rank, world):
    """The reference's CPU implementation of the path = its PyG model files; they need torch_geometric, which is
    not installable here nor on the box, so this arm runs the oracle PORT (same loops, same per-node edge scans,
    same per-(level, layer) scatter) with all host threads. Each step = one forward over a bounded sample of the
    workload batch, sized so that the whole run stays within a few minutes."""
    if rank != 0:
        return
    from dagnn_b200 import data as D
    use_all_host_threads()
    B = build_workload(wl, world)
    m = build_module(wl)
    p = {k: v.detach().clone() for k, v in m.state_dict().items()}
    ng = int(B.num_graphs)
    probe_n = min(8, ng)
    with torch.no_grad():
        sub = D.select_graphs(B, range(probe_n))
        oracle_forward(p, sub, wl)
        t0 = time.perf_counter()
        oracle_forward(p, sub, wl)
        t_probe = time.perf_counter() - t0
    budget = 150.0
    per_graph = t_probe / probe_n
    n_s = int(max(1, min(ng, budget / max(1, args.steps + args.warmup) / max(per_graph, 1e-6))))
    # cost per graph grows with batch size (the per-node edge scan is O(E)): re-probe once at the chosen size
    S = D.select_graphs(B, range(n_s))
    with torch.no_grad():
        for _ in range(args.warmup):
            oracle_forward(p, S, wl)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            oracle_forward(p, S, wl)
        dt = time.perf_counter() - t0
    val = n_s * args.steps / dt
    line = {"impl": "reference", "metric": "graphs/sec forward", "value": val, "unit": "graphs/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["desc"], "graphs_per_gpu": wl["graphs"], "global_batch": wl["graphs"] * world},
            "cpu_baseline": {"value": val, "unit": "graphs/s", "cores": int(torch.get_num_threads()), "host_cpus": os.cpu_count(),
                             "kind": "port (the reference's files need torch_geometric, which exists neither on the box nor in "
                                     "this image: oracle/dagnn_oracle.py, pinned against the reference on the PyG shim)",
                             "sample": "each step = one forward over the first %d graphs of the %d-graph "
                             "batch (oracle/dagnn_oracle.py; smaller batches make the reference's O(N*E) edge scan cheaper "
                             "per graph, i.e. this favours the reference)" % (n_s, ng)},
            "e2e": {"value": val, "unit": "graphs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ main
def gather_stats(x, world, dev):
    """[value of every rank] (one tiny all_gather; outside the timed regions)."""
    import torch.distributed as dist
    if world == 1:
        return [float(x)]
    t = torch.tensor([float(x)], device=dev, dtype=torch.float64)
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [float(o.item()) for o in out]


def main():
    args = parse()
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, wl, rank, world)
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("launch with: python -m torch.distributed.run --nnodes=1 --nproc-per-node %d --master-addr 127.0.0.1 "
                         "bench.py --gpus %d ..." % (args.gpus, args.gpus))
    import torch.distributed as dist
    from dagnn_b200 import _lib, data as D, runtime as rt, sharding

    _lib.build_library()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: dagnn_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    Bglobal = build_workload(wl, world)
    # depth-aware graph shards (data.shard_graph_ids): the deepest graphs go to different ranks and take fewer nodes with them
    Bcpu, my_graphs = sharding.shard_for_rank(Bglobal, rank, world)
    if Bcpu is None:
        raise SystemExit("rank %d owns no graph (more ranks than graphs)" % rank)
    n_graphs_local = int(Bcpu.num_graphs)
    m_cpu = build_module(wl)
    p_cpu = {k: v.detach().clone() for k, v in m_cpu.state_dict().items()}
    m = build_module(wl).to(dev)
    G = Bcpu.to(dev)
    Bpin = Bcpu.pin_memory()
    h2d_bytes = Bpin.nbytes()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)    # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        """`steps` calls of fn, each bracketed by CUDA events on the launching (current) stream, L2 flushed between
        calls outside the events; returns (sum of per-step device ms, wall s of the whole loop)."""
        for _ in range(warmup):
            fn()
        barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        t0 = time.perf_counter()
        for a, b in ev:
            if not args.no_flush:
                flush.fill_(1)
            a.record()
            fn()
            b.record()
        barrier()
        wall = time.perf_counter() - t0
        return sum(a.elapsed_time(b) for a, b in ev), wall

    def max_over_ranks(x):
        return max(gather_stats(x, world, dev))

    if args.train_step:
        return run_train_step(args, wl, m, G, Bcpu, rank, world, dev, timed, max_over_ranks, barrier)

    with torch.no_grad():
        # ---- parity gate before any timing is accepted: the WHOLE shard of rank 0 against the oracle
        if rank == 0:
            ref = oracle_forward(p_cpu, Bcpu, wl)
            got = hot_path(m, G, wl).cpu()
            err = (got - ref).abs().max().item()
            if not err <= 1e-4:
                raise SystemExit("parity gate failed: max-abs err %g vs the oracle on the %d-graph shard of rank 0" % (err, n_graphs_local))
            parity = {"checked_graphs": n_graphs_local, "max_abs_err_vs_oracle": err, "bar": 1e-4}

        clocks = ClockSampler(local)
        if rank == 0:
            clocks.start()
        # ---- value: inputs resident in HBM
        n0 = _lib.launch_count()
        ms_dev_local, _ = timed(lambda: hot_path(m, G, wl), args.steps, args.warmup)
        launches = (_lib.launch_count() - n0) // (args.steps + args.warmup) * args.steps
        ms_dev_all = gather_stats(ms_dev_local, world, dev)
        ms_dev = max(ms_dev_all)

        # ---- e2e: pinned host batch -> H2D -> forward -> D2H readout, per step
        out_host = None

        def e2e_step():
            nonlocal out_host
            Gd = Bpin.to(dev, non_blocking=True)
            out = hot_path(m, Gd, wl)
            if out_host is None:
                out_host = torch.empty(out.shape, dtype=out.dtype, pin_memory=True)
            out_host.copy_(out, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        ms_e2e, _ = timed(e2e_step, args.steps, max(3, args.warmup // 2))
        ms_e2e = max_over_ranks(ms_e2e)
        d2h_bytes = int(out_host.numel() * out_host.element_size())

        # ---- roofline of the dominant kernel: the sweep's launches alone (schedule + X prebuilt)
        if wl["kind"] == "code2":
            X, _, sched = m.node_states(G)
            packed, Din, nvid, use_ea = m._pack(dev), wl["emb"], 0, True
        else:
            X, _, sched = m.node_states(G)
            packed, Din, nvid, use_ea = m._packed, wl["emb"], (wl["emb"] if wl["kind"] == "NA" else 0), False
        n1 = _lib.launch_count()
        ms_sweep, _ = timed(lambda: rt.sweep(sched, X, packed, Din, wl["hid"], wl["layers"], nvid, use_ea), args.steps, 3)
        sweep_launches = (_lib.launch_count() - n1) // (args.steps + 3)
        clk = clocks.stop() if rank == 0 else None

    total_graphs = int(sum(gather_stats(n_graphs_local, world, dev)))
    value = total_graphs * args.steps / (ms_dev * 1e-3)
    e2e = total_graphs * args.steps / (ms_e2e * 1e-3)
    alg_bytes, e_prime = sweep_algorithmic_bytes(Bcpu, wl)
    peak, peak_src = measured_peaks()
    achieved = alg_bytes * args.steps / (ms_sweep * 1e-3) / 1e9
    H, layers, dirs = wl["hid"], wl["layers"], (2 if wl["bidir"] else 1)
    N = int(Bcpu.x.shape[0])
    flops = sum(6 * H * ((wl["emb"] if i == 0 else H) + H) for i in range(layers)) * N * dirs
    node_steps = N * dirs * layers
    cluster_path = H <= 256 and wl["emb"] <= 256 and node_steps <= 600000 and os.environ.get("DAGNN_SWEEP_PATH") != "grid"
    kernel = "k_sweep_cluster (cluster-resident level sweep)" if cluster_path else "k_sweep (grid-wide level sweep)"
    per_rank = {"nodes": [int(v) for v in gather_stats(N, world, dev)], "graphs": [int(v) for v in gather_stats(n_graphs_local, world, dev)],
                "levels": [int(v) for v in gather_stats(int(sched.num_levels[0]), world, dev)],
                "forward_ms": [v / args.steps for v in ms_dev_all],
                "sweep_ms": [v / args.steps for v in gather_stats(ms_sweep, world, dev)]}
    if rank == 0:
        traffic, traffic_src = ncu_traffic(args.workload, kernel.split(" ")[0])
        line = {
            "metric": "graphs/sec forward", "value": value, "unit": "graphs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["desc"], "graphs_per_gpu": wl["graphs"], "global_batch": total_graphs},
            "details": {"nodes_rank0": N, "edges_rank0": int(Bcpu.edge_index.shape[1]), "levels": int(sched.num_levels[0]),
                        "l2": "flushed between steps (256 MiB fill)" if not args.no_flush else "not flushed",
                        "timed_region": "node encoder + schedule build (incl. its one D2H of level offsets) + level sweeps + readout",
                        "parallelism": "graph-sharded dp%d (depth-aware shards), no forward collective" % world},
            "e2e": {"value": e2e, "unit": "graphs/s", "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": d2h_bytes,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches),
            "roofline": {"kernel": "%s, %d launch per forward" % (kernel, sweep_launches), "bound": "hbm",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "traffic_source": traffic_src, "peak_source": peak_src, "algorithmic_bytes_per_sweep": int(alg_bytes),
                         "sweep_ms": ms_sweep / args.steps, "gathered_edges": e_prime,
                         "gate_gemm_tflops_fp32": flops * args.steps / (ms_sweep * 1e-3) / 1e12},
            "parity": parity,
            "per_rank": per_rank,
            "clocks": clk,
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(p_cpu, Bcpu, wl)
        elif world > 1:
            line["cpu_baseline"] = None
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_train_step(args, wl, m, G, Bcpu, rank, world, dev, timed, max_over_ranks, barrier):
    """forward + loss + backward + one all-reduce of the flat gradient buffer + Adam step, per step (main_pyg.py:55-65 with the
    reference's DataParallel reduce-add, tg/data_parallel.py:59-62, replaced by NCCL)."""
    import torch.distributed as dist
    from dagnn_b200 import _lib, sharding
    if wl["kind"] != "code2":
        raise SystemExit("--train-step is defined for the code2-shaped workloads")
    m.train()
    params = [p for p in m.parameters()]
    opt = torch.optim.Adam(params, lr=1e-3)
    flat = sharding.FlatGradients(params)            # every .grad is a view into one buffer: the exchange is one NCCL call
    nb = int(Bcpu.num_graphs)
    g = torch.Generator().manual_seed(7 + rank)
    from dagnn_b200 import data as D
    y = torch.randint(0, D.CODE2_NUM_VOCAB, (D.CODE2_MAX_SEQ_LEN, nb), generator=g).to(dev)
    ev = {k: [] for k in ("fwd", "bwd", "wait", "ar", "opt")}
    tiny = torch.zeros(1, device=dev)
    nelem = {"n": 0}

    def step():
        e = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        e[0].record()
        flat.zero()
        pred = m(G)
        loss = sum(torch.nn.functional.cross_entropy(pred[k], y[k]) for k in range(len(pred))) / len(pred)
        e[1].record()
        loss.backward()
        e[2].record()
        if world > 1:
            dist.all_reduce(tiny)                   # absorbs the skew between ranks: the next interval is the exchange alone
        e[5].record()
        nelem["n"] = flat.allreduce(average=True)
        e[3].record()
        opt.step()
        e[4].record()
        ev["fwd"].append((e[0], e[1])); ev["bwd"].append((e[1], e[2])); ev["wait"].append((e[2], e[5])); ev["ar"].append((e[5], e[3]))
        ev["opt"].append((e[3], e[4]))

    clocks = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    if rank == 0:
        clocks.start()
    n0 = _lib.launch_count()
    ms, _ = timed(step, args.steps, args.warmup)
    launches = (_lib.launch_count() - n0) // (args.steps + args.warmup) * args.steps
    clk = clocks.stop() if rank == 0 else None
    ms = max_over_ranks(ms)
    parts = {k: sum(a.elapsed_time(b) for a, b in v[-args.steps:]) / args.steps for k, v in ev.items()}
    parts = {k: max_over_ranks(v) for k, v in parts.items()}
    total_graphs = int(sum(gather_stats(nb, world, dev)))
    grad_bytes = 4 * nelem["n"]
    busbw = (2.0 * (world - 1) / world) * grad_bytes / (parts["ar"] * 1e-3) / 1e9 if world > 1 else None
    if rank == 0:
        line = {"metric": "graphs/sec training step", "value": total_graphs * args.steps / (ms * 1e-3), "unit": "graphs/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": wl["desc"], "graphs_per_gpu": wl["graphs"], "global_batch": total_graphs,
                           "timed_region": "zero_grad + forward (all heads) + cross entropy + backward + NCCL all-reduce of the flat fp32 "
                                           "gradient buffer + Adam step",
                           "parallelism": "graph-sharded dp%d; one gradient all-reduce per step" % world},
                "gpu_launches": int(launches),
                "train": {"forward_ms": parts["fwd"], "backward_ms": parts["bwd"], "rank_skew_wait_ms": parts["wait"], "allreduce_ms": parts["ar"],
                          "optimizer_ms": parts["opt"],
                          "gradient_elements": nelem["n"], "gradient_bytes": grad_bytes, "allreduce_busbw_GBs": busbw,
                          "nvlink_pool_GBs": 725.0},
                "clocks": clk}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
