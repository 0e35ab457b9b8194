"""GPU, two or more devices on the box (skipped otherwise; run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`):
graph-sharded data parallelism over NCCL — the gathered shard readouts equal the unsharded forward, and the all-reduced shard
gradients equal the single-GPU gradients of the whole batch (replaces tg/data_parallel.py:59-62)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _model(dev):
    from dagnn_b200 import data as D, ogb
    enc = ogb.ASTNodeEncoder(64, D.CODE2_NUM_NODETYPES, D.CODE2_NUM_NODEATTRS, D.CODE2_MAX_DEPTH)
    m = ogb.DAGNN(50, 5, 64, 64, None, encoder=enc, num_layers=2, bidirectional=True, out_wx=False, out_pool_all=False)
    D.deterministic_init_(m, 4)
    return m.to(dev)


def _loss(pred, ws):
    return sum((p * w).sum() for p, w in zip(pred, ws))


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from dagnn_b200 import data as D, sharding
        B = D.make_code2_batch(24, 99)
        ids = D.shard_graph_ids(D.graph_node_counts(B), world, D.graph_depths(B))
        mine, my_ids = sharding.shard_for_rank(B)
        g = torch.Generator().manual_seed(5)
        ws_full = [torch.randn(24, 50, generator=g) for _ in range(5)]
        # ---- forward: gathered shard readouts == unsharded readout
        m = _model(dev)
        with torch.no_grad():
            ro = m.forward_readout(mine.to(dev))
            got = sharding.unshard_rows(sharding.gather_rows(ro, [len(i) for i in ids]), ids)
            full = m.forward_readout(B.to(dev))
        assert (got - full).abs().max().item() <= 1e-6
        # ---- training: all-reduced shard gradients == gradients of the whole batch on one GPU
        m.train()
        pred = m(mine.to(dev))
        sel = torch.as_tensor(my_ids)
        _loss(pred, [w[sel].to(dev) for w in ws_full]).backward()
        n = sharding.allreduce_gradients(list(m.parameters()))
        ref = _model(dev)
        ref.train()
        _loss(ref(B.to(dev)), [w.to(dev) for w in ws_full]).backward()
        worst = 0.0
        for (name, p), (_, q) in zip(m.named_parameters(), ref.named_parameters()):
            if q.grad is None:
                continue
            scale = q.grad.abs().max().item()
            err = (p.grad - q.grad).abs().max().item()
            worst = max(worst, err / max(scale, 1e-12)) if scale > 1e-6 else worst
            assert err <= 1e-5 * scale + 1e-6, (name, err, scale)
        out[rank] = (n, worst)
    finally:
        dist.destroy_process_group()


def test_sharded_forward_and_allreduced_gradients_nccl(built_lib):
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs on the box")
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert len(out) == world and all(v[0] > 0 for v in out.values())
