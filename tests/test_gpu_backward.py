"""GPU: backward of the path against torch.autograd on the CPU oracle — every parameter gradient and
the gradients of the embedding tables, on code2-shaped and random DAG batches, OGB and D-VAE flavours; the tensor-core
GEMM behind the dense layers in all operand layouts; one optimizer step end to end (main_pyg.py:55-65 semantics)."""
import numpy as np
import pytest
import torch

from helpers import state_dict_cpu

pytestmark = pytest.mark.gpu

RTOL = 1e-4     # gradients: max-abs error relative to the largest entry of the reference gradient


@pytest.fixture(scope="module")
def dev(built_lib):
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


@pytest.mark.parametrize("M,N,K,ak,bk,acc", [(300, 200, 192, 1, 1, 0), (77, 130, 100, 1, 0, 1), (256, 64, 1000, 0, 0, 0), (5, 7, 3, 1, 1, 0),
                                             (129, 257, 65, 0, 1, 1)])
def test_gemm_all_layouts_match_fp64(built_lib, dev, M, N, K, ak, bk, acc):
    from dagnn_b200 import _lib
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g)
    B = torch.randn(N, K, generator=g)
    C0 = torch.randn(M, N, generator=g)
    Ad = (A if ak else A.t().contiguous()).to(dev)
    Bd = (B if bk else B.t().contiguous()).to(dev)
    Cd = C0.clone().to(dev)
    _lib.check(built_lib.dagnn_gemm_f32(Ad.data_ptr(), Ad.stride(0), ak, Bd.data_ptr(), Bd.stride(0), bk, Cd.data_ptr(), N, M, N, K, acc,
                                        torch.cuda.current_stream().cuda_stream), "dagnn_gemm_f32")
    ref = A.double() @ B.double().t() + (C0.double() if acc else 0)
    err = (Cd.cpu().double() - ref).abs().max().item()
    assert err <= 1e-5 * ref.abs().max().item() + 1e-5, err


def test_linear_fn_forward_backward(dev):
    from dagnn_b200 import autograd as ag
    torch.manual_seed(0)
    lin = torch.nn.Linear(300, 77).to(dev)
    x = torch.randn(50, 300, device=dev, requires_grad=True)
    w = torch.randn(50, 77, device=dev)
    y = ag.linear(x, lin)
    (y * w).sum().backward()
    xr = x.detach().double().cpu().requires_grad_(True)
    W, b = lin.weight.detach().double().cpu().requires_grad_(True), lin.bias.detach().double().cpu().requires_grad_(True)
    yr = xr @ W.t() + b
    (yr * w.double().cpu()).sum().backward()
    assert (y.detach().cpu().double() - yr.detach()).abs().max() <= 1e-5 * yr.abs().max()
    for got, ref in ((x.grad, xr.grad), (lin.weight.grad, W.grad), (lin.bias.grad, b.grad)):
        assert (got.cpu().double() - ref).abs().max() <= 1e-5 * ref.abs().max() + 1e-6


def _grad_check(named_got, named_ref, tag):
    worst = 0.0
    top = max(float(r.abs().max()) for r in named_ref.values())
    for name, ref in named_ref.items():
        got = named_got[name]
        assert got is not None, "%s: no gradient for %s" % (tag, name)
        scale = ref.abs().max().item()
        err = (got.detach().cpu() - ref).abs().max().item()
        if scale < 1e-6 * top:           # parameters whose exact gradient is zero (autograd leaves rounding noise there)
            assert err <= 1e-5 * top, "%s %s: err %g on a zero gradient" % (tag, name, err)
            continue
        worst = max(worst, err / scale)
        assert err <= RTOL * scale + 1e-7, "%s %s: max-abs err %g, reference scale %g" % (tag, name, err, scale)
    return worst


OGB_CASES = [
    # graphs, seed, emb, hid, layers, bidir, kind, kwargs
    (6, 11, 32, 32, 2, True, "code2", {}),
    (5, 12, 24, 40, 3, True, "rand", {}),
    (7, 13, 48, 36, 1, False, "rand", {}),
    (4, 14, 64, 64, 2, True, "code2", dict(out_wx=True, out_pool="add")),
    (6, 15, 20, 28, 2, True, "rand", dict(w_edge_attr=False, out_pool_all=True, out_pool="mean")),
    (3, 16, 256, 256, 2, True, "code2", {}),
    (5, 17, 32, 48, 2, True, "code2", dict(agg="self_attn_h")),
]


@pytest.mark.parametrize("ng,seed,emb,hid,layers,bidir,kind,kw", OGB_CASES)
def test_ogb_gradients_match_oracle_autograd(ng, seed, emb, hid, layers, bidir, kind, kw, dev):
    from dagnn_b200 import data as D, ogb
    from oracle import dagnn_oracle as O
    B = D.make_code2_batch(ng, seed) if kind == "code2" else D.make_random_dag_batch(ng, seed, n_hi=30, with_attr=kw.get("w_edge_attr", True))
    enc = ogb.ASTNodeEncoder(emb, D.CODE2_NUM_NODETYPES, D.CODE2_NUM_NODEATTRS, D.CODE2_MAX_DEPTH)
    args = dict(num_layers=layers, bidirectional=bidir, out_wx=False, out_pool_all=False)
    args.update(kw)
    m = ogb.DAGNN(50, 5, emb, hid, None, encoder=enc, **args)
    D.deterministic_init_(m, seed)
    m.train()
    # ---- reference: autograd through the CPU oracle
    p = {k: v.clone().requires_grad_(True) for k, v in state_dict_cpu(m).items()}
    preds, _, _ = O.ogb_forward(p, B, num_layers=layers, bidirectional=bidir, out_wx=args["out_wx"], out_pool_all=args["out_pool_all"],
                                out_pool=args.get("out_pool", "max"), w_edge_attr=args.get("w_edge_attr", True), agg=args.get("agg", "attn_h"))
    g = torch.Generator().manual_seed(seed)
    ws = [torch.randn(preds[0].shape, generator=g) for _ in preds]
    sum((pr * w).sum() for pr, w in zip(preds, ws)).backward()
    ref = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in p.items()}
    # ---- this package
    m = m.to(dev)
    out = m(B.to(dev))
    sum((pr * w.to(dev)).sum() for pr, w in zip(out, ws)).backward()
    got = {k: v.grad for k, v in m.named_parameters()}
    if not bidir:      # the reference model keeps unused aggregators of direction 1 (no gradient on either side)
        ref = {k: v for k, v in ref.items() if not k.startswith("node_aggr_1")}
    _grad_check(got, {k: v for k, v in ref.items() if k in got and got[k] is not None or v.abs().max() > 0}, "ogb")


@pytest.mark.parametrize("kind,hs,layers,bidir", [("NA", 40, 2, False), ("NA", 32, 2, True), ("BN", 48, 2, True), ("BN", 36, 3, False)])
def test_dvae_gradients_match_oracle_autograd(kind, hs, layers, bidir, dev):
    from dagnn_b200 import data as D, dvae
    from oracle import dagnn_oracle as O
    nvt = 8 if kind == "NA" else 10
    B = D.make_random_dvae_batch(12, 3 + hs, kind)
    cls = dvae.DAGNN if kind == "NA" else dvae.DAGNN_BN
    m = cls(nvt, hs, hs, nvt, nvt, 0, 1, hs=hs, nz=16, num_nodes=nvt, num_layers=layers, bidirectional=bidir)
    D.deterministic_init_(m, hs)
    m.train()
    p = {k: v.clone().requires_grad_(True) for k, v in state_dict_cpu(m).items()}
    mu_r, lv_r = O.dvae_encode(p, B, num_layers=layers, bidirectional=bidir, num_nodes=nvt, vid=(kind == "NA"))
    g = torch.Generator().manual_seed(hs)
    w1, w2 = torch.randn(mu_r.shape, generator=g), torch.randn(lv_r.shape, generator=g)
    ((mu_r * w1).sum() + (lv_r * w2).sum()).backward()
    ref = {k: v.grad for k, v in p.items() if v.grad is not None}
    m = m.to(dev)
    mu, lv = m.encode([B.to(dev)])
    ((mu * w1.to(dev)).sum() + (lv * w2.to(dev)).sum()).backward()
    got = {k: v.grad for k, v in m.named_parameters()}
    # cells_0 / cells_1 alias grue_forward / grue_backward: the oracle sees the grue_* names only
    ref = {k: v for k, v in ref.items() if k in got}
    assert any(k.startswith("grue_forward") for k in ref) and any(k.startswith("node_aggr_0") for k in ref)
    _grad_check(got, ref, "dvae")


def test_one_training_step_like_main_pyg(dev):
    """main_pyg.py:55-65: forward, multi-head cross entropy, backward, clip, optimizer step — and the next forward sees the
    updated parameters (the packed-parameter cache follows the tensors' version counters)."""
    from dagnn_b200 import data as D, ogb
    B = D.make_code2_batch(8, 5)
    enc = ogb.ASTNodeEncoder(64, D.CODE2_NUM_NODETYPES, D.CODE2_NUM_NODEATTRS, D.CODE2_MAX_DEPTH)
    m = ogb.DAGNN(50, 5, 64, 64, None, encoder=enc, out_wx=False, out_pool_all=False).to(dev)
    D.deterministic_init_(m, 2)
    m.train()
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    G = B.to(dev)
    y = torch.randint(0, 50, (5, 8), generator=torch.Generator().manual_seed(0)).to(dev)
    losses = []
    for _ in range(3):
        opt.zero_grad()
        pred = m(G)
        loss = sum(torch.nn.functional.cross_entropy(pred[k].float(), y[k]) for k in range(5)) / 5
        loss.backward()
        torch.nn.utils.clip_grad_norm_(m.parameters(), 0.25)
        opt.step()
        losses.append(loss.item())
    assert all(np.isfinite(losses)) and losses[2] < losses[0]
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for n, p in m.named_parameters() if not n.startswith("node_aggr_1") or True)
