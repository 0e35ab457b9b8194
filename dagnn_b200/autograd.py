"""Autograd bridges over the C ABI: `loss.backward()` on the modules of this package.

The reference trains through torch.autograd over its Python level loop (main_pyg.py:55-65, dvae/train.py:255-264). Here the
forward is one library call per stage, so the backward is too: `EmbedFn` (node encoder, ogbg-code/utils.py:26-28),
`SweepReadoutFn` (schedule + level sweep + pooled readout, ogbg-code/model/dagnn.py:141-202) and `LinearFn` (the small dense
layers, dagnn.py:209-215) are `torch.autograd.Function`s whose backward calls `dagnn_embed_backward_f32`,
`dagnn_readout_backward_f32` + `dagnn_sweep_backward_f32` and `dagnn_gemm_f32`. No arithmetic of the path runs in torch.
"""
from __future__ import annotations

import ctypes as C
from typing import List

import torch

from . import _lib, runtime as rt
from ._lib import DagnnReadoutBlock, DagnnSweepBwdArgs, check, lib


class EmbedFn(torch.autograd.Function):
    """X = T[x0] + A[x1] + P[min(depth, max_depth)]; backward scatters dX into the three tables."""

    @staticmethod
    def forward(ctx, x, depth, T, A, P, max_depth):
        X = rt.embed(x, depth, T, A, P, max_depth)
        ctx.save_for_backward(x, depth)
        ctx.shapes = (T.shape, A.shape, P.shape, int(max_depth))
        return X

    @staticmethod
    def backward(ctx, dX):
        x, depth = ctx.saved_tensors
        ts, as_, ps, max_depth = ctx.shapes
        dX = dX.contiguous()
        dev = dX.device
        dT = torch.zeros(ts, device=dev, dtype=torch.float32)
        dA = torch.zeros(as_, device=dev, dtype=torch.float32)
        dP = torch.zeros(ps, device=dev, dtype=torch.float32)
        xx = x.contiguous()
        dd = depth.view(-1).contiguous()
        check(lib().dagnn_embed_backward_f32(xx.data_ptr(), dd.data_ptr(), max_depth, int(ts[0]), int(as_[0]), int(xx.shape[0]), int(ts[1]),
                                             dX.data_ptr(), dX.stride(0), dT.data_ptr(), dA.data_ptr(), dP.data_ptr(), rt._stream()),
              "dagnn_embed_backward_f32")
        return None, None, dT, dA, dP, None


class LinearFn(torch.autograd.Function):
    """y = x w^T + b on the tensor cores (fp16 x 3 split) — nn.Linear without cuBLAS, forward and backward."""

    @staticmethod
    def forward(ctx, x, w, b):
        x2 = x.reshape(-1, x.shape[-1]).contiguous()
        w = w.contiguous()
        y = torch.empty(x2.shape[0], w.shape[0], device=x.device, dtype=torch.float32)
        if x2.shape[0] > 0:
            check(lib().dagnn_linear_f32(x2.data_ptr(), x2.stride(0), w.data_ptr(), w.stride(0), None if b is None else b.contiguous().data_ptr(),
                                         y.data_ptr(), y.stride(0), x2.shape[0], w.shape[0], w.shape[1], rt._stream()), "dagnn_linear_f32")
        ctx.save_for_backward(x2, w)
        ctx.has_bias = b is not None
        ctx.xshape = x.shape
        return y.reshape(*x.shape[:-1], w.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, w = ctx.saved_tensors
        dy2 = dy.reshape(-1, dy.shape[-1]).contiguous()
        M, N, K = x2.shape[0], w.shape[0], w.shape[1]
        dx = torch.zeros(M, K, device=dy.device, dtype=torch.float32)
        dw = torch.zeros(N, K, device=dy.device, dtype=torch.float32)
        if M > 0:
            st = rt._stream()
            # dx[M, K] = dy[M, N] w[N, K]: contraction over N — w read as the [K' = N, rows = K] view
            check(lib().dagnn_gemm_f32(dy2.data_ptr(), dy2.stride(0), 1, w.data_ptr(), w.stride(0), 0, dx.data_ptr(), K, M, K, N, 0, st), "dagnn_gemm_f32")
            # dw[N, K] = dy^T[N, M] x[M, K]: contraction over M — both operands as transposed views
            check(lib().dagnn_gemm_f32(dy2.data_ptr(), dy2.stride(0), 0, x2.data_ptr(), x2.stride(0), 0, dw.data_ptr(), K, N, K, M, 0, st), "dagnn_gemm_f32")
        db = dy2.sum(0) if ctx.has_bias else None
        return dx.reshape(ctx.xshape), dw, db


def linear(x: torch.Tensor, layer: torch.nn.Linear) -> torch.Tensor:
    """nn.Linear through `dagnn_linear_f32` (autograd-aware). CPU tensors raise: there is no eager fallback."""
    rt._req_cuda(x, "input of a dense layer", torch.float32)
    return LinearFn.apply(x, layer.weight, layer.bias)


class SweepReadoutFn(torch.autograd.Function):
    """schedule -> level sweep -> pooled readout, differentiable in X and in every cell / aggregator parameter.

    `host` is the module (duck-typed): it provides `build_schedule(G, max_levels)`, `_pack(device)`, `_sweep_dims()` ->
    (Din, H, layers, nvid, use_edge_attr), `_readout_blocks(G, X, Hs)` -> (blocks, pool, width) and `_cell_params()` -> the
    parameter list in the order `params` was built from (per direction, per layer: weight_ih, weight_hh, bias_ih, bias_hh,
    attn_lin.weight, attn_lin.bias[, edge_encoder.weight, edge_encoder.bias])."""

    @staticmethod
    def forward(ctx, host, G, X, *params):
        return _sweep_forward(ctx, host, G, X, *params)

    @staticmethod
    def backward(ctx, dout):
        return _sweep_backward(ctx, dout)


def _sweep_forward(ctx, host, G, X, *params):
    Din, H, layers, nvid, use_ea = host._sweep_dims()
    keep = {}

    def run(max_levels):
        sched = host.build_schedule(G, max_levels)
        packed = host._pack(X.device)
        Hs = rt.sweep(sched, X, packed, Din, H, layers, nvid, use_ea)
        blocks, pool, width = host._readout_blocks(G, X, Hs)
        keep["v"] = (Hs, blocks, pool, sched)
        return rt.readout(sched, blocks, pool, width, X.device), sched
    out = rt.run_checked(run)
    Hs, blocks, pool, sched = keep["v"]
    ctx.host, ctx.G, ctx.sched = host, G, sched
    ctx.dims = (Din, H, layers, nvid, use_ea)
    ctx.pool = pool
    ctx.block_specs = [{k: v for k, v in b.items() if k != "src"} for b in blocks]
    ctx.block_src = [("X", None) if b["src"].data_ptr() == X.data_ptr() else ("H", (int(b["dir"]), _layer_of(b, Hs))) for b in blocks]
    ctx.save_for_backward(X, Hs, out, *params)
    return out


def _layer_of(block, Hs) -> int:
    src, d = block["src"], int(block.get("dir", 0))
    for i in range(Hs.shape[1]):
        if src.data_ptr() == Hs[d, i].data_ptr():
            return i
    raise _lib.DagnnError("readout block source is neither X nor a state tensor")


def _sweep_backward(ctx, dout):
    X, Hs, out = ctx.saved_tensors[:3]
    params = ctx.saved_tensors[3:]
    host, G, sched = ctx.host, ctx.G, ctx.sched
    Din, H, layers, nvid, use_ea = ctx.dims
    dirs, N, ldh = Hs.shape[0], Hs.shape[2], Hs.shape[3]
    dev = X.device
    dout = dout.contiguous()
    dHs = torch.zeros_like(Hs)
    dX = torch.zeros(N, Din, device=dev, dtype=torch.float32)
    st = rt._stream()
    # ---- readout backward into dHs / dX
    nb = len(ctx.block_specs)
    arr = (DagnnReadoutBlock * nb)()
    fwd = (C.c_void_p * nb)()
    keep = []
    for k, (spec, (kind, where)) in enumerate(zip(ctx.block_specs, ctx.block_src)):
        if kind == "X":
            gsrc, fsrc = dX, X
        else:
            gsrc, fsrc = dHs[where[0], where[1]], Hs[where[0], where[1]]
        arr[k].src, arr[k].ld, arr[k].width = gsrc.data_ptr(), gsrc.stride(0), int(spec["width"])
        arr[k].index_mode, arr[k].dir, arr[k].filter = int(spec.get("index_mode", 0)), int(spec.get("dir", 0)), int(spec.get("filter", 0))
        fl = spec.get("filter_lvl")
        if fl is not None:
            fl = rt._req_cuda(fl, "filter_lvl", torch.int64)
            keep.append(fl)
        arr[k].filter_lvl = None if fl is None else fl.data_ptr()
        arr[k].out_col = int(spec["out_col"])
        fwd[k] = fsrc.data_ptr()
    check(lib().dagnn_readout_backward_f32(C.byref(sched.c), arr, fwd, nb, rt.POOLS[ctx.pool], out.data_ptr(), dout.data_ptr(), dout.stride(0), st),
          "dagnn_readout_backward_f32")
    # ---- sweep backward
    a = DagnnSweepBwdArgs()
    a.sched = C.pointer(sched.c)
    a.num_layers, a.Din, a.H, a.nvid = layers, Din, H, nvid
    a.X, a.ldx = X.data_ptr(), X.stride(0)
    a.ldh = ldh
    per = 8 if use_ea else 6
    grads: List[torch.Tensor] = []
    for d in range(dirs):
        for i in range(layers):
            w_ih, w_hh, b_ih, b_hh, aw, ab = params[(d * layers + i) * per: (d * layers + i) * per + 6]
            ew = params[(d * layers + i) * per + 6] if use_ea else None
            eb = params[(d * layers + i) * per + 7] if use_ea else None
            g = [torch.empty_like(w_ih), torch.empty_like(w_hh), torch.empty_like(b_ih), torch.empty_like(b_hh), torch.zeros_like(aw),
                 torch.zeros_like(ab)]
            if use_ea:
                g += [torch.zeros_like(ew), torch.zeros_like(eb)]
            a.Hs[d][i], a.dHs[d][i] = Hs[d, i].data_ptr(), dHs[d, i].data_ptr()
            p, q = a.params[d][i], a.grads[d][i]
            p.weight_ih, p.weight_hh, p.bias_ih, p.bias_hh = (t.contiguous().data_ptr() for t in (w_ih, w_hh, b_ih, b_hh))
            p.attn_w, p.Dq = aw.contiguous().data_ptr(), int(aw.shape[1] - H - nvid)
            p.edge_w = ew.contiguous().data_ptr() if use_ea else None
            q.weight_ih, q.weight_hh, q.bias_ih, q.bias_hh, q.attn_w = (t.data_ptr() for t in g[:5])
            q.edge_w = g[6].data_ptr() if use_ea else None
            grads += g
            keep += [w_ih, w_hh, b_ih, b_hh, aw, ew]
    a.dX, a.lddx = dX.data_ptr(), dX.stride(0)
    a.use_edge_attr = 1 if use_ea else 0
    a.num_levels = int(sched.num_levels[0])
    host_off = [torch.from_numpy(sched.lvl_off_host[d].copy()).contiguous() for d in range(dirs)]
    for d in range(dirs):
        a.lvl_off_host[d] = host_off[d].data_ptr()
    need = int(lib().dagnn_sweep_backward_workspace_bytes(Din, H, nvid, N, sched.c.E))
    ws = torch.empty((need + 3) // 4 + 64, device=dev, dtype=torch.int32)
    off = (-ws.data_ptr()) % 256
    a.workspace, a.workspace_bytes = ws.data_ptr() + off, need
    check(lib().dagnn_sweep_backward_f32(C.byref(a), st), "dagnn_sweep_backward_f32")
    return (None, None, dX if ctx.needs_input_grad[2] else None) + tuple(grads)

