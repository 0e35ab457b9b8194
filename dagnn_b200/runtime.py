"""Host-side plumbing between torch tensors and the C ABI (include/dagnn_b200.h).

PyTorch is used here for device memory, the current stream and nothing else: every array is allocated as a
torch tensor and handed to libdagnn_sm100.so as a raw pointer. No arithmetic of the path happens in torch.
"""
from __future__ import annotations

import ctypes as C
import threading
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import (DagnnPackLayout, DagnnReadoutBlock, DagnnSchedule, DagnnSweepArgs, check, lib)

POOLS = {"max": 0, "mean": 1, "add": 2}
FILTER_ALL, FILTER_LVL0, FILTER_LAST, FILTER_FIRST = 0, 1, 2, 3


_RAW_STREAM = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream() -> int:
    """cudaStream_t of torch's current stream on the current device (the launching stream of every C-ABI call)."""
    if _RAW_STREAM is not None:        # same value, a fraction of the host time of current_stream().cuda_stream
        return _RAW_STREAM(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _req_cuda(t: torch.Tensor, name: str, dtype=None):
    if not t.is_cuda:
        raise _lib.DagnnError("%s must live on a CUDA device (dagnn_b200 has no CPU path); got %s" % (name, t.device))
    if dtype is not None and t.dtype != dtype:
        raise _lib.DagnnError("%s must be %s, got %s" % (name, dtype, t.dtype))
    return t.contiguous()


def _align(n: int, a: int = 64) -> int:
    return (n + a - 1) // a * a


def embed(x: torch.Tensor, depth: torch.Tensor, type_tab, attr_tab, depth_tab, max_depth: int) -> torch.Tensor:
    """ASTNodeEncoder.forward (ogbg-code/utils.py:26-28) -> fp32 [N, D]. When D % 4 == 0 the kernel also writes the
    operand image of X for the sweep's first projection; it rides on the returned tensor as `._dagnn_image`."""
    x = _req_cuda(x, "x", torch.int64)
    depth = _req_cuda(depth.view(-1), "node_depth", torch.int64)
    T, A, P = (_req_cuda(t.detach(), "embedding table", torch.float32) for t in (type_tab, attr_tab, depth_tab))
    N, D = x.shape[0], T.shape[1]
    X = torch.empty(N, D, device=x.device, dtype=torch.float32)
    img = None
    if D % 4 == 0:
        nb = int(lib().dagnn_operand_image_bytes(N, D))
        img = torch.empty(nb + 1024, device=x.device, dtype=torch.uint8)
        off = (-img.data_ptr()) % 1024
        img = img[off: off + nb]
    if P.shape[0] != int(max_depth) + 1:
        raise _lib.DagnnError("depth table has %d rows, max_depth + 1 = %d expected" % (P.shape[0], int(max_depth) + 1))
    check(lib().dagnn_embed_f32(_ptr(x), _ptr(depth), _ptr(T), _ptr(A), _ptr(P), int(max_depth), int(T.shape[0]), int(A.shape[0]),
                                N, D, _ptr(X), D, _ptr(img), _stream()), "dagnn_embed_f32")
    X._dagnn_image = img
    return X


def dag_levels(edge_index: torch.Tensor, num_nodes: int, max_passes: int = 257):
    """Longest-path levels of a batch of DAGs on the device: (levels on the edges as given, levels on the reversed edges),
    int64 [N] each — `_bi_layer_idx0`, `_bi_layer_idx1` of the reference (`top_sort` / `add_order_info_01`,
    src/utils_dag.py:8-35,39-52; the node-id rows `_bi_layer_index0/1` are arange(N)). OGB: pass the AST edges only.
    One D2H of the 4-int summary at the end; a batch deeper than max_passes - 1 levels is retried with more passes,
    an edge list that is still moving after N + 1 passes is not acyclic (ValueError, like the host version)."""
    edge_index = _req_cuda(edge_index, "edge_index", torch.int64)
    if edge_index.dim() != 2 or edge_index.shape[0] != 2:
        raise _lib.DagnnError("edge_index must be [2, E]")
    edge_index = edge_index.contiguous()
    N, E, dev = int(num_nodes), int(edge_index.shape[1]), edge_index.device
    passes = max(1, int(max_passes))
    while True:
        ws = torch.empty((int(lib().dagnn_levels_workspace_bytes(N, passes)) + 3) // 4, device=dev, dtype=torch.int32)
        lf = torch.empty(N, device=dev, dtype=torch.int64)
        lb = torch.empty(N, device=dev, dtype=torch.int64)
        summary = torch.empty(4, device=dev, dtype=torch.int32)
        check(lib().dagnn_levels_build(_ptr(edge_index), N, E, passes, _ptr(lf), _ptr(lb), _ptr(summary), _ptr(ws),
                                       ws.numel() * 4, _stream()), "dagnn_levels_build")
        status = int(summary[1].item())
        if status == 0:
            return lf, lb
        if status == 2:
            raise _lib.DagnnError("dag_levels: an edge endpoint is outside [0, %d)" % N)
        if passes > N + 1:
            raise ValueError("edge list is not acyclic")
        passes = min(passes * 8, N + 2)


def dvae_rows_to_tensor(rows, n: int) -> torch.Tensor:
    """Text rows [[t0], [t1, f10], [t2, f20, f21], ...] (dvae/data/*.txt after ast.literal_eval) -> int32 [B, n, n] on the host."""
    out = np.zeros((len(rows), n, n), dtype=np.int32)
    for g, row in enumerate(rows):
        if len(row) != n:
            raise _lib.DagnnError("row %d has %d variables, expected %d" % (g, len(row), n))
        for i, node in enumerate(row):
            out[g, i, : len(node)] = node
    return torch.from_numpy(out)


def dvae_batch_from_rows(rows_dev: torch.Tensor, kind: str, nvt: int):
    """D-VAE rows (int32 [B, n, n] on the device, `dvae_rows_to_tensor`) -> the collated batch `forward(G)` reads, built on the
    device: decode_ENAS_to_pygraph / decode_BN_to_pygraph (dvae/util.py:343-385 / :290-339) + dvae/batch.py:26-145.
    One D2H of the edge count at the end (edge_index is trimmed to it)."""
    from .data import DagBatch
    rows_dev = _req_cuda(rows_dev, "rows", torch.int32)
    B, n = int(rows_dev.shape[0]), int(rows_dev.shape[1])
    nn, dev = n + 2, rows_dev.device
    N, ecap = B * nn, B * (nn * (nn - 1) // 2)
    x = torch.empty(N, nvt, device=dev, dtype=torch.float32)
    ei = torch.empty(2, ecap, device=dev, dtype=torch.int64)
    bi = torch.empty(2, 2, N, device=dev, dtype=torch.int64)
    batch = torch.empty(N, device=dev, dtype=torch.int64)
    counts = torch.empty(4, device=dev, dtype=torch.int32)
    nws = int(lib().dagnn_dvae_rows_workspace_bytes(B))
    ws = torch.empty((nws + 3) // 4, device=dev, dtype=torch.int32)
    check(lib().dagnn_dvae_rows_build(_ptr(rows_dev), B, n, 0 if kind == "NA" else 1, int(nvt), _ptr(x), _ptr(ei), ecap, _ptr(bi), _ptr(batch),
                                      _ptr(counts), _ptr(ws), nws, _stream()), "dagnn_dvae_rows_build")
    c = counts.cpu()
    if int(c[1]) != 0:
        raise _lib.DagnnError("dvae rows: a node type is outside [0, %d)" % nvt)
    E = int(c[0])
    return DagBatch(x=x, edge_index=ei[:, :E].contiguous(), bi_layer_index=bi, batch=batch, num_graphs=B)


_LAYOUTS = {}      # (N, E, B, dirs, max_levels, has edge attributes) -> (offsets, total, workspace bytes) of a schedule buffer
_LOCK = threading.Lock()     # module-level caches (_LAYOUTS, _WS): forward may run in one host thread per GPU (tg/data_parallel.py:60-61)


class Schedule(object):
    """Level-sorted node order + in-edge CSR per direction, built on the device (schedule.cu).

    `build` only enqueues kernels; nothing in a forward waits for the host. `finalize()` brings the 8-int summary and
    the level offsets to the host (one small D2H + sync): the module calls it once at the END of a forward to check the
    status word, tests call it to inspect the arrays."""

    def __init__(self):
        self.c = DagnnSchedule()
        self.buf = None
        self.host_head = None
        self._final = False
        self._lvl_off_host: List[np.ndarray] = []
        self._num_levels: List[int] = []
        self._keep = []

    @staticmethod
    def build(edge_index: torch.Tensor, levels: Sequence[torch.Tensor], node_ids: Sequence[Optional[torch.Tensor]],
              edge_attr: Optional[torch.Tensor], batch: Optional[torch.Tensor], num_graphs: int,
              max_levels: int = 256) -> "Schedule":
        dirs = len(levels)
        dev = edge_index.device
        edge_index = _req_cuda(edge_index, "edge_index", torch.int64)
        levels = [_req_cuda(l, "level array", torch.int64) for l in levels]
        node_ids = [None if n is None else _req_cuda(n, "node id array", torch.int64) for n in node_ids]
        if edge_attr is not None:
            edge_attr = _req_cuda(edge_attr, "edge_attr", torch.float32)
            if edge_attr.dim() != 2 or edge_attr.shape[1] != 2:
                raise _lib.DagnnError("edge_attr must be [E, 2] (ogbg-code/utils2.py:45,68)")
        if batch is not None:
            batch = _req_cuda(batch, "batch", torch.int64)
        N, E, B = int(levels[0].shape[0]), int(edge_index.shape[1]), int(num_graphs)
        s = Schedule()
        ML = int(max_levels)
        head = 8 + dirs * (ML + 1)
        lkey = (N, E, B, dirs, ML, edge_attr is not None)
        with _LOCK:
            lay = _LAYOUTS.get(lkey)
        if lay is None:
            sizes = [("head", head)]
            for d in range(dirs):
                sizes += [("perm%d" % d, N), ("pos%d" % d, N), ("rowptr%d" % d, N + 1), ("col%d" % d, E), ("eid%d" % d, E)]
                if edge_attr is not None:
                    sizes.append(("eattr%d" % d, 2 * E))
            sizes.append(("gptr", B + 1))
            sizes.append(("gdepth", B))
            offs, tot = {}, 0
            for k, n in sizes:
                offs[k] = tot
                tot += _align(max(n, 1))
            lay = (offs, tot, int(lib().dagnn_schedule_workspace_bytes(N, E, ML)))
            with _LOCK:
                if len(_LAYOUTS) > 64:
                    _LAYOUTS.clear()
                _LAYOUTS[lkey] = lay
        offs, tot, ws_bytes = lay
        buf = torch.empty(tot + (ws_bytes + 3) // 4, device=dev, dtype=torch.int32)
        s.buf, s.head_len, s.max_levels = buf, head, ML
        # the C struct gets plain addresses (base + offset); the tensor views of the same ranges are only built when
        # somebody asks for them (tests, tools) — a forward never does, and 15 slices cost more host time than the launch
        s._offs, s._dims = offs, (N, E, B, dirs, ML, edge_attr is not None)
        base = buf.data_ptr()
        addr = lambda k: base + 4 * offs[k]
        c = s.c
        c.N, c.E, c.B, c.dirs, c.max_levels = N, E, B, dirs, ML
        c.summary = addr("head")
        for d in range(dirs):
            c.perm[d], c.pos[d] = addr("perm%d" % d), addr("pos%d" % d)
            c.lvl_off[d] = base + 4 * (offs["head"] + 8 + d * (ML + 1))
            c.rowptr[d], c.col[d], c.eid[d] = addr("rowptr%d" % d), addr("col%d" % d), addr("eid%d" % d)
            c.eattr[d] = addr("eattr%d" % d) if edge_attr is not None else None
        c.gptr = addr("gptr")
        c.gdepth = addr("gdepth")
        ws_ptr = base + 4 * tot
        check(lib().dagnn_schedule_build(_ptr(edge_index), _ptr(levels[0]), _ptr(levels[1]) if dirs == 2 else None,
                                         _ptr(node_ids[0]), _ptr(node_ids[1]) if dirs == 2 else None,
                                         _ptr(edge_attr), _ptr(batch), C.byref(c), ws_ptr, ws_bytes, _stream()),
              "dagnn_schedule_build")
        s._keep = [edge_index, levels, node_ids, edge_attr, batch]
        s.has_edge_attr = edge_attr is not None
        return s

    def finalize(self) -> "Schedule":
        """D2H of the summary + level offsets, then validate. status 1 (a level >= max_levels) raises
        `ScheduleOverflow` so the caller can rebuild with a larger table."""
        if self._final:
            return self
        host = torch.empty(self.head_len, dtype=torch.int32, pin_memory=True)
        host.copy_(self.buf[: self.head_len], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        hn = host.numpy()
        status = int(hn[2])
        if status == 1:
            raise ScheduleOverflow(self.max_levels)
        if status != 0:
            raise _lib.DagnnError("schedule build: status %d (2: node id / edge endpoint / batch vector out of range)" % status)
        dirs, ML = self.c.dirs, self.max_levels
        self.host_head = host
        self._num_levels = [int(hn[d]) for d in range(dirs)]
        self._lvl_off_host = [hn[8 + d * (ML + 1): 8 + (d + 1) * (ML + 1)] for d in range(dirs)]
        if self._num_levels[0] < 1:
            raise _lib.DagnnError("empty batch")
        if dirs == 2 and self._num_levels[1] > self._num_levels[0]:
            raise _lib.DagnnError("direction 1 has more levels (%d) than direction 0 (%d): level arrays are not the "
                                  "longest-path levels of one DAG" % (self._num_levels[1], self._num_levels[0]))
        self._final = True
        return self

    # ---- tensor views of the device arrays (built on demand)
    def _view(self, key: str, n: int) -> torch.Tensor:
        o = self._offs[key]
        return self.buf[o: o + n]

    @property
    def summary(self) -> torch.Tensor:
        return self._view("head", 8)

    @property
    def perm(self) -> List[torch.Tensor]:
        return [self._view("perm%d" % d, self._dims[0]) for d in range(self._dims[3])]

    @property
    def pos(self) -> List[torch.Tensor]:
        return [self._view("pos%d" % d, self._dims[0]) for d in range(self._dims[3])]

    @property
    def lvl_off(self) -> List[torch.Tensor]:
        ML, h = self._dims[4], self._offs["head"]
        return [self.buf[h + 8 + d * (ML + 1): h + 8 + (d + 1) * (ML + 1)] for d in range(self._dims[3])]

    @property
    def rowptr(self) -> List[torch.Tensor]:
        return [self._view("rowptr%d" % d, self._dims[0] + 1) for d in range(self._dims[3])]

    @property
    def col(self) -> List[torch.Tensor]:
        return [self._view("col%d" % d, self._dims[1]) for d in range(self._dims[3])]

    @property
    def eid(self) -> List[torch.Tensor]:
        return [self._view("eid%d" % d, self._dims[1]) for d in range(self._dims[3])]

    @property
    def eattr(self) -> List[Optional[torch.Tensor]]:
        E = self._dims[1]
        if not self._dims[5]:
            return [None] * self._dims[3]
        return [self._view("eattr%d" % d, 2 * E).view(torch.float32).view(E, 2) for d in range(self._dims[3])]

    @property
    def gptr(self) -> torch.Tensor:
        return self._view("gptr", self._dims[2] + 1)

    @property
    def num_levels(self) -> List[int]:
        return self.finalize()._num_levels

    @property
    def lvl_off_host(self) -> List[np.ndarray]:
        return self.finalize()._lvl_off_host

    # ---- views used by tests (bit-exact parity of the integer pre-pass) ----
    def level_nodes(self, d: int, l: int) -> torch.Tensor:
        a, b = int(self.lvl_off_host[d][l]), int(self.lvl_off_host[d][l + 1])
        return self.perm[d][a:b].long()

    def level_edges(self, d: int, l: int) -> torch.Tensor:
        a, b = int(self.lvl_off_host[d][l]), int(self.lvl_off_host[d][l + 1])
        rp = self.rowptr[d]
        e0, e1 = int(rp[a]), int(rp[b])
        return self.eid[d][e0:e1].long()


class ScheduleOverflow(_lib.DagnnError):
    def __init__(self, max_levels):
        super().__init__("a level index >= max_levels=%d" % max_levels)
        self.max_levels = max_levels


def run_checked(fn, max_levels: int = 256):
    """fn(max_levels) -> (result, schedule). Runs the whole forward asynchronously, then checks the schedule status
    once at the end (the single host<->device round trip of a forward; the reference syncs at dagnn.py:137 and once per
    node at :155). Rebuilds with a larger level table if the batch is deeper than max_levels."""
    while True:
        res, sched = fn(max_levels)
        try:
            sched.finalize()
            return res
        except ScheduleOverflow:
            if max_levels >= (1 << 22):
                raise
            max_levels *= 8


class PackedParams(object):
    """Packed GRU + attention parameters of every (direction, layer); re-packed when a parameter changes.

    The cache key is (storage pointer, tensor version) of every parameter: optimizer steps, `load_state_dict`, `.to()` and
    in-place ops on the parameter bump it. Writes through `param.data` (e.g. `p.data.copy_()`, hand-rolled EMA) do NOT bump
    the version — call `invalidate()` (the modules do it from `load_state_dict` / `train()` / `_apply`) after such edits."""

    def __init__(self):
        self.key = None
        self.blobs: List[List[torch.Tensor]] = []
        self.layouts: List[DagnnPackLayout] = []
        self._lock = threading.Lock()

    def invalidate(self):
        with self._lock:
            self.key = None

    # a module carrying this cache stays deep-copyable / picklable (copy.deepcopy(model), torch.save(model)): the copy starts empty
    def __deepcopy__(self, memo):
        return PackedParams()

    def __getstate__(self):
        return {}

    def __setstate__(self, state):
        self.__init__()

    @staticmethod
    def layout(Din: int, H: int, nvid: int, first: bool, last: bool) -> DagnnPackLayout:
        L = DagnnPackLayout()
        check(lib().dagnn_pack_layout(Din, H, nvid, int(first), int(last), C.byref(L)), "dagnn_pack_layout")
        return L

    def update(self, cells, aggrs, Din: int, H: int, nvid: int, use_edge_attr: bool, device):
        """cells[d][i]: nn.GRUCell; aggrs[d][i]: module with .attn_lin (and .edge_encoder when use_edge_attr)."""
        params = []
        for d in range(len(cells)):
            for i in range(len(cells[d])):
                cell, ag = cells[d][i], aggrs[d][i]
                params += [cell.weight_ih, cell.weight_hh, cell.bias_ih, cell.bias_hh, ag.attn_lin.weight]
                if use_edge_attr:
                    params.append(ag.edge_encoder.weight)
        key = tuple((p.data_ptr(), p._version) for p in params) + (Din, H, nvid, use_edge_attr, str(device))
        with self._lock:
            if key == self.key:
                return self
            self._repack(cells, aggrs, Din, H, nvid, use_edge_attr, device)
            self.key = key
        return self

    def _repack(self, cells, aggrs, Din: int, H: int, nvid: int, use_edge_attr: bool, device):
        blobs, layouts = [], []
        for d in range(len(cells)):
            row = []
            for i in range(len(cells[d])):
                cell, ag = cells[d][i], aggrs[d][i]
                din = Din if i == 0 else H
                nl = len(cells[d])
                # the reference's GRUCell raises on a width mismatch; the pack kernel would read out of bounds instead
                if tuple(cell.weight_ih.shape) != (3 * H, din) or tuple(cell.weight_hh.shape) != (3 * H, H) or \
                        tuple(cell.bias_ih.shape) != (3 * H,) or tuple(cell.bias_hh.shape) != (3 * H,):
                    raise _lib.DagnnError("GRU cell (%d, %d): weight_ih %s / weight_hh %s do not match input width %d, hidden %d"
                                          % (d, i, tuple(cell.weight_ih.shape), tuple(cell.weight_hh.shape), din, H))
                L = self.layout(din, H, nvid, i == 0, i + 1 == nl)
                nxt = cells[d][i + 1].weight_ih.detach().contiguous() if i + 1 < nl else None
                for p in (cell.weight_ih, cell.weight_hh, cell.bias_ih, cell.bias_hh, ag.attn_lin.weight):
                    _req_cuda(p, "parameter", torch.float32)
                aw = ag.attn_lin.weight.detach().contiguous()
                Dq = aw.shape[1] - H - nvid
                if Dq < 0:
                    raise _lib.DagnnError("attn_lin.weight has %d columns < H + nvid" % aw.shape[1])
                ew = ag.edge_encoder.weight.detach().contiguous() if use_edge_attr else None
                if ew is not None and tuple(ew.shape) != (H, 2):
                    raise _lib.DagnnError("edge_encoder.weight must be [H, 2]")
                blob = torch.empty(int(L.total_floats), device=device, dtype=torch.float32)
                check(lib().dagnn_pack_params_f32(_ptr(cell.weight_ih.detach().contiguous()), _ptr(cell.weight_hh.detach().contiguous()),
                                                  _ptr(cell.bias_ih.detach().contiguous()), _ptr(cell.bias_hh.detach().contiguous()),
                                                  _ptr(aw), int(Dq), _ptr(ew), _ptr(nxt), C.byref(L), _ptr(blob), _stream()),
                      "dagnn_pack_params_f32")
                row.append(blob)
                if d == 0:
                    layouts.append(L)
            blobs.append(row)
        self.blobs, self.layouts = blobs, layouts


_WS = {}


def _sweep_workspace(device, dirs, layers, Din, H, N, E, max_levels) -> torch.Tensor:
    need = int(lib().dagnn_sweep_workspace_bytes(dirs, layers, Din, H, N, E, max_levels))
    key = (device.type, device.index, _stream(), threading.get_ident())     # never shared between concurrent forwards
    with _LOCK:
        w = _WS.get(key)
    if w is None or w.numel() * 4 < need:
        w = torch.zeros((need + 3) // 4, device=device, dtype=torch.int32)
        with _LOCK:
            _WS[key] = w
    return w


def sweep(sched: Schedule, X: torch.Tensor, packed: PackedParams, Din: int, H: int, num_layers: int, nvid: int = 0,
          use_edge_attr: bool = False, trace_steps: int = 0):
    """The level sweep (one persistent kernel). Returns Hs fp32 [dirs, layers, N, ldh] in POSITION order
    (row p = node perm[d][p]). Asynchronous: level offsets / level count are read on the device."""
    X = _req_cuda(X, "X", torch.float32)
    dirs, N = sched.c.dirs, sched.c.N
    if X.dim() != 2 or X.shape[0] != N or X.shape[1] != Din:
        raise _lib.DagnnError("X must be [N=%d, Din=%d], got %s" % (N, Din, tuple(X.shape)))
    ldh = (H + 3) // 4 * 4
    Hs = torch.empty(dirs, num_layers, N, ldh, device=X.device, dtype=torch.float32)
    ws = _sweep_workspace(X.device, dirs, num_layers, Din, H, N, sched.c.E, sched.c.max_levels)
    a = DagnnSweepArgs()
    a.sched = C.pointer(sched.c)
    hs0, hs_stride = Hs.data_ptr(), N * ldh * 4
    for d in range(dirs):
        for i in range(num_layers):
            a.Hs[d][i] = hs0 + (d * num_layers + i) * hs_stride
            a.packed[d][i] = packed.blobs[d][i].data_ptr()
    a.num_layers = num_layers
    a.Din, a.H, a.nvid = Din, H, nvid
    a.X, a.ldx, a.ldh = X.data_ptr(), X.stride(0), ldh
    ximg = getattr(X, "_dagnn_image", None)
    a.X_image = ximg.data_ptr() if ximg is not None else None
    a.use_edge_attr = 1 if use_edge_attr else 0
    a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel() * 4
    trace = None
    if trace_steps > 0:     # profiling: per (step, CTA) clock64 stamps, see include/dagnn_b200.h
        trace = torch.zeros(int(lib().dagnn_sweep_trace_bytes(trace_steps)) // 8, device=X.device, dtype=torch.int64)
        a.trace = trace.data_ptr()
    check(lib().dagnn_sweep_forward_f32(C.byref(a), _stream()), "dagnn_sweep_forward_f32")
    if trace is not None:
        return Hs, trace.view(-1, 256, 16)
    return Hs


def readout(sched: Schedule, blocks: Sequence[dict], pool: str, out_width: int, device) -> torch.Tensor:
    """blocks: dicts with src (2-D fp32 tensor), width, index_mode, dir, filter, filter_lvl, out_col."""
    arr = (DagnnReadoutBlock * len(blocks))()
    keep = []
    for k, b in enumerate(blocks):
        src = b["src"]
        arr[k].src, arr[k].ld, arr[k].width = src.data_ptr(), src.stride(0), int(b["width"])
        arr[k].index_mode, arr[k].dir, arr[k].filter = int(b.get("index_mode", 0)), int(b.get("dir", 0)), int(b.get("filter", 0))
        fl = b.get("filter_lvl")
        if fl is not None:
            fl = _req_cuda(fl, "filter_lvl", torch.int64)
            keep.append(fl)
        arr[k].filter_lvl = _ptr(fl)
        arr[k].out_col = int(b["out_col"])
    out = torch.empty(sched.c.B, out_width, device=device, dtype=torch.float32)
    check(lib().dagnn_readout_f32(C.byref(sched.c), arr, len(blocks), POOLS[pool], out.data_ptr(), out.stride(0), _stream()),
          "dagnn_readout_f32")
    return out


def states_to_node_order(sched: Schedule, Hs: torch.Tensor, H: int) -> List[List[torch.Tensor]]:
    """Hs [dirs, layers, N, ldh] (position order) -> G.h-style nested list of [N, H] tensors in node order."""
    out = []
    for d in range(Hs.shape[0]):
        row = []
        for i in range(Hs.shape[1]):
            dst = torch.empty(Hs.shape[2], H, device=Hs.device, dtype=torch.float32)
            check(lib().dagnn_states_to_node_order_f32(C.byref(sched.c), d, Hs[d, i].data_ptr(), Hs.stride(2), H, dst.data_ptr(),
                                                       H, _stream()), "dagnn_states_to_node_order_f32")
            row.append(dst)
        out.append(row)
    return out
