"""Graph operators with the reference's `ntsGraphOp` interface (core/ntsBaseOp.hpp:24-48):
constructed from `(PartitionedGraph, active)`, `forward(x)` / `forward(x, w)`, `backward(grad)`,
`get_additional_grad()`; outputs are freshly allocated zero tensors the kernels accumulate into
(NtsScheduler::NewKeyTensor / NewLeafTensor, core/NtsScheduler.hpp:378-394).

Every operator calls the sm_100a kernels through the C ABI (`_lib.call`); tensors only provide device
memory and the current CUDA stream.  There is no CPU path: a CPU tensor raises.
"""
from __future__ import annotations

import torch

from . import _lib


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _check_input(t, name="input"):
    if not t.is_cuda:
        raise _lib.NtsError("%s must be a CUDA tensor (libnts_b200 has no CPU fallback)" % name)
    if t.dtype != torch.float32 or t.dim() != 2:
        raise _lib.NtsError("%s must be a 2-D float32 tensor" % name)
    if not t.is_contiguous():
        # the reference borrows packed_accessor storage (core/NtsScheduler.hpp:505-515): contiguous only
        raise _lib.NtsError("%s must be contiguous" % name)
    return t


def _ptr(t):
    return 0 if t is None else t.data_ptr()


class KernelTimer:
    """Optional CUDA-event bracket around every aggregation launch (bench.py's live roofline measurement).
    Events are recorded on the stream the kernel is launched on; nothing is synchronised until `summary()`."""

    def __init__(self):
        self.records = []  # (tag, F, edges, rows, start, stop)

    def bracket(self, tag, F, edges, rows):
        start = torch.cuda.Event(enable_timing=True)
        stop = torch.cuda.Event(enable_timing=True)
        self.records.append((tag, int(F), int(edges), int(rows), start, stop))
        return start, stop

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for tag, F, edges, rows, a, b in self.records:
            d = out.setdefault((tag, F), {"calls": 0, "ms": 0.0, "edges": 0, "rows": 0})
            d["calls"] += 1
            d["ms"] += a.elapsed_time(b)
            d["edges"] += edges
            d["rows"] += rows
        return out


_timer = None


def set_kernel_timer(timer):
    global _timer
    _timer = timer


class _timed:
    """`with _timed(tag, F, edges, rows): launch` - CUDA events around the launch when a KernelTimer is installed."""

    def __init__(self, tag, F, edges, rows):
        self.ev = _timer.bracket(tag, F, edges, rows) if _timer else None

    def __enter__(self):
        if self.ev:
            self.ev[0].record()

    def __exit__(self, *exc):
        if self.ev:
            self.ev[1].record()
        return False


def segment_gather_sum(out, x, weight, indices, offsets, index_base, n_rows, n_edges):
    """out[r,:] += sum_e x[indices[e]-index_base,:] * weight[e]  (nts_segment_gather_sum)."""
    _lib.call("nts_segment_gather_sum", _ptr(x), _ptr(out), _ptr(weight), _ptr(indices), _ptr(offsets),
              int(index_base), int(n_rows), int(n_edges), int(x.shape[1]), _stream())
    return out


class GatherPlan:
    """nts_gather_plan (include/nts_b200.h): one chunk direction preprocessed once for repeated aggregation -
    source-slab bucketing (L2 residency), interleaved (row, weight) pairs, 16-byte aligned gathers."""

    def __init__(self, offsets, indices, weight, index_base, n_rows, n_edges, gather_rows, slabs, slot_of=None,
                 tune_for=0):
        """slabs > 0: that many source slabs; slabs == 0: the slab count is MEASURED for feature width `tune_for`
        (nts_gather_plan_create_tuned)."""
        L = _lib.load()
        if slabs > 0:
            self.handle = L.nts_gather_plan_create(_ptr(offsets), _ptr(indices), _ptr(weight), _ptr(slot_of),
                                                   int(index_base), int(n_rows), int(n_edges), int(gather_rows),
                                                   int(slabs), _stream())
        else:
            self.handle = L.nts_gather_plan_create_tuned(_ptr(offsets), _ptr(indices), _ptr(weight), _ptr(slot_of),
                                                         int(index_base), int(n_rows), int(n_edges),
                                                         int(gather_rows), int(tune_for), _stream())
        if not self.handle:
            raise _lib.NtsError("nts_gather_plan_create failed: " + L.nts_last_error().decode(errors="replace"))
        self.slabs = int(L.nts_gather_plan_slabs(self.handle))
        self.n_rows, self.n_edges = int(n_rows), int(n_edges)

    def run(self, x, out):
        _lib.call("nts_gather_plan_run", self.handle, _ptr(x), _ptr(out), int(x.shape[1]), _stream())
        return out

    def bytes(self):
        return int(_lib.load().nts_gather_plan_bytes(self.handle))

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _lib.load().nts_gather_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


# Plan policy: "auto" preprocesses chunks with at least PLAN_MIN_EDGES edges on first use (the arrays of a chunk are
# immutable, like the reference's CopyGraphToDevice uploads); "off" always takes the plain kernel on the reference
# layout; "on" plans every chunk.  NTS_PLAN / NTS_PLAN_SLABS are measurement overrides.
import os as _os

PLAN_MIN_EDGES = 1 << 20
_plan_mode = {"0": "off", "1": "on"}.get(_os.environ.get("NTS_PLAN", ""), "auto")
_plan_slabs = int(_os.environ.get("NTS_PLAN_SLABS", "0"))


def set_plan_mode(mode, slabs=0):
    """mode in {"auto", "on", "off"}; slabs > 0 forces the slab count (0 = nts_gather_plan_pick_slabs)."""
    global _plan_mode, _plan_slabs
    if mode not in ("auto", "on", "off"):
        raise ValueError("plan mode must be auto, on or off")
    _plan_mode, _plan_slabs = mode, int(slabs)


def _chunk_plan(chunk, direction, F):
    """The GatherPlan of one chunk direction for feature width F, or None when the plain kernel should run."""
    if _plan_mode == "off" or (_plan_mode == "auto" and chunk.edge_size < PLAN_MIN_EDGES):
        return None
    if direction == "fwd":
        n_rows, gather_rows = chunk.batch_size_forward, chunk.batch_size_backward
    else:
        n_rows, gather_rows = chunk.batch_size_backward, chunk.batch_size_forward
    plans = chunk.__dict__.setdefault("_gather_plans", {})      # (direction, slabs) -> plan
    tuned = chunk.__dict__.setdefault("_gather_plan_for", {})   # (direction, F) -> plan picked by measurement
    key = (direction, _plan_slabs) if _plan_slabs else (direction, "F", int(F))
    plan = plans.get(key) if _plan_slabs else tuned.get(key)
    if plan is None:
        if direction == "fwd":
            plan = GatherPlan(chunk.column_offset_gpu, chunk.row_indices_gpu, chunk.edge_weight_forward_gpu,
                              chunk.src_range[0], n_rows, chunk.edge_size, gather_rows, _plan_slabs, tune_for=int(F))
        else:
            plan = GatherPlan(chunk.row_offset_gpu, chunk.column_indices_gpu, chunk.edge_weight_backward_gpu,
                              chunk.dst_range[0], n_rows, chunk.edge_size, gather_rows, _plan_slabs, tune_for=int(F))
        if (direction, plan.slabs) in plans:     # another width already settled on this slab count: share the arrays
            plan = plans[(direction, plan.slabs)]
        plans[(direction, plan.slabs)] = plan
        if not _plan_slabs:
            tuned[key] = plan
    return plan


def gather_by_dst_from_src(chunk, out, x, with_weight=True):
    """NtsScheduler::GatherByDstFromSrc (core/NtsScheduler.hpp:151-191) on one chunk."""
    plan = _chunk_plan(chunk, "fwd", x.shape[1]) if with_weight else None
    ev = _timer.bracket("fwd", x.shape[1], chunk.edge_size, chunk.batch_size_forward) if _timer else None
    if ev:
        ev[0].record()
    if plan is not None:
        plan.run(x, out)
    else:
        _lib.call("nts_gather_by_dst_from_src", _ptr(x), _ptr(out), _ptr(chunk.edge_weight_forward_gpu),
                  _ptr(chunk.row_indices_gpu), _ptr(chunk.column_offset_gpu), chunk.src_range[0], chunk.src_range[1],
                  chunk.dst_range[0], chunk.dst_range[1], chunk.edge_size, chunk.batch_size_forward,
                  int(x.shape[1]), 1 if with_weight else 0, _stream())
    if ev:
        ev[1].record()
    return out


def gather_by_src_from_dst(chunk, out, grad, with_weight=True):
    """NtsScheduler::GatherBySrcFromDst (core/NtsScheduler.hpp:257-293) on one chunk."""
    plan = _chunk_plan(chunk, "bwd", grad.shape[1]) if with_weight else None
    ev = _timer.bracket("bwd", grad.shape[1], chunk.edge_size, chunk.batch_size_backward) if _timer else None
    if ev:
        ev[0].record()
    if plan is not None:
        plan.run(grad, out)
    else:
        _lib.call("nts_gather_by_src_from_dst", _ptr(grad), _ptr(out), _ptr(chunk.edge_weight_backward_gpu),
                  _ptr(chunk.row_offset_gpu), _ptr(chunk.column_indices_gpu), chunk.src_range[0], chunk.src_range[1],
                  chunk.dst_range[0], chunk.dst_range[1], chunk.edge_size, chunk.batch_size_backward,
                  int(grad.shape[1]), 1 if with_weight else 0, _stream())
    if ev:
        ev[1].record()
    return out


class ntsGraphOp:
    """core/ntsBaseOp.hpp:24-48."""

    def __init__(self, partitioned_graph, active=None):
        self.partitioned_graph_ = partitioned_graph
        self.active_ = active

    def forward(self, f_input, f_input1=None):
        raise NotImplementedError

    def backward(self, output_grad):
        raise NotImplementedError

    def get_additional_grad(self):
        raise NotImplementedError("get_additional_grad is not implemented")


class ForwardSingleGPUfuseOp(ntsGraphOp):
    """core/ntsSingleGPUFusedGraphOp.hpp:48-71 -> Graph::forward_single / backward_single
    (core/graph.hpp:3805-3855): Y = A X on chunk 0, dX = A^T dY, no communication."""

    def forward(self, f_input, f_input1=None):
        x = _check_input(f_input)
        c = self.partitioned_graph_.graph_chunks[0]
        y = torch.zeros((c.batch_size_forward, x.shape[1]), dtype=torch.float32, device=x.device)
        return gather_by_dst_from_src(c, y, x)

    def backward(self, f_output_grad):
        g = _check_input(f_output_grad, "output_grad")
        c = self.partitioned_graph_.graph_chunks[0]
        dx = torch.zeros((c.batch_size_backward, g.shape[1]), dtype=torch.float32, device=g.device)
        return gather_by_src_from_dst(c, dx, g)


class ForwardGPUfuseOp(ntsGraphOp):
    """core/ntsDistGPUFusedGraphOp.hpp:48-90: the distributed fused GCN aggregation.  The reference drives it
    through Graph::sync_compute_decoupled / compute_sync_decoupled with host-staged MPI messages
    (core/graph.hpp:3455-3719); here the exchange is device-resident (neutronstarlite_b200.exchange)."""

    def __init__(self, partitioned_graph, active=None, exchange=None):
        super().__init__(partitioned_graph, active)
        if exchange is None:
            from .exchange import default_exchange
            exchange = default_exchange(partitioned_graph)
        self.exchange = exchange

    def forward(self, f_input, f_input1=None):
        return self.exchange.forward(_check_input(f_input))

    def backward(self, f_output_grad):
        return self.exchange.backward(_check_input(f_output_grad, "output_grad"))


# ---- edge-granular operators (core/ntsDistGPUGraphOp.hpp) ------------------------------------------------------
class _EdgeOp(ntsGraphOp):
    def _topo(self):
        pg = self.partitioned_graph_
        if pg.column_offset_gpu is None or pg.row_indices_gpu is None:
            raise _lib.NtsError("whole-partition CSC missing: call PartitionedGraph.generate_all(dist=True, device=...)")
        return pg


class DistGPUGetDepNbrOp(ntsGraphOp):
    """core/ntsDistGPUGraphOp.hpp:48-143: fetch the feature rows of every (local or remote) source of a local
    in-edge into the mirror matrix [owned_mirrors, F]; backward returns mirror gradients to the owners.
    Device-resident here (f4 of SURVEY 8f): no .cpu()/MPI/.cuda() round trip."""

    def __init__(self, partitioned_graph, active=None, exchange=None):
        super().__init__(partitioned_graph, active)
        if exchange is None:
            from .exchange import default_exchange
            exchange = default_exchange(partitioned_graph)
        self.exchange = exchange

    def forward(self, f_input, f_input1=None):
        return self.exchange.fetch_mirrors(_check_input(f_input))

    def backward(self, f_output_grad):
        return self.exchange.return_mirror_grads(_check_input(f_output_grad, "output_grad"))


class DistGPUScatterSrc(_EdgeOp):
    """core/ntsDistGPUGraphOp.hpp:100-176: mirror [M,F] -> edge messages [E_p,F]."""

    def forward(self, f_input, f_input1=None):
        pg = self._topo()
        x = _check_input(f_input)
        msg = torch.zeros((pg.owned_edges, x.shape[1]), dtype=torch.float32, device=x.device)
        _lib.call("nts_scatter_src_mirror_to_msg", _ptr(msg), _ptr(x), _ptr(pg.row_indices_gpu),
                  _ptr(pg.column_offset_gpu), _ptr(pg.mirror_index_gpu), pg.owned_vertices, x.shape[1], _stream())
        return msg

    def backward(self, f_output_grad):
        pg = self._topo()
        g = _check_input(f_output_grad, "output_grad")
        out = torch.zeros((pg.owned_mirrors, g.shape[1]), dtype=torch.float32, device=g.device)
        _lib.call("nts_gather_msg_to_src_mirror", _ptr(out), _ptr(g), _ptr(pg.row_indices_gpu),
                  _ptr(pg.column_offset_gpu), _ptr(pg.mirror_index_gpu), pg.owned_vertices, g.shape[1], _stream())
        return out


class DistGPUScatterDst(_EdgeOp):
    """core/ntsDistGPUGraphOp.hpp:178-238: local vertices [V_p,F] -> edge messages [E_p,F]."""

    def forward(self, f_input, f_input1=None):
        pg = self._topo()
        x = _check_input(f_input)
        msg = torch.zeros((pg.owned_edges, x.shape[1]), dtype=torch.float32, device=x.device)
        _lib.call("nts_scatter_dst_to_msg", _ptr(msg), _ptr(x), _ptr(pg.row_indices_gpu),
                  _ptr(pg.column_offset_gpu), pg.owned_vertices, x.shape[1], _stream())
        return msg

    def backward(self, f_output_grad):
        pg = self._topo()
        g = _check_input(f_output_grad, "output_grad")
        out = torch.zeros((pg.owned_vertices, g.shape[1]), dtype=torch.float32, device=g.device)
        _lib.call("nts_gather_msg_to_dst", _ptr(out), _ptr(g), _ptr(pg.row_indices_gpu),
                  _ptr(pg.column_offset_gpu), pg.owned_vertices, g.shape[1], _stream())
        return out


class DistGPUAggregateDst(_EdgeOp):
    """core/ntsDistGPUGraphOp.hpp:240-300: edge messages [E_p,F] -> sum per destination [V_p,F]."""

    def forward(self, f_input, f_input1=None):
        pg = self._topo()
        m = _check_input(f_input)
        out = torch.zeros((pg.owned_vertices, m.shape[1]), dtype=torch.float32, device=m.device)
        _lib.call("nts_gather_msg_to_dst", _ptr(out), _ptr(m), _ptr(pg.row_indices_gpu),
                  _ptr(pg.column_offset_gpu), pg.owned_vertices, m.shape[1], _stream())
        return out

    def backward(self, f_output_grad):
        pg = self._topo()
        g = _check_input(f_output_grad, "output_grad")
        msg = torch.zeros((pg.owned_edges, g.shape[1]), dtype=torch.float32, device=g.device)
        _lib.call("nts_scatter_dst_to_msg", _ptr(msg), _ptr(g), _ptr(pg.row_indices_gpu),
                  _ptr(pg.column_offset_gpu), pg.owned_vertices, g.shape[1], _stream())
        return msg


class DistGPUEdgeSoftMax(_EdgeOp):
    """core/ntsDistGPUGraphOp.hpp:302-361; numerics follow the CPU operator DistEdgeSoftMax
    (core/ntsDistCPUGraphOp.hpp:442-492): column-wise, max-subtracted."""

    def __init__(self, partitioned_graph, active=None):
        super().__init__(partitioned_graph, active)
        self.IntermediateResult = None

    def forward(self, f_input, f_input1=None):
        pg = self._topo()
        m = _check_input(f_input)
        out = torch.zeros_like(m)
        self.IntermediateResult = torch.zeros_like(m)
        _lib.call("nts_edge_softmax_forward", _ptr(out), _ptr(m), _ptr(self.IntermediateResult),
                  _ptr(pg.row_indices_gpu), _ptr(pg.column_offset_gpu), pg.owned_vertices, m.shape[1], _stream())
        return out

    def backward(self, f_output_grad):
        pg = self._topo()
        g = _check_input(f_output_grad, "output_grad")
        out = torch.zeros_like(g)
        _lib.call("nts_edge_softmax_backward", _ptr(out), _ptr(g), _ptr(self.IntermediateResult),
                  _ptr(pg.row_indices_gpu), _ptr(pg.column_offset_gpu), pg.owned_vertices, g.shape[1], _stream())
        return out


class DistGPUAggregateDstFuseWeight(_EdgeOp):
    """GPU twin of DistAggregateDstFuseWeight (core/ntsDistCPUGraphOp.hpp:499-594), the fused GAT aggregation
    of toolkits/GAT_CPU_DIST_OPTM.hpp:196-241: y[d,:] = sum_e a[e] * mirror[MirrorIndex[src(e)],:], never
    materialising an [E,F] message.  backward returns d_mirror; `get_additional_grad()` returns d_a [E,1]."""

    def __init__(self, partitioned_graph, active=None):
        super().__init__(partitioned_graph, active)
        self._mirror = None
        self._a = None
        self.e_weight_grad = None

    def forward(self, f_input, e_weight=None):
        pg = self._topo()
        x = _check_input(f_input)
        a = _check_input(e_weight, "edge weight")
        self._mirror, self._a = x, a
        heads = int(a.shape[1])  # [E, H]: head h weights columns [h*D, (h+1)*D); H = 1 is the reference's operator
        out = torch.zeros((pg.owned_vertices, x.shape[1]), dtype=torch.float32, device=x.device)
        _lib.call("nts_segment_gather_sum_heads", _ptr(x), _ptr(out), _ptr(a), _ptr(pg.row_indices_gpu),
                  _ptr(pg.column_offset_gpu), _ptr(pg.mirror_index_gpu), 0, pg.owned_vertices, pg.owned_edges,
                  x.shape[1], heads, _stream())
        return out

    def backward(self, f_output_grad):
        pg = self._topo()
        g = _check_input(f_output_grad, "output_grad")
        dm = torch.zeros((pg.owned_mirrors, g.shape[1]), dtype=torch.float32, device=g.device)
        self.e_weight_grad = torch.zeros_like(self._a)
        _lib.call("nts_aggregate_dst_fuse_weight_backward_heads", _ptr(dm), _ptr(self.e_weight_grad),
                  _ptr(self._mirror), _ptr(self._a), _ptr(g), _ptr(pg.row_indices_gpu), _ptr(pg.column_offset_gpu),
                  _ptr(pg.mirror_index_gpu), pg.owned_vertices, g.shape[1], int(self._a.shape[1]), _stream())
        return dm

    def get_additional_grad(self):
        return self.e_weight_grad


class DistGPUFusedGATOp(_EdgeOp):
    """K7: the whole attention + aggregation of one GAT layer in two kernels forward and two passes backward, never
    materialising an edge-sized tensor (toolkits/GAT_CPU_DIST_OPTM.hpp:196-241 keeps [E,1] logits / attention;
    toolkits/GAT_GPU_DIST.hpp:187-219 keeps four [E,F] messages).

        forward(mirror [M, H*D], src_score [M, H], dst_score [V, H]) -> out [V, H*D]
            a[e,h] = softmax over the in-edges of dst(e) of leaky_relu(src_score[slot(e),h] + dst_score[dst(e),h])
            out[d, hD:(h+1)D] = sum_e a[e,h] * mirror[slot(e), hD:(h+1)D]
        backward(grad_out) -> (d_mirror, d_src_score, d_dst_score)
    """

    def __init__(self, partitioned_graph, active=None, negative_slope=0.2, two_pass_backward=True):
        super().__init__(partitioned_graph, active)
        self.slope = float(negative_slope)
        self.two_pass_backward = bool(two_pass_backward)
        self._saved = None

    @staticmethod
    def slot_indices(pg):
        """row_indices with the MirrorIndex lookup already applied (mirror slot of every CSC edge), cached on the
        graph: the kernels then skip one dependent load per edge (the C ABI accepts mirror_index = NULL for this)."""
        cached = getattr(pg, "_slot_indices_gpu", None)
        if cached is None:
            cached = pg.mirror_index_gpu[pg.row_indices_gpu.long()].to(torch.int32).contiguous()
            pg._slot_indices_gpu = cached
        return cached

    @staticmethod
    def slot_csr(pg):
        """Out-edges of every mirror slot: (slot_row_offset [M+1], slot_column_indices [E], local destination ids) -
        the CSR twin of the whole-partition CSC, built once per PartitionedGraph on the device and cached on it."""
        cached = getattr(pg, "_slot_csr_gpu", None)
        if cached is not None:
            return cached
        col = pg.column_offset_gpu.long()
        V, M = pg.owned_vertices, pg.owned_mirrors
        slot = DistGPUFusedGATOp.slot_indices(pg).long()
        dst = torch.repeat_interleave(torch.arange(V, device=col.device), col[1:V + 1] - col[:V])
        order = torch.sort(slot, stable=True).indices
        csr_dst = dst[order].to(torch.int32).contiguous()
        off = torch.zeros(M + 1, dtype=torch.int64, device=col.device)
        off[1:] = torch.cumsum(torch.bincount(slot, minlength=M), 0)
        pg._slot_csr_gpu = (off.to(torch.int32).contiguous(), csr_dst)
        return pg._slot_csr_gpu

    def forward(self, mirror, src_score, dst_score):
        pg = self._topo()
        x = _check_input(mirror, "mirror")
        s = _check_input(src_score, "src_score")
        d = _check_input(dst_score, "dst_score")
        H = int(s.shape[1])
        seg_max = torch.empty((pg.owned_vertices, H), dtype=torch.float32, device=x.device)
        seg_sum = torch.empty_like(seg_max)
        slots = self.slot_indices(pg)
        with _timed("gat_stats", x.shape[1], pg.owned_edges, pg.owned_vertices):
            _lib.call("nts_gat_softmax_stats", _ptr(seg_max), _ptr(seg_sum), _ptr(s), _ptr(d), _ptr(slots),
                      _ptr(pg.column_offset_gpu), 0, pg.owned_vertices, H, self.slope, _stream())
        F = int(x.shape[1])
        xk, Fk = x, F
        if H == 1 and F % 4 != 0:
            # odd single-head width (the 41-wide output layer of config D): gather from a copy padded to a multiple
            # of 4 columns so that the kernel uses 16-byte loads and packed virtual warps instead of 4-byte gathers
            # (11.9 -> ~4 ms per call on config D); the zero columns are dropped again below
            Fk = (F + 3) // 4 * 4
            xk = torch.nn.functional.pad(x, (0, Fk - F))
        out = torch.zeros((pg.owned_vertices, Fk), dtype=torch.float32, device=x.device)
        with _timed("gat_fwd", F, pg.owned_edges, pg.owned_vertices):
            _lib.call("nts_gat_fused_aggregate_forward", _ptr(xk), _ptr(out), _ptr(s), _ptr(d), _ptr(seg_max),
                      _ptr(seg_sum), _ptr(slots), _ptr(pg.column_offset_gpu), 0,
                      pg.owned_vertices, pg.owned_edges, Fk, H, self.slope, _stream())
        if Fk != F:
            out = out[:, :F].contiguous()
        self._saved = (x, s, d, seg_max, seg_sum, out)
        return out

    def backward(self, f_output_grad):
        pg = self._topo()
        g = _check_input(f_output_grad, "output_grad")
        x, s, d, seg_max, seg_sum, out = self._saved
        H = int(s.shape[1])
        D = x.shape[1] // H
        # sum_e a[e,h] * <mirror[slot(e),h], g[d,h]> == <out[d,h], g[d,h]>: the softmax backward needs no edge pass
        out_dot_g = (out.detach() * g).view(-1, H, D).sum(-1).contiguous()
        dm = torch.zeros_like(x)
        ds = torch.zeros_like(s)
        dd = torch.zeros_like(d)
        if self.two_pass_backward:
            # no per-edge atomics: a destination-major and a source-major pass, each with register accumulators
            slot_off, slot_dst = self.slot_csr(pg)
            pack = torch.empty((pg.owned_vertices, H, 4), dtype=torch.float32, device=x.device)
            with _timed("gat_bwd", x.shape[1], 2 * pg.owned_edges, pg.owned_vertices):
                _lib.call("nts_gat_fused_aggregate_backward_two_pass", _ptr(dm), _ptr(ds), _ptr(dd), _ptr(pack),
                          _ptr(x), _ptr(s), _ptr(d), _ptr(seg_max), _ptr(seg_sum), _ptr(out_dot_g), _ptr(g),
                          _ptr(self.slot_indices(pg)), _ptr(pg.column_offset_gpu), 0,
                          _ptr(slot_off), _ptr(slot_dst), pg.owned_vertices, x.shape[0], x.shape[1], H, self.slope,
                          _stream())
            return dm, ds, dd
        _lib.call("nts_gat_fused_aggregate_backward", _ptr(dm), _ptr(ds), _ptr(dd), _ptr(x), _ptr(s), _ptr(d),
                  _ptr(seg_max), _ptr(seg_sum), _ptr(out_dot_g), _ptr(g), _ptr(pg.row_indices_gpu),
                  _ptr(pg.column_offset_gpu), _ptr(pg.mirror_index_gpu), pg.owned_vertices, x.shape[1], H,
                  self.slope, _stream())
        return dm, ds, dd
