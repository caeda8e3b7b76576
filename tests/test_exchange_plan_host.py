"""The C++ exchange-plan builder (nts_exchange_plan_*, csrc/nts_exchange_plan.cu) on the reference's own chunk arrays
(golden dumps of the unmodified reference at P = 2, 4, 8): all ranks are simulated in one process, the "all-to-all" of
the need lists is a Python loop, and the exchange itself is replayed with numpy on the plan's host view -

    forward : staging = rows the plan says to read from every peer;  Y_p = local CSC + merged remote CSC over staging
    backward: partial rows over the merged compact CSR, returned to the owners through send_rows / peer_bwd_offset

- and compared with the reference's ForwardCPUfuseOp results (gcn_Y / gcn_dX of the golden files).  No GPU involved:
only the host half of the builder runs."""
import ctypes as C

import numpy as np
import pytest

from neutronstarlite_b200 import _lib


def _u32p(a):
    return a.ctypes.data_as(C.c_void_p)


def _segment_sum(off, idx, w, x, rows):
    out = np.zeros((rows, x.shape[1]), dtype=np.float64)
    off = off.astype(np.int64)
    seg = np.repeat(np.arange(rows), off[1:] - off[:-1])
    np.add.at(out, seg, x[idx.astype(np.int64)].astype(np.float64) * w[:, None].astype(np.float64))
    return out


def _close(a, ref):
    assert a.shape == ref.shape
    if ref.size:
        assert np.abs(a - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max())


def _np(ptr, n, dtype):
    if n == 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)


def _need(plan, i):
    cnt = C.c_uint32(0)
    ptr = _lib.load().nts_exchange_plan_need(plan, i, C.byref(cnt))
    return _np(ptr, cnt.value, np.uint32)


class _Rank:
    def __init__(self, g, r):
        L = _lib.load()
        self.keep = []
        P = g.P
        arr = (_lib.HostChunk * P)()
        self.chunks = []
        for i in range(P):
            t = "chunk%d_" % i
            meta = g.get(r, t + "meta")
            c = {name: np.ascontiguousarray(g.get(r, t + name)) for name in
                 ("column_offset", "row_indices", "row_offset", "column_indices", "edge_weight_forward",
                  "edge_weight_backward")}
            c["meta"] = [int(m) for m in meta]
            self.chunks.append(c)
            h = arr[i]
            h.column_offset, h.row_indices = _u32p(c["column_offset"]), _u32p(c["row_indices"])
            h.row_offset, h.column_indices = _u32p(c["row_offset"]), _u32p(c["column_indices"])
            h.edge_weight_forward, h.edge_weight_backward = _u32p(c["edge_weight_forward"]), _u32p(c["edge_weight_backward"])
            h.src_start, h.src_end, h.dst_start, h.dst_end = (int(meta[3]), int(meta[4]), int(meta[5]), int(meta[6]))
            h.edges = int(meta[0])
        self.arr = arr
        self.plan = L.nts_exchange_plan_create(arr, P, r)
        assert self.plan, L.nts_last_error()
        n = int(L.nts_exchange_plan_packed_rows(self.plan))
        self.counts = np.zeros(P, dtype=np.uint32)
        self.rows = np.zeros(max(n, 1), dtype=np.uint32)
        assert L.nts_exchange_plan_pack_needs(self.plan, _u32p(self.counts), _u32p(self.rows)) == 0
        self.rows = self.rows[:n]

    def view(self):
        L = _lib.load()
        v = _lib.ExchangePlanView()
        assert L.nts_exchange_plan_get_view(self.plan, C.byref(v)) == 0, L.nts_last_error()
        P = v.partitions
        out = {"P": P, "rank": v.rank, "Vp": v.owned_vertices, "recv_total": v.recv_total, "send_total": v.send_total,
               "remote_edges": int(v.remote_edges),
               "need_count": _np(v.need_count, P, np.uint32), "send_count": _np(v.send_count, P, np.uint32),
               "peer_bwd_offset": _np(v.peer_bwd_offset, P, np.uint32),
               "fwd_push_offset": _np(v.fwd_push_offset, P, np.uint32),
               "bwd_push_offset": _np(v.bwd_push_offset, P, np.uint32),
               "remote_column_offset": _np(v.remote_column_offset, v.owned_vertices + 1 if v.remote_edges else 0, np.uint32),
               "remote_slots": _np(v.remote_slots, int(v.remote_edges), np.uint32),
               "remote_weight": _np(v.remote_weight, int(v.remote_edges), np.float32),
               "backward_offsets": _np(v.backward_offsets, v.backward_rows + 1, np.uint32),
               "backward_indices": _np(v.backward_indices, int(v.remote_edges), np.uint32),
               "backward_weight": _np(v.backward_weight, int(v.remote_edges), np.float32),
               "send_rows_all": _np(v.send_rows_all, v.send_total, np.uint32)}
        return out

    def close(self):
        _lib.load().nts_exchange_plan_destroy(self.plan)
        self.plan = None


def test_plan_replays_the_reference_exchange(golden):
    g = golden
    if g.P == 1:
        pytest.skip("single partition: nothing to exchange")
    L = _lib.load()
    ranks = [_Rank(g, r) for r in range(g.P)]
    try:
        for a in ranks:                                   # the control plane: everybody learns everybody's lists
            for j, b in enumerate(ranks):
                assert L.nts_exchange_plan_set_peer_needs(a.plan, j, _u32p(b.counts), _u32p(b.rows)) == 0, \
                    L.nts_last_error()
            assert L.nts_exchange_plan_finalize(a.plan) == 0, L.nts_last_error()
        views = [a.view() for a in ranks]
        po = g.partition_offset.astype(np.int64)
        X = [g.mat(r, "X").astype(np.float32) for r in range(g.P)]
        G = [g.mat(r, "G").astype(np.float32) for r in range(g.P)]
        need = [[_need(a.plan, i) for i in range(g.P)] for a in ranks]
        # consistency of the two directions: what j sends to r is what r needs from j
        for r in range(g.P):
            v = views[r]
            assert v["recv_total"] == sum(len(need[r][i]) for i in range(g.P) if i != r)
            pos = 0
            for j in range(g.P):
                if j == r:
                    continue
                n = int(v["send_count"][j])
                assert np.array_equal(v["send_rows_all"][pos:pos + n], need[j][r])
                pos += n
                assert int(v["peer_bwd_offset"][j]) == sum(len(need[j][q]) for q in range(r) if q != j)
        # ---- forward
        for r in range(g.P):
            v = views[r]
            Vp = int(po[r + 1] - po[r])
            lc = ranks[r].chunks[r]
            y = np.zeros((Vp, g.F))
            if lc["meta"][0]:
                y += _segment_sum(lc["column_offset"], lc["row_indices"].astype(np.int64) - po[r],
                                  lc["edge_weight_forward"], X[r], Vp)
            staging = [X[i][need[r][i]] for i in range(g.P) if i != r]
            staging = np.concatenate(staging, axis=0) if staging else np.zeros((0, g.F), dtype=np.float32)
            assert staging.shape[0] == v["recv_total"]
            if v["remote_edges"]:
                y += _segment_sum(v["remote_column_offset"], v["remote_slots"], v["remote_weight"], staging, Vp)
            ref = g.mat(r, "gcn_Y")
            _close(y, ref)
        # ---- backward: partial rows of every rank, then the owners add their slices
        partial = []
        for r in range(g.P):
            v = views[r]
            rows = len(v["backward_offsets"]) - 1
            assert rows == v["recv_total"]
            if v["remote_edges"]:
                part = _segment_sum(v["backward_offsets"], v["backward_indices"].astype(np.int64) - po[r],
                                    v["backward_weight"], G[r], rows)
            else:
                part = np.zeros((rows, g.F))
            partial.append(part)
        for r in range(g.P):
            v = views[r]
            Vp = int(po[r + 1] - po[r])
            lc = ranks[r].chunks[r]
            dx = np.zeros((Vp, g.F))
            if lc["meta"][0]:
                dx += _segment_sum(lc["row_offset"], lc["column_indices"].astype(np.int64) - po[r],
                                   lc["edge_weight_backward"], G[r], Vp)
            pos = 0
            for j in range(g.P):
                if j == r:
                    continue
                n = int(v["send_count"][j])
                rows = v["send_rows_all"][pos:pos + n].astype(np.int64)
                o = int(v["peer_bwd_offset"][j])
                np.add.at(dx, rows, partial[j][o:o + n])
                pos += n
            ref = g.mat(r, "gcn_dX")
            _close(dx, ref)
    finally:
        for a in ranks:
            a.close()


def _plan_chunk(plan, i, edges, need_n):
    sl, oc = C.POINTER(C.c_uint32)(), C.POINTER(C.c_uint32)()
    assert _lib.load().nts_exchange_plan_chunk(plan, i, C.byref(sl), C.byref(oc)) == 0
    return _np(sl, edges, np.uint32), _np(oc, need_n + 1, np.uint32)


def test_plan_replays_the_push_engine(golden):
    """The data path of csrc/nts_exchange.cu (PUSH model, one stage per source partition) walked in numpy on the C++
    plan's per-chunk arrays and push offsets: every rank stores the rows its peers read straight into THEIR receive
    staging at fwd_push_offset, aggregates chunk after chunk from the staged slices; backward partials go to the
    owner's gradient staging at bwd_push_offset and are added in send_rows_all order."""
    g = golden
    if g.P == 1:
        pytest.skip("single partition: nothing to exchange")
    L = _lib.load()
    P = g.P
    ranks = [_Rank(g, r) for r in range(P)]
    try:
        for a in ranks:
            for j, b in enumerate(ranks):
                assert L.nts_exchange_plan_set_peer_needs(a.plan, j, _u32p(b.counts), _u32p(b.rows)) == 0
            assert L.nts_exchange_plan_finalize(a.plan) == 0, L.nts_last_error()
        views = [a.view() for a in ranks]
        po = g.partition_offset.astype(np.int64)
        X = [g.mat(r, "X").astype(np.float32) for r in range(P)]
        G = [g.mat(r, "G").astype(np.float32) for r in range(P)]

        def offs(counts, skip):
            o = np.zeros(P + 1, dtype=np.int64)
            for i in range(P):
                o[i + 1] = o[i] + (0 if i == skip else int(counts[i]))
            return o

        recv_offs = [offs(views[r]["need_count"], r) for r in range(P)]
        srecv_offs = [offs(views[r]["send_count"], r) for r in range(P)]
        # ---- forward pushes: NaN-filled windows prove every slot is written exactly where it is read
        window = [np.full((views[r]["recv_total"], g.F), np.nan, dtype=np.float32) for r in range(P)]
        for p in range(P):
            v = views[p]
            for j in range(P):
                if j == p:
                    continue
                rows = v["send_rows_all"][srecv_offs[p][j]:srecv_offs[p][j + 1]].astype(np.int64)
                o = int(v["fwd_push_offset"][j])
                assert o == recv_offs[j][p]
                window[j][o:o + rows.shape[0]] = X[p][rows]
        for r in range(P):
            assert not np.isnan(window[r]).any()
            Vp = int(po[r + 1] - po[r])
            lc = ranks[r].chunks[r]
            y = np.zeros((Vp, g.F))
            if lc["meta"][0]:
                y += _segment_sum(lc["column_offset"], lc["row_indices"].astype(np.int64) - po[r],
                                  lc["edge_weight_forward"], X[r], Vp)
            for s in range(1, P):                      # ring order of the engine
                i = (r + s) % P
                c = ranks[r].chunks[i]
                n_i = int(views[r]["need_count"][i])
                if not c["meta"][0]:
                    continue
                slots, _ = _plan_chunk(ranks[r].plan, i, c["meta"][0], n_i)
                assert slots.max() < n_i
                y += _segment_sum(c["column_offset"], slots, c["edge_weight_forward"],
                                  window[r][recv_offs[r][i]:recv_offs[r][i + 1]], Vp)
            _close(y, g.mat(r, "gcn_Y"))
        # ---- backward pushes
        gwin = [np.full((views[r]["send_total"], g.F), np.nan) for r in range(P)]
        for p in range(P):
            for i in range(P):
                if i == p:
                    continue
                c = ranks[p].chunks[i]
                n_i = int(views[p]["need_count"][i])
                _, offc = _plan_chunk(ranks[p].plan, i, c["meta"][0], n_i)
                part = _segment_sum(offc, c["column_indices"].astype(np.int64) - po[p], c["edge_weight_backward"],
                                    G[p], n_i) if c["meta"][0] else np.zeros((n_i, g.F))
                o = int(views[p]["bwd_push_offset"][i])
                assert o == srecv_offs[i][p]
                gwin[i][o:o + n_i] = part
        for r in range(P):
            assert not np.isnan(gwin[r]).any()
            Vp = int(po[r + 1] - po[r])
            lc = ranks[r].chunks[r]
            dx = np.zeros((Vp, g.F))
            if lc["meta"][0]:
                dx += _segment_sum(lc["row_offset"], lc["column_indices"].astype(np.int64) - po[r],
                                   lc["edge_weight_backward"], G[r], Vp)
            np.add.at(dx, views[r]["send_rows_all"].astype(np.int64), gwin[r])
            _close(dx, g.mat(r, "gcn_dX"))
    finally:
        for a in ranks:
            a.close()


def test_plan_rejects_incomplete_and_bad_input():
    L = _lib.load()
    assert not L.nts_exchange_plan_create(None, 2, 0)
    off = np.zeros(3, dtype=np.uint32)
    arr = (_lib.HostChunk * 2)()
    for i in range(2):
        arr[i].column_offset = arr[i].row_offset = _u32p(off)
        arr[i].src_start, arr[i].src_end = 2 * i, 2 * i + 2
        arr[i].dst_start, arr[i].dst_end = 0, 2
        arr[i].edges = 0
    plan = L.nts_exchange_plan_create(arr, 2, 0)
    assert plan
    try:
        L.nts_set_abort_on_error(0) if hasattr(L, "nts_set_abort_on_error") else None
        assert L.nts_exchange_plan_finalize(plan) != 0          # peer 1 never reported its lists
        counts = np.zeros(2, dtype=np.uint32)
        assert L.nts_exchange_plan_set_peer_needs(plan, 1, _u32p(counts), None) == 0
        assert L.nts_exchange_plan_finalize(plan) == 0
        v = _lib.ExchangePlanView()
        assert L.nts_exchange_plan_get_view(plan, C.byref(v)) == 0
        assert v.recv_total == 0 and v.send_total == 0 and v.remote_edges == 0 and v.backward_rows == 0
    finally:
        L.nts_exchange_plan_destroy(plan)


@pytest.mark.parametrize("V,E,P,seed", [(5000, 40000, 3, 1), (3000, 9000, 4, 2), (9000, 200, 5, 3), (2048, 30000, 2, 4),
                                        (7000, 60000, 6, 5)])
def test_plan_on_random_graphs_matches_dense_product(V, E, P, seed):
    """Chunks from the product's own host builder (random multigraphs, hubs, partitions left empty by the 1024-aligned
    partitioner, ranks without remote edges), plan from the C++ builder, exchange replayed in numpy: Y = A X and
    dX = A^T G against a float64 scatter over the edge list."""
    from neutronstarlite_b200.graph import HostGraph, PartitionedGraph
    L = _lib.load()
    rng = np.random.default_rng(seed)
    src = rng.integers(0, V, E, dtype=np.uint32)
    dst = rng.integers(0, V, E, dtype=np.uint32)
    hub = rng.integers(0, V)
    dst[: E // 5] = hub                                  # one hub destination
    src[E // 5: E // 3] = rng.integers(0, V)             # and one hub source
    edges = np.stack([src, dst], 1)
    hg = HostGraph(edges, V)
    F = 3
    X = rng.uniform(-1, 1, (V, F)).astype(np.float32)
    G = rng.uniform(-1, 1, (V, F)).astype(np.float32)
    pgs = [PartitionedGraph(hg, P, r).generate_all(dist=True) for r in range(P)]
    po = pgs[0].partition_offset.astype(np.int64)
    out_d, in_d = hg.degrees()
    w = (1.0 / (np.sqrt(out_d[src].astype(np.float64)).astype(np.float32) *
                np.sqrt(in_d[dst].astype(np.float64)).astype(np.float32))).astype(np.float32)
    Y_ref = np.zeros((V, F))
    np.add.at(Y_ref, dst.astype(np.int64), X[src.astype(np.int64)].astype(np.float64) * w[:, None])
    dX_ref = np.zeros((V, F))
    np.add.at(dX_ref, src.astype(np.int64), G[dst.astype(np.int64)].astype(np.float64) * w[:, None])

    plans, packs, keep = [], [], []
    try:
        for r in range(P):
            arr = (_lib.HostChunk * P)()
            for i, c in enumerate(pgs[r].graph_chunks):
                h = arr[i]
                for name in ("column_offset", "row_indices", "row_offset", "column_indices", "edge_weight_forward",
                             "edge_weight_backward"):
                    setattr(h, name, getattr(c, name).ctypes.data_as(C.c_void_p))
                h.src_start, h.src_end = int(c.src_range[0]), int(c.src_range[1])
                h.dst_start, h.dst_end = int(c.dst_range[0]), int(c.dst_range[1])
                h.edges = int(c.edge_size)
            keep.append(arr)
            pl = L.nts_exchange_plan_create(arr, P, r)
            assert pl, L.nts_last_error()
            plans.append(pl)
            n = int(L.nts_exchange_plan_packed_rows(pl))
            counts = np.zeros(P, dtype=np.uint32)
            rows = np.zeros(n + 1, dtype=np.uint32)
            assert L.nts_exchange_plan_pack_needs(pl, _u32p(counts), _u32p(rows)) == 0
            packs.append((counts, rows))
        views = []
        for r in range(P):
            for j in range(P):
                assert L.nts_exchange_plan_set_peer_needs(plans[r], j, _u32p(packs[j][0]), _u32p(packs[j][1])) == 0, \
                    L.nts_last_error()
            assert L.nts_exchange_plan_finalize(plans[r]) == 0, L.nts_last_error()
            v = _lib.ExchangePlanView()
            assert L.nts_exchange_plan_get_view(plans[r], C.byref(v)) == 0
            views.append(v)
        need = [[_need(plans[r], i) for i in range(P)] for r in range(P)]
        partial = []
        for r in range(P):
            v = views[r]
            Vp = int(po[r + 1] - po[r])
            c = pgs[r].graph_chunks[r]
            y = np.zeros((Vp, F))
            if c.edge_size:
                y += _segment_sum(c.column_offset, c.row_indices.astype(np.int64) - po[r], c.edge_weight_forward,
                                  X[po[r]:po[r + 1]], Vp)
            stag = [X[po[i]:po[i + 1]][need[r][i]] for i in range(P) if i != r]
            stag = np.concatenate(stag) if stag else np.zeros((0, F), dtype=np.float32)
            assert stag.shape[0] == v.recv_total
            E_r = int(v.remote_edges)
            if E_r:
                y += _segment_sum(_np(v.remote_column_offset, Vp + 1, np.uint32), _np(v.remote_slots, E_r, np.uint32),
                                  _np(v.remote_weight, E_r, np.float32), stag, Vp)
            _close(y, Y_ref[po[r]:po[r + 1]])
            rows = v.backward_rows
            assert rows == v.recv_total
            if E_r:
                part = _segment_sum(_np(v.backward_offsets, rows + 1, np.uint32),
                                    _np(v.backward_indices, E_r, np.uint32).astype(np.int64) - po[r],
                                    _np(v.backward_weight, E_r, np.float32), G[po[r]:po[r + 1]], rows)
            else:
                part = np.zeros((rows, F))
            partial.append(part)
        for r in range(P):
            v = views[r]
            Vp = int(po[r + 1] - po[r])
            c = pgs[r].graph_chunks[r]
            dx = np.zeros((Vp, F))
            if c.edge_size:
                dx += _segment_sum(c.row_offset, c.column_indices.astype(np.int64) - po[r], c.edge_weight_backward,
                                   G[po[r]:po[r + 1]], Vp)
            send_rows = _np(v.send_rows_all, v.send_total, np.uint32).astype(np.int64)
            send_count = _np(v.send_count, P, np.uint32)
            off = _np(v.peer_bwd_offset, P, np.uint32)
            pos = 0
            for j in range(P):
                if j == r:
                    continue
                n = int(send_count[j])
                np.add.at(dx, send_rows[pos:pos + n], partial[j][int(off[j]):int(off[j]) + n])
                pos += n
            _close(dx, dX_ref[po[r]:po[r + 1]])
    finally:
        for pl in plans:
            L.nts_exchange_plan_destroy(pl)
