"""Host-side layout (nts_graph_host.cpp through the C ABI) against the golden artefacts dumped by the unmodified
reference at P = 1, 2, 4, 8: partition offsets, degrees, every chunk's CSC/CSR arrays and weights, MirrorIndex and
the whole-partition CSC - all bit-exact (CSR rows as multisets: their order is a race in the reference)."""
import numpy as np

from neutronstarlite_b200.graph import HostGraph, PartitionedGraph


def rows_as_multisets(offsets, idx):
    out = idx.copy()
    off = offsets.astype(np.int64)
    for r in range(off.shape[0] - 1):
        out[off[r]:off[r + 1]] = np.sort(out[off[r]:off[r + 1]])
    return out


def test_degrees_partition_offsets(golden):
    g = golden
    hg = HostGraph(g.edges, g.V)
    out_d, in_d = hg.degrees()
    assert np.array_equal(out_d, g.get(0, "out_degree"))
    assert np.array_equal(in_d, g.get(0, "in_degree"))
    assert np.array_equal(hg.partition_offsets(g.P), g.partition_offset)


def test_chunks_mirror_index_whole_topo(golden):
    g = golden
    hg = HostGraph(g.edges, g.V)
    for r in range(g.P):
        pg = PartitionedGraph(hg, g.P, r).generate_all(dist=True)
        assert pg.owned_vertices == int(g.get(r, "meta")[4])
        assert pg.owned_edges == int(g.get(r, "meta")[5])
        assert pg.owned_mirrors == int(g.get(r, "meta")[6])
        for i, c in enumerate(pg.graph_chunks):
            t = "chunk%d_" % i
            meta = g.get(r, t + "meta")
            assert (meta[0], meta[1], meta[2]) == (c.edge_size, c.batch_size_forward, c.batch_size_backward)
            assert (meta[3], meta[4]) == c.src_range and (meta[5], meta[6]) == c.dst_range
            assert np.array_equal(c.column_offset, g.get(r, t + "column_offset"))
            assert np.array_equal(c.row_indices, g.get(r, t + "row_indices"))
            assert np.array_equal(c.edge_weight_forward.view(np.uint32),
                                  g.get(r, t + "edge_weight_forward").view(np.uint32))
            assert np.array_equal(c.row_offset, g.get(r, t + "row_offset"))
            ref_ci = g.get(r, t + "column_indices")
            assert np.array_equal(c.column_indices, rows_as_multisets(c.row_offset, ref_ci))
            # backward weights: compare as per-row multisets of (dst, weight-bits) pairs
            ref_w = g.get(r, t + "edge_weight_backward").view(np.uint32).astype(np.uint64)
            mine_w = c.edge_weight_backward.view(np.uint32).astype(np.uint64)
            ref_pairs = (ref_ci.astype(np.uint64) << np.uint64(32)) | ref_w
            my_pairs = (c.column_indices.astype(np.uint64) << np.uint64(32)) | mine_w
            assert np.array_equal(rows_as_multisets(c.row_offset, my_pairs),
                                  rows_as_multisets(c.row_offset, ref_pairs))
            assert np.array_equal(c.source_active, g.get(r, t + "source_active"))
        assert np.array_equal(pg.MirrorIndex, g.get(r, "mirror_index"))
        assert np.array_equal(pg.column_offset, g.get(r, "whole_column_offset"))
        assert np.array_equal(pg.row_indices, g.get(r, "whole_row_indices"))


def test_empty_graph_and_single_vertex():
    hg = HostGraph(np.zeros((0, 2), dtype=np.uint32), 5)
    pg = PartitionedGraph(hg, 1, 0).generate_all(dist=True)
    assert pg.owned_edges == 0 and pg.owned_mirrors == 0
    assert np.array_equal(pg.graph_chunks[0].column_offset, np.zeros(6, dtype=np.uint32))
    out_d, in_d = hg.degrees()
    assert (out_d == 1).all() and (in_d == 1).all()


def test_partition_offsets_from_degrees_matches_reference(golden):
    from neutronstarlite_b200.graph import partition_offsets_from_out_degree
    g = golden
    raw = np.bincount(g.edges[:, 0], minlength=g.V)
    assert np.array_equal(partition_offsets_from_out_degree(raw, g.E, g.P), g.partition_offset)


def test_host_builder_rejects_bad_input():
    """Out-of-range vertex ids and bad ranks are reported through the status code (no silent clamping)."""
    from neutronstarlite_b200 import _lib
    L = _lib.load()
    edges = np.array([[0, 1], [7, 2]], dtype=np.uint32)   # vertex 7 does not exist in a 5-vertex graph
    out_d = np.zeros(5, dtype=np.uint32)
    in_d = np.zeros(5, dtype=np.uint32)
    assert L.nts_host_degrees(edges.ctypes.data, 2, 5, out_d.ctypes.data, in_d.ctypes.data) != 0
    po = np.zeros(3, dtype=np.uint32)
    assert L.nts_host_partition_offsets(edges.ctypes.data, 2, 5, 2, po.ctypes.data) != 0
    ok = np.array([[0, 1], [4, 2]], dtype=np.uint32)
    assert L.nts_host_partition_offsets(ok.ctypes.data, 2, 5, 0, po.ctypes.data) != 0      # zero partitions
    counts = np.zeros(2, dtype=np.uint64)
    po2 = np.array([0, 0, 5], dtype=np.uint32)
    assert L.nts_host_chunk_edge_counts(ok.ctypes.data, 2, po2.ctypes.data, 2, 5, counts.ctypes.data) != 0  # bad rank


def test_duplicates_and_self_loops_count_in_degrees():
    """The reference counts duplicate edges and self loops in both degrees (data/cora.2708.edge.self has 302
    duplicate pairs); weights follow from those counts."""
    e = np.array([[0, 1], [0, 1], [1, 1], [2, 1], [2, 0]], dtype=np.uint32)
    hg = HostGraph(e, 4)
    out_d, in_d = hg.degrees()
    assert out_d.tolist() == [2, 1, 2, 1] and in_d.tolist() == [1, 4, 1, 1]   # vertex 3 isolated -> clamped to 1
    pg = PartitionedGraph(hg, 1, 0).generate_all()
    c = pg.graph_chunks[0]
    assert c.column_offset.tolist() == [0, 1, 5, 5, 5]
    assert c.row_indices.tolist() == [2, 0, 0, 1, 2]
    w = 1 / (np.sqrt(out_d[c.row_indices].astype(np.float64)).astype(np.float32) *
             np.sqrt(in_d[[0, 1, 1, 1, 1]].astype(np.float64)).astype(np.float32))
    assert np.array_equal(c.edge_weight_forward, w.astype(np.float32))


def test_host_builder_equals_the_pinned_oracle_on_random_graphs():
    """Beyond the golden cases: random multigraphs x partition counts (incl. non-powers of two and partitions the
    1024-aligned partitioner leaves empty).  The numpy restatement is pinned against the reference's own dumps
    (tests/test_oracle_golden.py); the product's C++ host builder must agree with it bit for bit here."""
    import sys
    import os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import nts_oracle as O   # checker only
    rng = np.random.default_rng(7)
    for V, E, P in [(3000, 20000, 3), (7000, 50000, 5), (12000, 9000, 7), (2048, 4000, 2), (1500, 6000, 4),
                    (20000, 120000, 6)]:
        src = rng.integers(0, V, E).astype(np.uint32)
        dst = (rng.zipf(1.3, E) % V).astype(np.uint32)          # skewed destinations
        src[: E // 10] = rng.integers(0, V)                     # one hub source
        edges = np.stack([src, dst], 1)
        hg = HostGraph(edges, V)
        out_d, in_d = hg.degrees()
        o_out, o_in = O.degrees(edges, V)
        assert np.array_equal(out_d, o_out) and np.array_equal(in_d, o_in)
        po = hg.partition_offsets(P)
        assert np.array_equal(po, O.partition_offsets(edges, V, P))
        for r in range(P):
            pg = PartitionedGraph(hg, P, r).generate_all(dist=True)
            ref = O.build_chunks(edges, V, po, r, o_out, o_in)
            for c, rc in zip(pg.graph_chunks, ref):
                assert c.edge_size == rc.edge_size and tuple(c.src_range) == tuple(rc.src_range)
                assert np.array_equal(c.column_offset, rc.column_offset)
                assert np.array_equal(c.row_indices, rc.row_indices)
                assert np.array_equal(c.edge_weight_forward.view(np.uint32), rc.edge_weight_forward.view(np.uint32))
                assert np.array_equal(c.row_offset, rc.row_offset)
                assert np.array_equal(rows_as_multisets(c.row_offset, c.column_indices),
                                      rows_as_multisets(rc.row_offset, rc.column_indices))
                assert np.array_equal(c.source_active, rc.source_active)
            mi, n_mirrors = O.mirror_index(edges, V, po, r)
            assert np.array_equal(pg.MirrorIndex, mi) and pg.owned_mirrors == n_mirrors


def test_streaming_generator_matches_one_shot_generator():
    """Config E is generated per partition in two streaming passes (synth.zipf_degrees / zipf_edges_owned): same
    graph as the one-shot generator - degrees exactly, owned edge multiset exactly - for every rank of a 3-way split."""
    import torch
    from neutronstarlite_b200 import synth
    d = torch.device("cpu")
    V, E = 6000, 150000
    s, t = synth.zipf_edges(V, E, d, chunk=1 << 15)
    od, idg = synth.zipf_degrees(V, E, d, chunk=1 << 15)
    assert torch.equal(od, torch.bincount(s, minlength=V)) and torch.equal(idg, torch.bincount(t, minlength=V))
    cuts = [0, 1024, 4096, V]
    seen = 0
    for r in range(3):
        so, to = synth.zipf_edges_owned(V, E, d, cuts[r], cuts[r + 1], chunk=1 << 15)
        keep = (t >= cuts[r]) & (t < cuts[r + 1])
        assert torch.equal((to * V + so).sort().values, (t[keep] * V + s[keep]).sort().values)
        seen += int(so.numel())
    assert seen == int(s.numel())
