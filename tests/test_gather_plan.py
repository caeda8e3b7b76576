"""nts_gather_plan (source-slab bucketing + interleaved pairs + padded 16-byte gathers, csrc/nts_plan.cu) against the C
oracle of the reference's aggregation loop (core/ntsCPUFusedGraphOp.hpp:81-106 -> oracle/nts_oracle.c) and against the
plain kernel on the reference layout.  Integer artefacts of the plan (segment sizes) are checked exactly through the
all-ones product; float results per ROW relative to that row's own magnitude (1e-4, north_star)."""
import numpy as np
import pytest

import oracle_c

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def up_u32(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.uint32).view(np.int32)).to(dev())


def row_close(actual, desired, rtol=1e-4):
    """max |err| of every row <= rtol * max |desired| of THAT row (+ a denormal-sized floor)."""
    err = np.abs(actual.astype(np.float64) - desired.astype(np.float64)).max(axis=1)
    scale = np.abs(desired).max(axis=1).astype(np.float64)
    bad = np.nonzero(err > rtol * scale + 1e-30)[0]
    assert bad.size == 0, "rows %s: err %s vs scale %s" % (bad[:5], err[bad[:5]], scale[bad[:5]])


def make_graph(rng, n_rows, n_src, n_edges, hub_frac=0.2, empty_every=7):
    dst = rng.integers(0, n_rows, n_edges)
    dst[: int(n_edges * hub_frac)] = n_rows // 3            # one hub destination
    dst = dst[dst % empty_every != 1] if empty_every else dst  # some destinations without edges
    src = rng.integers(0, n_src, dst.shape[0])
    src[: dst.shape[0] // 10] = n_src - 1                   # a hub source in the last slab
    order = np.lexsort((src, dst))
    dst, src = dst[order], src[order]
    off = np.zeros(n_rows + 1, dtype=np.uint32)
    np.add.at(off, dst + 1, 1)
    off = np.cumsum(off).astype(np.uint32)
    w = rng.uniform(0.1, 1.0, dst.shape[0]).astype(np.float32)
    return off, src.astype(np.uint32), w


def run_plan(off, idx, w, X, base, slabs, n_src, slot_of=None):
    from neutronstarlite_b200 import ops
    d_off, d_idx = up_u32(off), up_u32(idx)
    d_w = None if w is None else torch.from_numpy(w).to(dev())
    d_s = None if slot_of is None else up_u32(slot_of)
    plan = ops.GatherPlan(d_off, d_idx, d_w, base, off.shape[0] - 1, idx.shape[0], n_src, slabs, slot_of=d_s)
    x = torch.from_numpy(X).to(dev())
    out = torch.zeros((off.shape[0] - 1, X.shape[1]), dtype=torch.float32, device=dev())
    plan.run(x, out)
    torch.cuda.synchronize()
    return plan, out.cpu().numpy()


@pytest.mark.parametrize("F", [602, 128, 100, 41, 7, 1, 1433, 64])
@pytest.mark.parametrize("slabs", [1, 3, 16])
def test_plan_matches_oracle(F, slabs):
    rng = np.random.default_rng(1000 + F + slabs)
    n_rows, n_src, n_edges = 700, 900, 40000
    off, idx, w = make_graph(rng, n_rows, n_src, n_edges)
    base = 5000
    X = rng.uniform(-1, 1, (n_src, F)).astype(np.float32)
    plan, got = run_plan(off, idx + base, w, X, base, slabs, n_src)
    assert plan.slabs == slabs
    ref = oracle_c.segment_gather_sum(off, idx, w, X)
    row_close(got, ref)
    # segment sizes exactly: all-ones input, unit weights -> in-degree counts
    _, cnt = run_plan(off, idx + base, None, np.ones((n_src, 4), dtype=np.float32), base, slabs, n_src)
    assert np.array_equal(cnt[:, 0], np.diff(off.astype(np.int64)).astype(np.float32))


@pytest.mark.parametrize("F", [602, 128, 41, 1433])
def test_plan_tma_row_staging_variant_matches_oracle(F):
    """Variant 1 (feature rows staged in shared memory by per-row cp.async.bulk, the north star's TMA staging): same
    results as the oracle; kept for measurement, the default stays variant 0."""
    from neutronstarlite_b200 import _lib
    rng = np.random.default_rng(2000 + F)
    n_rows, n_src, n_edges = 500, 700, 30000
    off, idx, w = make_graph(rng, n_rows, n_src, n_edges)
    X = rng.uniform(-1, 1, (n_src, F)).astype(np.float32)
    ref = oracle_c.segment_gather_sum(off, idx, w, X)
    _lib.call("nts_gather_plan_set_variant", 1)
    try:
        for slabs in (1, 4):
            _, got = run_plan(off, idx, w, X, 0, slabs, n_src)
            row_close(got, ref)
    finally:
        _lib.call("nts_gather_plan_set_variant", 0)


def test_plan_slot_table_and_unaligned_views():
    """Indices through a slot table (the receive-staging slots of the exchange) and a feature matrix whose rows are
    16-byte multiples but whose base pointer is only 4-byte aligned (a view): the padded workspace must kick in."""
    rng = np.random.default_rng(77)
    n_rows, n_src, n_edges, F = 300, 500, 20000, 128
    off, idx, w = make_graph(rng, n_rows, n_src, n_edges)
    ids = rng.permutation(4000)[:n_src].astype(np.uint32)         # global ids
    slot_of = np.zeros(4000, dtype=np.uint32)
    slot_of[ids] = np.arange(n_src, dtype=np.uint32)
    X = rng.uniform(-1, 1, (n_src, F)).astype(np.float32)
    _, got = run_plan(off, ids[idx], w, X, 0, 4, n_src, slot_of=slot_of)
    ref = oracle_c.segment_gather_sum(off, idx, w, X)
    row_close(got, ref)
    from neutronstarlite_b200 import ops
    flat = torch.zeros(n_src * F + 1, dtype=torch.float32, device=dev())
    flat[1:] = torch.from_numpy(X).to(dev()).reshape(-1)
    xv = flat[1:].view(n_src, F)
    assert xv.data_ptr() % 16 != 0
    plan = ops.GatherPlan(up_u32(off), up_u32(idx), torch.from_numpy(w).to(dev()), 0, n_rows, idx.shape[0], n_src, 2)
    out = torch.zeros((n_rows, F), dtype=torch.float32, device=dev())
    plan.run(xv, out)
    torch.cuda.synchronize()
    row_close(out.cpu().numpy(), ref)


def test_plan_accumulates_and_handles_degenerate_inputs():
    from neutronstarlite_b200 import ops
    rng = np.random.default_rng(5)
    off, idx, w = make_graph(rng, 64, 64, 3000, empty_every=0)
    X = rng.uniform(-1, 1, (64, 12)).astype(np.float32)
    plan = ops.GatherPlan(up_u32(off), up_u32(idx), torch.from_numpy(w).to(dev()), 0, 64, idx.shape[0], 64, 2)
    out = torch.ones((64, 12), dtype=torch.float32, device=dev())
    plan.run(torch.from_numpy(X).to(dev()), out)
    plan.run(torch.from_numpy(X).to(dev()), out)       # accumulate semantics, like every aggregation entry
    torch.cuda.synchronize()
    ref = oracle_c.segment_gather_sum(off, idx, w, X)
    row_close(out.cpu().numpy(), 1.0 + 2.0 * ref, rtol=2e-4)
    # no edges at all
    empty = ops.GatherPlan(up_u32(np.zeros(9, dtype=np.uint32)), None, None, 0, 8, 0, 8, 4)
    z = torch.zeros((8, 12), dtype=torch.float32, device=dev())
    empty.run(torch.from_numpy(X[:8].copy()).to(dev()), z)
    torch.cuda.synchronize()
    assert float(z.abs().max()) == 0.0


def test_ops_use_plans_and_match_plain_kernel():
    """ops.gather_by_* switch to plans for large chunks ("auto"); same result as the plain kernel on the reference
    layout, forward and backward, F = 602 (padded) and F = 128."""
    from neutronstarlite_b200 import ops
    from neutronstarlite_b200.graph import HostGraph, PartitionedGraph
    rng = np.random.default_rng(11)
    V, E = 5000, 300000
    edges = np.stack([rng.integers(0, V, E), rng.integers(0, V, E)], 1).astype(np.uint32)
    edges[: E // 6, 1] = 17
    pg = PartitionedGraph(HostGraph(edges, V), 1, 0).generate_all(device=dev())
    c = pg.graph_chunks[0]
    for F in (602, 128):
        x = torch.from_numpy(rng.uniform(-1, 1, (V, F)).astype(np.float32)).to(dev())
        res = {}
        for mode, slabs in (("off", 0), ("on", 0), ("on", 5)):
            ops.set_plan_mode(mode, slabs)
            try:
                y = ops.gather_by_dst_from_src(c, torch.zeros_like(x), x)
                dx = ops.gather_by_src_from_dst(c, torch.zeros_like(x), x)
                torch.cuda.synchronize()
                res[(mode, slabs)] = (y.cpu().numpy(), dx.cpu().numpy())
            finally:
                ops.set_plan_mode("auto", 0)
        for k in (("on", 0), ("on", 5)):
            row_close(res[k][0], res[("off", 0)][0])
            row_close(res[k][1], res[("off", 0)][1])
        assert ("fwd", 5) in c._gather_plans and ("bwd", 5) in c._gather_plans
