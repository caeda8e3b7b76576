"""Synthetic inputs of the shapes BASELINE.json names: power-law multigraphs generated on the GPU.

Edges: both endpoints drawn from a Zipf(s) law over a random permutation of the vertices (hubs on both sides,
duplicates kept - they count in degrees like in the reference), plus one self loop per vertex.  Features U(-1,1),
labels U{0..C-1}, mask v % 3 (the reference's `random_generate`, core/ntsDataloador.hpp:63-71, uses all-ones
features which would hide index bugs)."""
from __future__ import annotations

import torch

WORKLOADS = {
    # name: (V, E_random, layers)            SURVEY.md 8 preamble / 8d
    "reddit": (232965, 114615892, [602, 128, 41]),
    "products": (2449029, 61859140, [100, 128, 47]),
    "cora_sized": (2708, 10858, [1433, 128, 7]),
    "tiny": (20000, 400000, [602, 128, 41]),
}

SEED_GRAPH = 0x5EED0001
SEED_FEATURES = 0x5EED0002


def zipf_edges(V, E, device, s=1.0, seed=SEED_GRAPH, self_loops=True, chunk=1 << 26):
    """int64 (src, dst) on `device`; deterministic for a given (V, E, s, seed) and GPU architecture."""
    gen = torch.Generator(device=device).manual_seed(seed)
    w = torch.arange(1, V + 1, device=device, dtype=torch.float64).pow_(-s)
    cdf = torch.cumsum(w / w.sum(), 0)
    perm_s = torch.randperm(V, generator=gen, device=device)
    perm_d = torch.randperm(V, generator=gen, device=device)
    src_parts, dst_parts = [], []
    done = 0
    while done < E:
        n = min(chunk, E - done)
        u = torch.rand(n, generator=gen, device=device, dtype=torch.float64)
        src_parts.append(perm_s[torch.searchsorted(cdf, u).clamp_(max=V - 1)])
        u = torch.rand(n, generator=gen, device=device, dtype=torch.float64)
        dst_parts.append(perm_d[torch.searchsorted(cdf, u).clamp_(max=V - 1)])
        done += n
    if self_loops:
        loops = torch.arange(V, device=device, dtype=torch.int64)
        src_parts.append(loops)
        dst_parts.append(loops)
    return torch.cat(src_parts), torch.cat(dst_parts)


def features_labels_mask(V, F, classes, device, seed=SEED_FEATURES, rows=None):
    """Features of rows [rows[0], rows[1]) (default all), generated per row block so every rank of a partitioned
    run sees the same global matrix."""
    lo, hi = (0, V) if rows is None else rows
    gen = torch.Generator(device=device).manual_seed(seed)
    block = 1 << 15
    out = torch.empty((hi - lo, F), dtype=torch.float32, device=device)
    labels = torch.empty(hi - lo, dtype=torch.int64, device=device)
    # generate whole blocks and keep the overlap: deterministic per global row id
    b0 = (lo // block) * block
    gen_state_skip = b0 // block
    for _ in range(gen_state_skip):
        torch.rand((block, F), generator=gen, device=device)
        torch.randint(0, classes, (block,), generator=gen, device=device)
    pos = b0
    while pos < hi:
        x = torch.rand((block, F), generator=gen, device=device) * 2 - 1
        y = torch.randint(0, classes, (block,), generator=gen, device=device)
        a, b = max(pos, lo), min(pos + block, hi)
        if b > a:
            out[a - lo:b - lo] = x[a - pos:b - pos]
            labels[a - lo:b - lo] = y[a - pos:b - pos]
        pos += block
    mask = (torch.arange(lo, hi, device=device) % 3).to(torch.int64)
    return out, labels, mask
