#!/usr/bin/env python3
"""Config D of BASELINE.json: 3-layer GAT, 8 heads, on the Reddit-shaped synthetic graph, 1 GPU, through the fused
multi-head aggregation path (toolkits.GATImpl).  Prints one JSON line: ms per epoch (CUDA events, after warm-up),
aggregated edges/s (2 aggregations per layer per epoch: forward + backward), peak memory.

    python tools/run_gat.py [--workload reddit] [--heads 8] [--layers 602-64-64-41] [--steps 3] [--warmup 1]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from neutronstarlite_b200 import _lib, synth
from neutronstarlite_b200.exchange import GpuExchange
from neutronstarlite_b200.graph import PartitionedGraph
from neutronstarlite_b200.toolkits import GATImpl


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="reddit")
    ap.add_argument("--heads", type=int, default=8)
    ap.add_argument("--layers", default="602-64-64-41")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--fused-kernel", type=int, default=1, help="1: K7 DistGPUFusedGATOp, 0: [E,H] operator chain")
    ap.add_argument("--two-pass", type=int, default=1, help="K7 backward: 1 = dst-major + src-major passes without "
                    "per-edge atomics, 0 = single destination-major pass with atomics")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    V, E_rand, _ = synth.WORKLOADS[a.workload]
    layers = [int(x) for x in a.layers.split("-")]
    src, dst = synth.zipf_edges(V, E_rand, dev)
    E = int(src.numel())
    pg = PartitionedGraph.from_device_edges(src, dst, V)
    # whole-partition CSC + MirrorIndex for the edge operators (P = 1: chunk 0 is the whole partition)
    c = pg.graph_chunks[0]
    pg.owned_vertices, pg.owned_edges = V, c.edge_size
    pg.column_offset_gpu, pg.row_indices_gpu = c.column_offset_gpu, c.row_indices_gpu
    has_src = torch.zeros(V + 1, dtype=torch.int32, device=dev)
    has_src[1:][torch.unique(src)] = 1
    pg.mirror_index_gpu = torch.cumsum(has_src, 0).to(torch.int32)
    pg.owned_mirrors = int(pg.mirror_index_gpu[-1].item())
    del src, dst, has_src
    torch.cuda.empty_cache()
    feats, labels, mask = synth.features_labels_mask(V, layers[0], layers[-1], dev)
    model = GATImpl(pg, layers, feats, labels, mask, heads=a.heads, exchange=GpuExchange(pg),
                    fused_kernel=bool(a.fused_kernel), two_pass_backward=bool(a.two_pass))
    for _ in range(a.warmup):
        model.run_epoch()
    torch.cuda.synchronize()
    # per-entry-point device time (CUDA events around every C-ABI call; nothing is synchronised until the end)
    records = []
    raw_call = _lib.call

    def timed_call(name, *args):
        e_a, e_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e_a.record()
        raw_call(name, *args)
        e_b.record()
        records.append((name, e_a, e_b))
    _lib.call = timed_call
    l0 = _lib.load().nts_kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        loss = model.run_epoch()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    _lib.call = raw_call
    per_call = {}
    for name, e_a, e_b in records:
        per_call[name] = per_call.get(name, 0.0) + e_a.elapsed_time(e_b) / a.steps
    n_layers = len(layers) - 1
    print(json.dumps({
        "workload": "%s-shaped, %d V, %d E, %d-layer GAT %s, %d heads (hidden layers), %s" % (
            a.workload, V, E, n_layers, a.layers, a.heads,
            "K7 fully fused attention+aggregation" if a.fused_kernel else "[E,H] operator chain + fused aggregation"),
        "ms_per_epoch": ms, "epochs_per_sec": 1e3 / ms,
        "aggregated_edges_per_sec": 2 * n_layers * E / (ms * 1e-3),
        "ms_per_epoch_by_entry_point": {k: round(v, 3) for k, v in sorted(per_call.items(), key=lambda kv: -kv[1])},
        "loss": float(loss.item()), "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9,
        "gpu_launches_per_epoch": (_lib.load().nts_kernel_launch_count() - l0) / a.steps}))


if __name__ == "__main__":
    main()
