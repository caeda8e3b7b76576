#!/usr/bin/env python3
"""Diagnostic: where does an end-to-end (host-fed) step of a toolkit spend its time?  Times, per step, the pinned
H2D copy alone, the epoch with swapped input buffers but no copy, and both together (device events + host clock)."""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from neutronstarlite_b200 import synth
from neutronstarlite_b200.graph import PartitionedGraph
from neutronstarlite_b200.toolkits import GCNEagerImpl, GCNImpl


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--toolkit", default="gcn_eager")
    ap.add_argument("--div", type=int, default=4)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    V, E_rand, layers = synth.WORKLOADS["reddit"]
    src, dst = synth.zipf_edges(V, E_rand // a.div, dev)
    pg = PartitionedGraph.from_device_edges(src, dst, V)
    del src, dst
    feats, labels, mask = synth.features_labels_mask(V, layers[0], layers[-1], dev)
    eager = a.toolkit == "gcn_eager"
    model = (GCNEagerImpl if eager else GCNImpl)(pg, layers, feats, labels, mask, drop_rate=0.5)
    host = torch.empty(feats.shape, dtype=torch.float32).pin_memory()
    host.copy_(feats.detach())   # not feats itself: that would tie `host` (and every buffer filled from it) into autograd
    bufs = [torch.empty_like(feats), torch.empty_like(feats)]
    cs = torch.cuda.Stream(device=dev)
    res = {"toolkit": a.toolkit, "host_is_pinned": host.is_pinned()}

    def timed(name, fn, n=5):
        for _ in range(2):
            fn(0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for k in range(n):
            fn(k)
        e1.record()
        t_issue = time.perf_counter() - t0
        torch.cuda.synchronize()
        res[name] = {"device_ms_per_step": e0.elapsed_time(e1) / n, "host_issue_ms_per_step": 1e3 * t_issue / n,
                     "host_wall_ms_per_step": 1e3 * (time.perf_counter() - t0) / n}

    def epoch_resident(k):
        model.run_epoch()

    def copy_only(k):
        with torch.no_grad():
            bufs[k & 1].copy_(host, non_blocking=True)

    def copy_side_stream(k):
        with torch.no_grad(), torch.cuda.stream(cs):
            bufs[k & 1].copy_(host, non_blocking=True)
        torch.cuda.current_stream().wait_stream(cs)

    def epoch_swapped(k):
        model.X[0] = bufs[k & 1] if eager else bufs[k & 1].requires_grad_(True)
        model.run_epoch()

    def epoch_with_copy(k):
        with torch.no_grad(), torch.cuda.stream(cs):
            bufs[(k + 1) & 1].copy_(host, non_blocking=True)
        model.X[0] = bufs[k & 1] if eager else bufs[k & 1].requires_grad_(True)
        model.run_epoch()
        torch.cuda.current_stream().wait_stream(cs)

    timed("epoch_resident", epoch_resident)
    timed("copy_only_main_stream", copy_only)
    timed("copy_only_side_stream", copy_side_stream)
    try:
        timed("epoch_swapped_buffers", epoch_swapped)
    except Exception as exc:
        res["epoch_swapped_buffers"] = repr(exc)
    try:
        timed("epoch_with_overlapped_copy", epoch_with_copy)
    except Exception as exc:
        res["epoch_with_overlapped_copy"] = repr(exc)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
