#!/usr/bin/env python3
"""Kernel-level sweep of the aggregation kernel on the Reddit-shaped graph (one graph build, many configurations).
Prints one JSON line per configuration: avg ms of the F=602 and F=128 forward launches (CUDA events, 5 launches
after 2 warm-ups).  Usage: python tools/tune_aggregate.py [--zipf-s 1.0] [--configs "v,Q,U,B;..."]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from neutronstarlite_b200 import _lib, ops, synth
from neutronstarlite_b200.graph import PartitionedGraph


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--zipf-s", type=float, default=1.0)
    ap.add_argument("--workload", default="reddit")
    # variant, Q, U, MINB, tiles, tile_major   (0 = library default)
    ap.add_argument("--configs", default="1,0,0,0,0,0;2,0,0,0,0,0;2,256,0,0,0,0;2,1024,0,0,0,0;"
                                         "2,0,4,1,2,0;2,0,2,2,2,0;2,0,4,2,2,0;2,0,2,3,2,1;"
                                         "2,0,2,3,3,1;2,0,4,2,3,1;2,0,2,3,3,0;"
                                         "2,0,2,3,4,1;2,0,4,3,4,1;2,0,4,2,4,1;2,0,4,3,4,0;"
                                         "2,0,4,3,5,1;2,0,4,4,5,1;2,0,8,2,5,1;2,0,8,3,5,1;2,0,8,3,5,0;"
                                         "2,0,8,4,10,1;2,0,8,3,10,1;2,0,16,2,10,1;"
                                         "2,0,8,1,1,0;2,0,8,4,1,0;2,0,16,2,1,0;2,0,4,4,1,0")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    V, E_rand, layers = synth.WORKLOADS[a.workload]
    src, dst = synth.zipf_edges(V, E_rand, dev, s=a.zipf_s)
    pg = PartitionedGraph.from_device_edges(src, dst, V)
    del src, dst
    c = pg.graph_chunks[0]
    xs = {F: torch.rand((V, F), device=dev) * 2 - 1 for F in (layers[0], layers[1])}
    ys = {F: torch.zeros((V, F), device=dev) for F in xs}
    for cfg in a.configs.split(";"):
        v, q, u, b, t, m = (int(x) for x in cfg.split(","))
        _lib.call("nts_aggregate_set_variant", v, q)
        res = {"variant": v, "Q": q, "U": u, "minb": b, "tiles": t, "tile_major": m}
        for F in xs:
            # (U, MINB) and tile settings are per shape: tiles > 1 only make sense for the wide matrix
            wide = F == layers[0]
            os.environ.pop("NTS_AGG_TUNE", None)
            os.environ.pop("NTS_AGG_TILES", None)
            if u and (wide == (t != 1)):
                os.environ["NTS_AGG_TUNE"] = "%d,%d" % (u, b)
            if t and (wide == (t != 1)):
                os.environ["NTS_AGG_TILES"] = "%d,%d" % (t, m)
            try:
                ops.gather_by_dst_from_src(c, ys[F], xs[F])
            except Exception as exc:
                res["F%d_ms" % F] = "n/a: %s" % str(exc)[-60:]
                continue
            for _ in range(2):
                ops.gather_by_dst_from_src(c, ys[F], xs[F])
            torch.cuda.synchronize()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            for _ in range(5):
                ops.gather_by_dst_from_src(c, ys[F], xs[F])
            a1.record()
            torch.cuda.synchronize()
            res["F%d_ms" % F] = a0.elapsed_time(a1) / 5
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
