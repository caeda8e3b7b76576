"""Host-side pieces of bench.py that the driver's records depend on: both arms print the SAME `config` dict, the CPU
arm's scale model keeps the workload's mean degree, and the thread count honours the affinity mask."""
import argparse
import importlib.util
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("nts_bench", os.path.join(ROOT, "bench.py"))
bench = importlib.util.module_from_spec(spec)
spec.loader.exec_module(bench)


def _args(**kw):
    d = dict(workload="reddit", toolkit="gcn", drop_rate=0.0, zipf_s=1.0, heads=8)
    d.update(kw)
    return argparse.Namespace(**d)


def test_both_arms_describe_the_same_config():
    from neutronstarlite_b200 import synth
    V, E_rand, layers = synth.WORKLOADS["reddit"]
    a = bench._config(_args(), V, E_rand + V, layers)
    b = bench._config(_args(), V, E_rand + V, layers)   # what `--impl reference` prints
    assert a == b and a["aggregations_per_epoch"] == 3 and "602-128-41" in a["workload"]
    assert "inputs larger than L2" in a["l2"]
    g = bench._config(_args(toolkit="gat"), V, E_rand + V, [602, 64, 64, 41])
    assert g["aggregations_per_epoch"] == 6 and "GAT, 8 heads" in g["workload"]


def test_scale_model_keeps_the_mean_degree():
    V, E = 20000, 400000
    Vs, edges = bench._scale_model(V, E, 4)
    assert Vs == V // 4 and edges.dtype == np.uint32 and edges.shape[1] == 2
    assert edges.shape[0] == E // 4 + Vs                       # + one self loop per vertex
    assert abs(edges.shape[0] / Vs - (E + V) / V) < 1.0         # same mean degree as the workload
    assert int(edges.max()) < Vs
    div, probe = bench.pick_cpu_sample(V, E, [8, 4, 2], 16, 10.0, 25, 1)   # explicit div: no probe run
    assert (div, probe) == (16, None)
    assert "1/16 scale model" in bench._sample_text(16, Vs, 123, None)
    assert "workload itself" in bench._sample_text(1, V, E, None)


def test_usable_cores_is_bounded_by_the_affinity_mask():
    n = bench.usable_cores()
    assert 1 <= n <= len(os.sched_getaffinity(0))
