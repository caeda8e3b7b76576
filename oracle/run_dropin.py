#!/usr/bin/env python3
"""TEST INFRASTRUCTURE ONLY - run the reference's OWN host code on our kernels (the drop-in proof).

`oracle/_ref/nts_dropin_main` is the reference's `toolkits/main.cpp` (every GPU toolkit: GCN.hpp, GCN_EAGER*.hpp,
GAT_GPU_DIST.hpp, GIN_GPU.hpp, COMMNET_GPU.hpp, test_getdepneighbor_gpu.hpp) compiled unchanged with -DCUDA_ENABLE=1
against include/nts_dropin (our shadow of cuda/ntsCUDA.hpp) and linked with libnts_b200.so instead of the reference's
cuda_propagate library (oracle/Makefile target `dropin`).  This script runs it on the reference's Cora fixture for a
list of ALGORITHM values on the GPU box (single rank: the MPI stand-in is in-process) and, for comparison, the CPU
reference binary on the same cfg.  Prints one JSON object per algorithm with the parsed per-epoch losses and final
accuracies.

    python oracle/run_dropin.py [--epochs 30] [--algos GCNEAGERSINGLE,GCN,GATGPUDIST,test_getdep]
"""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")

CFG = """ALGORITHM:{algo}
VERTICES:2708
LAYERS:{layers}
EPOCHS:{epochs}
EDGE_FILE:{data}/cora.2708.edge.self
FEATURE_FILE:{data}/cora.featuretable
LABEL_FILE:{data}/cora.labeltable
MASK_FILE:{data}/cora.mask
PROC_OVERLAP:0
PROC_LOCAL:0
PROC_CUDA:1
PROC_REP:0
LOCK_FREE:1
LEARN_RATE:0.01
WEIGHT_DECAY:0.0001
DECAY_RATE:0.97
DECAY_EPOCH:100
DROP_RATE:0.0
"""


def _visible_devices(rank, n_gpus):
    """The reference never calls cudaSetDevice (always device 0, core/NtsScheduler.hpp:299), so rank r gets physical
    GPU r % n_gpus AS its device 0 - but every other GPU stays visible behind it: CUDA IPC can only map a peer's
    window if the exporting device is visible to the importing process."""
    first = rank % n_gpus
    return ",".join(str((first + k) % n_gpus) for k in range(n_gpus))


def _run_ranks(binary, cfg, nprocs, env, timeout):
    """rank 0's output and the first non-zero exit status of `nprocs` ranks under the MPI stand-in
    (oracle/shim/mpi.h: NTS_SHIM_SIZE / NTS_SHIM_RANK / NTS_SHIM_DIR).  The reference never calls cudaSetDevice
    (device 0 always), so rank r is pinned to GPU r % n_gpus through CUDA_VISIBLE_DEVICES; with one GPU all ranks
    share it - the host-side exchange of the reference does not care."""
    if nprocs == 1:
        p = subprocess.run([binary, cfg], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env,
                           timeout=timeout)
        return p.stdout, p.returncode
    import shutil
    try:
        import torch
        n_gpus = max(1, torch.cuda.device_count())
    except Exception:
        n_gpus = 1
    scratch = tempfile.mkdtemp(prefix="nts_shim_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    procs = []
    try:
        for r in range(nprocs):
            e = dict(env, NTS_SHIM_SIZE=str(nprocs), NTS_SHIM_RANK=str(r), NTS_SHIM_DIR=scratch,
                     CUDA_VISIBLE_DEVICES=_visible_devices(r, n_gpus))
            procs.append(subprocess.Popen([binary, cfg], env=e, text=True,
                                          stdout=subprocess.PIPE if r == 0 else subprocess.DEVNULL,
                                          stderr=subprocess.STDOUT if r == 0 else subprocess.DEVNULL))
        out, _ = procs[0].communicate(timeout=timeout)
        rc = procs[0].returncode
        for q in procs[1:]:
            q.wait(timeout=60)
            rc = rc or q.returncode
        return out, rc
    finally:
        for q in procs:
            if q.poll() is None:
                q.kill()
        shutil.rmtree(scratch, ignore_errors=True)


def run(binary, algo, epochs, layers="1433-128-7", timeout=600, nprocs=1):
    with tempfile.TemporaryDirectory() as d:
        cfg = os.path.join(d, "c.cfg")
        open(cfg, "w").write(CFG.format(algo=algo, layers=layers, epochs=epochs, data=os.path.join(REF, "data")))
        env = dict(os.environ)
        env.setdefault("NTS_THREADS", "8")
        env["OMP_NUM_THREADS"] = env["NTS_THREADS"]
        out, rc = _run_ranks(binary, cfg, nprocs, env, timeout)

    class p:   # keep the field names used below
        returncode = rc
    losses = [float(x) for x in re.findall(r"Epoch\[\d+\]:loss\s+([-0-9.eE+]+)", out)]
    accs = re.findall(r"(Train|Eval|Test)\s+ACC:\s+([0-9.]+)", out)
    last = {}
    for k, v in accs:
        last[k.lower()] = float(v)
    passed = re.findall(r"(\d+) is passed|passed", out)
    return {"algo": algo, "ranks": nprocs, "rc": p.returncode, "epochs": len(losses), "loss_first": losses[0] if losses else None,
            "loss_last": losses[-1] if losses else None, "acc": last, "tail": out[-600:] if p.returncode else "",
            "raw_passed_lines": [ln for ln in out.splitlines() if "pass" in ln.lower()][:12]}


SYN_CFG = """ALGORITHM:{algo}
VERTICES:{V}
LAYERS:{layers}
EPOCHS:{epochs}
EDGE_FILE:{edge}
FEATURE_FILE:random
LABEL_FILE:random
MASK_FILE:random
PROC_OVERLAP:0
PROC_LOCAL:0
PROC_CUDA:1
PROC_REP:0
LOCK_FREE:1
LEARN_RATE:0.01
WEIGHT_DECAY:0.0001
DECAY_RATE:0.97
DECAY_EPOCH:100
DROP_RATE:0.0
"""


def run_synthetic(binary, algo, V, layers, edge_file, epochs, timeout=1500, nprocs=1):
    """Per-epoch seconds of a reference toolkit on a synthetic edge file: from its own `Times[...(s)]` print when the
    toolkit has one (GCN_EAGER_single.hpp:250-252), else from the arrival times of its per-epoch loss lines."""
    import time
    with tempfile.TemporaryDirectory() as d:
        cfg = os.path.join(d, "c.cfg")
        open(cfg, "w").write(SYN_CFG.format(algo=algo, V=V, layers=layers, epochs=epochs, edge=edge_file))
        env = dict(os.environ)
        env.setdefault("NTS_THREADS", str(os.cpu_count()))
        env["OMP_NUM_THREADS"] = env["NTS_THREADS"]
        t0 = time.perf_counter()
        others, scratch = [], None
        if nprocs > 1:   # ranks 1.. in the background, rank 0 is the one we read
            import shutil
            import torch
            n_gpus = max(1, torch.cuda.device_count())
            scratch = tempfile.mkdtemp(prefix="nts_shim_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
            env = dict(env, NTS_SHIM_SIZE=str(nprocs), NTS_SHIM_DIR=scratch,
                       NTS_THREADS=str(max(1, int(env["NTS_THREADS"]) // nprocs)))
            env["OMP_NUM_THREADS"] = env["NTS_THREADS"]
            for r in range(1, nprocs):
                others.append(subprocess.Popen([binary, cfg], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL,
                                               env=dict(env, NTS_SHIM_RANK=str(r),
                                                        CUDA_VISIBLE_DEVICES=_visible_devices(r, n_gpus))))
            env = dict(env, NTS_SHIM_RANK="0", CUDA_VISIBLE_DEVICES=_visible_devices(0, n_gpus))
        p = subprocess.Popen([binary, cfg], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
        stamps, own = [], []
        for line in p.stdout:
            if "Running.Epoch[" in line:
                stamps.append(time.perf_counter())
                m = re.search(r"Times\[([0-9.eE+-]+)\(s\)\]", line)
                if m:
                    own.append(float(m.group(1)))
        p.wait(timeout=timeout)
        total = time.perf_counter() - t0
        for q in others:
            try:
                q.wait(timeout=120)
            except subprocess.TimeoutExpired:
                q.kill()
        if scratch:
            import shutil
            shutil.rmtree(scratch, ignore_errors=True)
    if own:
        per = own[1:] if len(own) > 1 else own
    else:
        per = [b - a for a, b in zip(stamps[:-1], stamps[1:])]
    return {"algo": algo, "ranks": nprocs, "rc": p.returncode, "epochs_seen": len(stamps), "s_per_epoch_after_first": per,
            "s_per_epoch_median": sorted(per)[len(per) // 2] if per else None, "wall_s_incl_load": total}


def synthetic_dist_main(div, epochs, nprocs):
    """The reference's own host code (toolkits/main.cpp + toolkits/GCN.hpp, unchanged) at P ranks, ONE GPU PER RANK,
    with the drop-in ForwardGPUfuseOp on the peer-memory exchange (nts_dropin_dist_main), on 1/div of the
    Reddit-shaped graph.  Every rank loads the edge file through the MPI stand-in like the reference does."""
    import numpy as np
    import torch
    sys.path.insert(0, os.path.dirname(HERE))
    from neutronstarlite_b200 import synth
    V, E_rand, layers = synth.WORKLOADS["reddit"]
    src, dst = synth.zipf_edges(V, E_rand // div, torch.device("cpu"))
    with tempfile.TemporaryDirectory() as d:
        efile = os.path.join(d, "syn.edge")
        torch.stack([src, dst], 1).numpy().astype(np.uint32).tofile(efile)
        res = {"graph": "reddit-shaped 1/%d: %d V, %d E, LAYERS %s" % (div, V, int(src.numel()), "-".join(map(str, layers))),
               "gpus": torch.cuda.device_count(), "ranks": nprocs}
        res["dropin_dist_GCN"] = run_synthetic(os.path.join(REF, "nts_dropin_dist_main"), "GCN", V,
                                               "-".join(map(str, layers)), efile, epochs, nprocs=nprocs)
    print(json.dumps(res, indent=1))
    return 0


def synthetic_main(div, epochs):
    """Reference host code + OUR kernels (nts_dropin_main) next to reference host code + ITS OWN kernels
    (nts_refgpu_main) on a Reddit-shaped synthetic edge file of 1/div of the edges."""
    import numpy as np
    import torch
    sys.path.insert(0, os.path.dirname(HERE))
    from neutronstarlite_b200 import synth
    V, E_rand, layers = synth.WORKLOADS["reddit"]
    src, dst = synth.zipf_edges(V, E_rand // div, torch.device("cpu"))
    with tempfile.TemporaryDirectory() as d:
        efile = os.path.join(d, "syn.edge")
        torch.stack([src, dst], 1).numpy().astype(np.uint32).tofile(efile)
        res = {"graph": "reddit-shaped 1/%d: %d V, %d E, LAYERS %s" % (div, V, int(src.numel()), "-".join(map(str, layers)))}
        for name, binary in (("ours", os.path.join(REF, "nts_dropin_main")), ("reference_kernels", os.path.join(REF, "nts_refgpu_main"))):
            if not os.path.exists(binary):
                continue
            for algo in ("GCNEAGERSINGLE", "GCN"):
                res["%s_%s" % (name, algo)] = run_synthetic(binary, algo, V, "-".join(map(str, layers)), efile, epochs)
    print(json.dumps(res, indent=1))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--epochs", type=int, default=30)
    ap.add_argument("--algos", default="GCNEAGERSINGLE,GCN,GATGPUDIST,test_getdep")
    ap.add_argument("-np", type=int, default=1, help="ranks of the reference's host code (MPI stand-in); P > 1 runs "
                    "its sync_compute_decoupled / compute_sync_decoupled exchange on top of our kernels")
    ap.add_argument("--no-cpu-reference", action="store_true", help="skip the CPU reference run of the same cfg")
    ap.add_argument("--dist-exchange", action="store_true",
                    help="use nts_dropin_dist_main (make -C oracle dropin_dist): ForwardGPUfuseOp on the device-resident "
                         "peer-memory exchange instead of the reference's host-staged MPI exchange")
    ap.add_argument("--synthetic", type=int, default=0, metavar="DIV",
                    help="compare ours vs the reference's own kernels through the reference's host code on 1/DIV of "
                         "the Reddit-shaped graph")
    a = ap.parse_args()
    if a.synthetic and a.dist_exchange:
        return synthetic_dist_main(a.synthetic, min(a.epochs, 8), a.np)
    if a.synthetic:
        return synthetic_main(a.synthetic, min(a.epochs, 6))
    res = {}
    cpu = os.path.join(REF, "nts_ref_main")
    gpu = os.path.join(REF, "nts_dropin_dist_main" if a.dist_exchange else "nts_dropin_main")
    res["binary"] = os.path.basename(gpu)
    if not a.no_cpu_reference:
        try:
            res["cpu_reference_GCNCPU"] = run(cpu, "GCNCPU", a.epochs, nprocs=a.np)
        except subprocess.TimeoutExpired:
            res["cpu_reference_GCNCPU"] = {"algo": "GCNCPU", "rc": "timeout"}
    for algo in a.algos.split(","):
        try:
            res["dropin_" + algo] = run(gpu, algo, a.epochs, nprocs=a.np, timeout=240)
        except subprocess.TimeoutExpired:
            res["dropin_" + algo] = {"algo": algo, "rc": "timeout"}
    print(json.dumps(res, indent=1))
    return 0


if __name__ == "__main__":
    sys.exit(main())
