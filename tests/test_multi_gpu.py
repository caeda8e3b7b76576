"""CUDA data path of the distributed fused aggregation on 2 (or 4) GPUs of one box, both transports ("nccl"
all-to-all and "p2p" peer-memory pull), against the golden vectors of the reference run at the same P.
Skipped when fewer GPUs are visible (the round-end single-GPU run); run with `gpurun --gpus 2 -- pytest -m gpu
tests/test_multi_gpu.py`."""
import os
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def _worker(rank, world, port, case, transport, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from neutronstarlite_b200 import ops
        from neutronstarlite_b200.exchange import GpuExchange
        from neutronstarlite_b200.graph import HostGraph, PartitionedGraph
        z = np.load(os.path.join(GOLD, case))
        V, E, P, F = (int(x) for x in z["case"])
        pg = PartitionedGraph(HostGraph(z["edges"], V), P, rank).generate_all(device=dev, dist=True)
        ex = GpuExchange(pg, transport=transport)
        op = ops.ForwardGPUfuseOp(pg, None, exchange=ex)
        x = torch.from_numpy(z["r%d/X" % rank].reshape(-1, F)).to(dev)
        g = torch.from_numpy(z["r%d/G" % rank].reshape(-1, F)).to(dev)
        ref_y = z["r%d/gcn_Y" % rank].reshape(-1, F)
        ref_dx = z["r%d/gcn_dX" % rank].reshape(-1, F)
        for it in range(3):  # repeated calls exercise buffer reuse and the p2p epoch protocol
            y = op.forward(x)
            dx = op.backward(g)
            torch.cuda.synchronize()
            np.testing.assert_allclose(y.cpu().numpy(), ref_y, rtol=1e-4, atol=1e-5)
            np.testing.assert_allclose(dx.cpu().numpy(), ref_dx, rtol=1e-4, atol=2e-5)
        # mirror fetch / return (DistGPUGetDepNbrOp) against the reference's DistGetDepNbrOp at the same P
        dep = ops.DistGPUGetDepNbrOp(pg, None, exchange=ex)
        mirror = dep.forward(x)
        torch.cuda.synchronize()
        assert np.array_equal(mirror.cpu().numpy(), z["r%d/dep_mirror" % rank].reshape(-1, F))
        gm = torch.from_numpy(z["r%d/dep_Gm" % rank].reshape(-1, F)).to(dev)
        dxm = dep.backward(gm)
        torch.cuda.synchronize()
        np.testing.assert_allclose(dxm.cpu().numpy(), z["r%d/dep_dX" % rank].reshape(-1, F), rtol=1e-4, atol=2e-5)
        # a wider feature matrix (second width through the same exchange object)
        F2 = 40
        gen = torch.Generator().manual_seed(1)
        Xg = torch.rand((V, F2), generator=gen) * 2 - 1
        po = pg.partition_offset
        y2 = op.forward(Xg[int(po[rank]):int(po[rank + 1])].contiguous().to(dev))
        torch.cuda.synchronize()
        # single-partition truth with the same kernels on this GPU
        pg1 = PartitionedGraph(HostGraph(z["edges"], V), 1, 0).generate_all(device=dev)
        y_full = ops.ForwardSingleGPUfuseOp(pg1).forward(Xg.to(dev))
        torch.testing.assert_close(y2, y_full[int(po[rank]):int(po[rank + 1])], rtol=1e-4, atol=1e-5)
        dist.barrier()
        q.put((rank, "ok"))
    except Exception as exc:  # pragma: no cover
        import traceback
        q.put((rank, "FAIL: %r\n%s" % (exc, traceback.format_exc())))
    finally:
        dist.destroy_process_group()


CASES = [("synth9k_P2_F2.npz", 2), ("cora_self_P2_F4.npz", 2), ("synth9k_P4_F2.npz", 4), ("cora_self_P4_F2.npz", 4)]


@pytest.mark.parametrize("transport", ["nccl", "p2p"])
@pytest.mark.parametrize("case,world", CASES)
def test_distributed_fused_aggregation(case, world, transport):
    if not torch.cuda.is_available() or torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    port = 29700 + (hash((case, transport)) % 200)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, case, transport, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = []
    try:
        for _ in range(world):
            results.append(q.get(timeout=300))
    finally:
        for p in procs:
            p.join(timeout=30)
            if p.is_alive():
                p.kill()
    for rank, msg in sorted(results):
        assert msg == "ok", "rank %d: %s" % (rank, msg)
