#!/bin/bash
# short 4-GPU check of the per-peer DMA streams: multi-GPU tests + config B
cd "$(dirname "$0")/.."
O=gpurun_out
N=$(python -c "import torch;print(torch.cuda.device_count())")
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q -k "distributed_fused" 2>&1 | tail -4 > $O/pytest_multi_gpu_r2_n${N}_b.log
cat $O/pytest_multi_gpu_r2_n${N}_b.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 20 --warmup 5 --no-e2e > $O/bench_r2_reddit_n${N}_b.json 2> $O/bench_r2_reddit_n${N}_b.err
head -c 600 $O/bench_r2_reddit_n${N}_b.json; tail -n 3 $O/bench_r2_reddit_n${N}_b.err
