#!/bin/bash
# One N-GPU gpurun call of round 2 (N = visible GPUs): multi-GPU parity tests, the headline bench at N (parity block
# included), config C (products-shaped) and - at 8 GPUs - config E (papers100M-shaped); the reference's own host code on
# the drop-in exchange with one GPU per rank.  Everything lands in gpurun_out/.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
N=$(python -c "import torch;print(torch.cuda.device_count())")
run() {  # run <tag> <bench args...>
  tag=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus $N "$@" > $O/${tag}_n$N.json 2> $O/${tag}_n$N.err
  echo "== $tag N=$N rc=$?"; head -c 1500 $O/${tag}_n$N.json; echo; tail -n 3 $O/${tag}_n$N.err
}
if [ "$N" -le 4 ]; then
  timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q 2>&1 | tail -15 > $O/pytest_multi_gpu_r2_n$N.log
  cat $O/pytest_multi_gpu_r2_n$N.log
fi
run bench_r2_reddit --steps 20 --warmup 5
run bench_r2_products --workload products --steps 20 --warmup 5
if [ "$N" -le 2 ]; then
  run bench_r2_reddit_nccl --steps 10 --warmup 3 --transport nccl --no-e2e
fi
if [ "$N" -ge 8 ]; then
  NTS_EXCHANGE_BUFFERS=1 run bench_r2_papers100m --workload papers100m --steps 5 --warmup 3 --no-e2e
fi
if [ "$N" -le 4 ]; then
  timeout 600 python oracle/run_dropin.py --synthetic 8 --dist-exchange -np $N --epochs 6 > $O/dropin_dist_r2_n$N.json 2> $O/dropin_dist_r2_n$N.err
  echo "== dropin rc=$?"; cat $O/dropin_dist_r2_n$N.json | head -40
fi
