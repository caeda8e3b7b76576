#!/bin/bash
# One 1-GPU gpurun call of round 2: GPU test suite, the headline bench, the GAT bench (config D), the TMA row-staging
# measurement, and the ncu evidence of the planned aggregation kernel.  Everything lands in gpurun_out/.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > $O/pytest_gpu_r2_n1.log
python bench.py --steps 20 --warmup 5 > $O/bench_r2_n1.json 2> $O/bench_r2_n1.err
python bench.py --steps 10 --warmup 3 --zipf-s 0 --no-cpu-baseline --no-ref-gpu > $O/bench_r2_n1_uniform.json 2> $O/bench_r2_n1_uniform.err
python bench.py --toolkit gat --steps 5 --warmup 3 > $O/bench_r2_gat.json 2> $O/bench_r2_gat.err
NTS_AGG_NO_SUBWARP=1 python bench.py --toolkit gat --steps 5 --warmup 3 --no-e2e > $O/bench_r2_gat_nosubwarp.json 2> $O/bench_r2_gat_nosubwarp.err
python tools/k1_sweep.py --quick --tma --slabs 1 --out $O/k1_sweep_r2_tma.jsonl > $O/k1_sweep_r2_tma.log 2>&1
python bench.py --workload products --steps 10 --warmup 3 --no-cpu-baseline --no-ref-gpu > $O/bench_r2_n1_products.json 2> $O/bench_r2_n1_products.err
# ncu: launch list of a short run, then a full capture of six consecutive planned-kernel launches in steady state
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 300 --csv --log-file $O/launches_r2_n1.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ref-gpu --no-e2e > $O/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:planned_gather_sum_kernel -s 60 -c 6 \
    -o $O/prof_r2_plan python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ref-gpu --no-e2e > $O/ncu_full.log 2>&1
tail -n 3 $O/pytest_gpu_r2_n1.log
for f in bench_r2_n1 bench_r2_n1_uniform bench_r2_gat bench_r2_gat_nosubwarp bench_r2_n1_products; do echo "== $f"; head -c 1200 $O/$f.json; echo; tail -n 2 $O/$f.err; done
