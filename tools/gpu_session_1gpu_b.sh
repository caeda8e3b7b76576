#!/bin/bash
# Second 1-GPU gpurun call of round 2: full GPU test suite with durations, ncu evidence of the steady-state planned
# kernels (Zipf and uniform graph), the launch list of the timed region, compute-sanitizer, the reference arm.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q --durations=12 2>&1 | tail -40 > $O/pytest_gpu_r2_n1.log
tail -n 22 $O/pytest_gpu_r2_n1.log
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ref-gpu --no-e2e"
# F=602 launches: 21 belong to the slab-count measurement (1, 2, 4 slabs x 3 launches), the rest are steady state
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:planned_gather_sum_kernel<5' -s 23 -c 2 -o $O/prof_r2_plan_F602 $B > $O/ncu_full_F602.log 2>&1
timeout 900 ncu --set full --clock-control none --kernel-name-base demangled \
    -k 'regex:planned_gather_sum_kernel<1' -s 120 -c 4 -o $O/prof_r2_plan_F128 $B > $O/ncu_full_F128.log 2>&1
timeout 900 ncu --set full --clock-control none --kernel-name-base demangled \
    -k 'regex:planned_gather_sum_kernel<5' -s 60 -c 8 -o $O/prof_r2_plan_F602_uniform $B --zipf-s 0 > $O/ncu_full_F602_uniform.log 2>&1
timeout 600 ncu --nvtx --nvtx-include "nts_timed/" --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $O/launches_r2_n1_timed.csv $B > $O/ncu_launches_timed.log 2>&1
timeout 900 compute-sanitizer --tool memcheck python tools/sanitizer_smoke.py > $O/sanitizer_r2_memcheck.txt 2>&1
timeout 900 compute-sanitizer --tool racecheck python tools/sanitizer_smoke.py > $O/sanitizer_r2_racecheck.txt 2>&1
tail -n 4 $O/sanitizer_r2_memcheck.txt $O/sanitizer_r2_racecheck.txt
(time python bench.py --impl reference --steps 20 --warmup 5) > $O/bench_r2_reference_arm.json 2> $O/bench_r2_reference_arm.err
tail -c 700 $O/bench_r2_reference_arm.json; tail -n 4 $O/bench_r2_reference_arm.err
python bench.py --toolkit gcn_eager --steps 10 --warmup 3 --no-cpu-baseline --no-ref-gpu > $O/bench_r2_n1_eager.json 2> $O/bench_r2_n1_eager.err
head -c 400 $O/bench_r2_n1_eager.json
