#!/bin/bash
# Third 1-GPU call: ncu --set full of the steady-state planned kernels (demangled names carry "(int)"), config D again.
cd "$(dirname "$0")/.."
O=gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ref-gpu --no-e2e"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:planned_gather_sum_kernel<\(int\)5' -s 23 -c 2 -o $O/prof_r2_plan_F602 $B > $O/ncu_full_F602.log 2>&1
timeout 900 ncu --set full --clock-control none --kernel-name-base demangled \
    -k 'regex:planned_gather_sum_kernel<\(int\)1' -s 120 -c 4 -o $O/prof_r2_plan_F128 $B > $O/ncu_full_F128.log 2>&1
timeout 900 ncu --set full --clock-control none --kernel-name-base demangled \
    -k 'regex:planned_gather_sum_kernel<\(int\)5' -s 60 -c 8 -o $O/prof_r2_plan_F602_uniform $B --zipf-s 0 > $O/ncu_full_F602_uniform.log 2>&1
grep -h "No kernels\|Report" $O/ncu_full_F602.log $O/ncu_full_F128.log $O/ncu_full_F602_uniform.log
python bench.py --toolkit gat --steps 5 --warmup 3 --no-e2e > $O/bench_r2_gat_b.json 2> $O/bench_r2_gat_b.err
head -c 300 $O/bench_r2_gat_b.json; tail -n 2 $O/bench_r2_gat_b.err
