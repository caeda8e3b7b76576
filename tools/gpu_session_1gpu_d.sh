#!/bin/bash
# ncu --set full of the steady-state F=602 launch (slab count forced to the measured value 1: no candidate launches)
# and of the F=128 launches (forced to the measured 2 slabs)
cd "$(dirname "$0")/.."
O=gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ref-gpu --no-e2e"
NTS_PLAN_SLABS=1 timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:planned_gather_sum_kernel<\(int\)5' -s 2 -c 2 -o $O/prof_r2_plan_F602_steady $B > $O/ncu_full_F602_steady.log 2>&1
NTS_PLAN_SLABS=2 timeout 300 ncu --set full --clock-control none --kernel-name-base demangled \
    -k 'regex:planned_gather_sum_kernel<\(int\)1' -s 8 -c 4 -o $O/prof_r2_plan_F128_steady $B > $O/ncu_full_F128_steady.log 2>&1
grep -h "No kernels\|Report" $O/ncu_full_F602_steady.log $O/ncu_full_F128_steady.log
