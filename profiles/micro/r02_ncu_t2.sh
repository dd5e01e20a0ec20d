#!/bin/bash
# Round-2 (second half) profile captures of the two-step sweep after the downward-march / pillar changes (one GPU).
# Each program is first run WITHOUT ncu and must exit 0.
set -e
B="python bench.py --steps 2 --warmup 1 --no-configs --no-cpu-baseline --e2e-steps 0"
$B > gpurun_out/r02_bench_short.json 2> gpurun_out/r02_bench_short.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv $B > gpurun_out/r02_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_heat_t2 -s 6 -c 1 -o gpurun_out/r02_t2 -f $B > gpurun_out/r02_ncu_t2.log 2>&1
C="python bench.py --config C3 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0"
$C > gpurun_out/r02_bench_c3_short.json 2> gpurun_out/r02_bench_c3_short.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_c3.csv $C > gpurun_out/r02_ncu_list_c3.log 2>&1
echo captures done
