#!/bin/bash
# ncu evidence for the bench command: launch list + one full capture of the chain kernel.
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-configs --no-sustained --no-ablation ${BENCH_ARGS:-}"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:${KREGEX:-tmm_cavity|tmm_typed} -s 3 -c 1 -f -o gpurun_out/prof $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_list.log; tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out
