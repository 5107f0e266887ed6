#!/bin/bash
# full ncu captures of the chain kernel: typed (default) and untyped (--no-layer-cse) variants
mkdir -p gpurun_out
A="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$A > gpurun_out/plainA.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tmm_chain -s 3 -c 1 -f -o gpurun_out/prof_typed $A > gpurun_out/ncuA.log 2>&1
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-layer-cse"
$B > gpurun_out/plainB.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tmm_chain -s 3 -c 1 -f -o gpurun_out/prof_untyped $B > gpurun_out/ncuB.log 2>&1
C="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload c4"
$C > gpurun_out/plainC.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tmm_chain -s 3 -c 1 -f -o gpurun_out/prof_c4 $C > gpurun_out/ncuC.log 2>&1
tail -1 gpurun_out/ncuA.log gpurun_out/ncuB.log gpurun_out/ncuC.log
