#!/bin/bash
# one full ncu capture of the chain kernel for `bench.py $BENCH_ARGS`
mkdir -p gpurun_out
A="python bench.py --steps 3 --warmup 3 --no-cpu-baseline ${BENCH_ARGS:-}"
$A > gpurun_out/plainA.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tmm_chain -s 3 -c 1 -f -o gpurun_out/${OUT:-prof} $A > gpurun_out/ncuA.log 2>&1
tail -n 2 gpurun_out/ncuA.log
