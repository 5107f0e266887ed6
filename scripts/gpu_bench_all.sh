#!/bin/bash
# quick benches of all workloads + ablation (no CPU baseline)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
for w in c2 c4 c3 c5 c1; do python bench.py --steps ${STEPS:-30} --no-cpu-baseline --workload $w > gpurun_out/bench_$w.log 2>gpurun_out/bench_$w.err; done
python bench.py --steps ${STEPS:-30} --no-cpu-baseline --no-layer-cse > gpurun_out/bench_c2_nocse.log 2>&1
