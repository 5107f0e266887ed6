#!/bin/bash
# GPU tests + plain benches of the workloads in $WLS (no CPU baseline)
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
for w in ${WLS:-c2 c3 c4 c5}; do
  python bench.py --steps ${STEPS:-30} --warmup 3 --no-cpu-baseline --workload $w > gpurun_out/bench_$w.log 2> gpurun_out/bench_$w.err
done
python scripts/show_bench.py
