#!/bin/bash
# iteration pass: GPU tests, then device-timed C2 (default and --no-layer-cse) and C4
mkdir -p gpurun_out
if [ "${TESTS:-1}" = "1" ]; then
  ( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
  tail -4 gpurun_out/pytest_gpu.log
fi
Q="--steps ${STEPS:-50} --warmup 3 --no-cpu-baseline --no-configs --no-sustained --no-ablation"
for v in "c2:" "nocse:--no-layer-cse" "c4:--workload c4" ${EXTRA:-}; do
  n=${v%%:*}; a=${v#*:}
  python bench.py $Q $a > gpurun_out/bench_$n.log 2> gpurun_out/bench_$n.err; echo "$n rc=$?"
  python scripts/show_bench.py gpurun_out/bench_$n.log 2>/dev/null || tail -c 600 gpurun_out/bench_$n.err
done
